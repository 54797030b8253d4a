# quick GPU check: parity tests + bench line (+ phase profile with the instrumented build)
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/q_pytest.log
python bench.py > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err
FCCQP_LIB=$PWD/fcc_qp_b200/libfccqp_b200_dev.so FCCQP_PROFILE=1 python tools/prof_run.py 65536 2 2>&1 | tail -17 > gpurun_out/q_phase.log
FCCQP_LIB=$PWD/fcc_qp_b200/libfccqp_b200_dev.so FCCQP_PROFILE=1 python tools/prof_run.py 65536 2 warm 2>&1 | tail -17 > gpurun_out/q_phase_warm.log
cat gpurun_out/q_pytest.log; tail -3 gpurun_out/q_bench.err; python - <<'PY'
import json
l = json.load(open("gpurun_out/q_bench.json"))
print("value %.3f M QP/s  e2e %.3f M QP/s (%.1f ms)  cpu %.0f  fp64 frac %.3f" % (l["value"]/1e6, l["e2e"]["value"]/1e6, l["e2e"]["ms_per_step"], l["cpu_baseline"]["value"], l["roofline_fp64"]["frac"]))
PY
cat gpurun_out/q_phase.log gpurun_out/q_phase_warm.log
