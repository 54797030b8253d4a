# quicker GPU check: bench value + phase profile + contended trace
python bench.py > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err
export DEV=$PWD/fcc_qp_b200/libfccqp_b200_dev.so
FCCQP_LIB=$DEV FCCQP_PROFILE=1 python tools/prof_run.py 65536 2 2>&1 | tail -17 > gpurun_out/q_phase.log
FCCQP_LIB=$DEV FCCQP_TRACE=gpurun_out/trace_b64k.txt python tools/prof_run.py 65536 1 > /dev/null 2>&1
FCCQP_CTAS_PER_SM=1 FCCQP_LIB=$DEV FCCQP_TRACE=gpurun_out/trace_b1.txt python tools/prof_run.py 592 1 > /dev/null 2>&1
python - <<'PY'
import json
l = json.load(open("gpurun_out/q_bench.json"))
print("value %.3f M QP/s  e2e %.3f M QP/s (%.1f ms)  cpu %.0f  fp64 frac %.3f" % (l["value"]/1e6, l["e2e"]["value"]/1e6, l["e2e"]["ms_per_step"], l["cpu_baseline"]["value"], l["roofline_fp64"]["frac"]))
PY
cat gpurun_out/q_phase.log
