# quick GPU check: parity tests + bench line + phase profile + traces (instrumented build)
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/q_pytest.log
cat gpurun_out/q_pytest.log
bash tools/run_quick2.sh
