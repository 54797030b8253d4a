set -x
timeout 900 python tools/struct_debug.py 65536 > gpurun_out/c2.log 2>&1
timeout 600 python tools/occ_sweep.py 32768 > gpurun_out/e_sweep.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/c4.log
cat gpurun_out/c2.log gpurun_out/e_sweep.log gpurun_out/c4.log
