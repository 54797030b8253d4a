set -x
timeout 900 python tools/struct_debug2.py > gpurun_out/c1.log 2>&1
timeout 900 python tools/struct_debug.py 65536 > gpurun_out/c2.log 2>&1
FCCQP_STRUCT_PREFETCH=0 timeout 300 python tools/prof_run.py 65536 4 cold > gpurun_out/c3.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/c4.log
cat gpurun_out/c1.log gpurun_out/c2.log gpurun_out/c3.log gpurun_out/c4.log
