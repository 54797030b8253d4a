set -x
timeout 900 python -m pytest tests/test_gpu_structure.py -x -q 2>&1 | tail -15 > gpurun_out/n_t1.log
timeout 900 python tools/struct_debug.py 65536 > gpurun_out/n_dbg.log 2>&1
FCCQP_STRUCT_FULLOP=0 timeout 900 python tools/struct_debug.py 65536 2>&1 | grep "^time" > gpurun_out/n_time_nofullop.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/n_t2.log
cat gpurun_out/n_t1.log gpurun_out/n_dbg.log gpurun_out/n_time_nofullop.log gpurun_out/n_t2.log
