set -x
timeout 600 python -m pytest tests/test_gpu_structure.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/t_t.log
timeout 300 python tools/prof_run.py 65536 4 cold 2>&1 | tail -2 > gpurun_out/t_time.log
timeout 300 python tools/prof_shape.py humanoid 65536 3 cold 2>&1 | tail -2 >> gpurun_out/t_time.log
timeout 300 python tools/prof_shape.py multicontact 16384 3 cold 2>&1 | tail -2 >> gpurun_out/t_time.log
timeout 300 python tools/prof_shape.py quadruped 131072 3 cold 2>&1 | tail -2 >> gpurun_out/t_time.log
FCCQP_PROFILE=1 FCCQP_LIB=$PWD/fcc_qp_b200/libfccqp_b200_dev.so timeout 300 python tools/prof_shape.py multicontact 16384 2 cold 2>&1 | tail -16 > gpurun_out/t_phase_mc.log
FCCQP_PROFILE=1 FCCQP_LIB=$PWD/fcc_qp_b200/libfccqp_b200_dev.so timeout 300 python tools/prof_shape.py humanoid 16384 2 cold 2>&1 | tail -16 > gpurun_out/t_phase_hu.log
cat gpurun_out/t_t.log gpurun_out/t_time.log gpurun_out/t_phase_mc.log gpurun_out/t_phase_hu.log
