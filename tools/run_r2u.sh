# quick loop: structure + parity tests, then timings of the four shapes (device-resident, cold)
set -x
timeout 600 python -m pytest tests/test_gpu_structure.py tests/test_gpu_parity.py tests/test_adaptive_rho.py tests/test_relaxation.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/u_t.log
timeout 300 python tools/prof_run.py 65536 4 cold 2>&1 | tail -2 > gpurun_out/u_time.log
timeout 300 python tools/prof_shape.py humanoid 65536 3 cold 2>&1 | tail -2 >> gpurun_out/u_time.log
timeout 300 python tools/prof_shape.py multicontact 16384 3 cold 2>&1 | tail -2 >> gpurun_out/u_time.log
timeout 300 python tools/prof_shape.py quadruped 131072 3 cold 2>&1 | tail -2 >> gpurun_out/u_time.log
cat gpurun_out/u_t.log gpurun_out/u_time.log
