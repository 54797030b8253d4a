set -x
./tools/ubench/h2d_ubench > gpurun_out/h2d_ubench.log 2>&1
FCCQP_LIB=$PWD/fcc_qp_b200/libfccqp_b200_dev.so FCCQP_PROFILE=1 python tools/prof_run.py 65536 2 2>&1 | tail -17 > gpurun_out/q_phase.log
FCCQP_LIB=$PWD/fcc_qp_b200/libfccqp_b200_dev.so FCCQP_PROFILE=1 python tools/prof_run.py 65536 2 warm 2>&1 | tail -17 > gpurun_out/q_phase_warm.log
cat gpurun_out/h2d_ubench.log gpurun_out/q_phase.log gpurun_out/q_phase_warm.log
