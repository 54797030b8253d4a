set -x
# solo phase profile on a log WITHOUT iterating QPs would be ideal; max_iter cannot be 0, so use the first 256 QPs (all exit at 0?) tiled
FCCQP_CTAS_PER_SM=1 FCCQP_STRUCT_REFINE=0 FCCQP_PROFILE=1 FCCQP_LIB=$PWD/fcc_qp_b200/libfccqp_b200_dev.so timeout 300 python tools/prof_run.py 32768 2 cold 2>&1 | tail -16 > gpurun_out/k1.log
FCCQP_CTAS_PER_SM=1 FCCQP_STRUCT_REFINE=1 FCCQP_PROFILE=1 FCCQP_LIB=$PWD/fcc_qp_b200/libfccqp_b200_dev.so timeout 300 python tools/prof_run.py 32768 2 cold 2>&1 | tail -16 > gpurun_out/k2.log
cat gpurun_out/k1.log gpurun_out/k2.log
