set -x
for c in 1 5; do
FCCQP_CTAS_PER_SM=$c FCCQP_STRUCT_REFINE=0 FCCQP_PROFILE=1 FCCQP_LIB=$PWD/fcc_qp_b200/libfccqp_b200_dev.so timeout 300 python tools/prof_run.py 32768 2 cold > gpurun_out/f_prof_c$c.log 2>&1
done
FCCQP_CTAS_PER_SM=1 FCCQP_STRUCT_REFINE=1 FCCQP_PROFILE=1 FCCQP_LIB=$PWD/fcc_qp_b200/libfccqp_b200_dev.so timeout 300 python tools/prof_run.py 32768 2 cold > gpurun_out/f_prof_c1_refine.log 2>&1
timeout 900 python -m pytest tests/test_gpu_structure.py tests/test_replay.py -q -x 2>&1 | grep -E "^E|assert|Error|^tests|passed|failed" | head -40 > gpurun_out/f_t1.log
timeout 900 python -m pytest tests/test_gpu_structure.py -q -k "host_arrays or warm_sequence or multicontact" 2>&1 | grep -E "^E|assert|Error|^tests|passed|failed" | head -60 > gpurun_out/f_t2.log
cat gpurun_out/f_prof_c1.log gpurun_out/f_prof_c5.log gpurun_out/f_prof_c1_refine.log gpurun_out/f_t1.log gpurun_out/f_t2.log
