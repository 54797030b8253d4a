set -x
timeout 900 python -m pytest tests/test_gpu_structure.py -x -q 2>&1 | tail -15 > gpurun_out/l_t1.log
for f in "" "staged"; do
echo "== FRONT=$f" >> gpurun_out/l_prof.log
for r in 0 1; do
FCCQP_STRUCT_FRONT=$f FCCQP_CTAS_PER_SM=1 FCCQP_STRUCT_REFINE=$r FCCQP_PROFILE=1 FCCQP_LIB=$PWD/fcc_qp_b200/libfccqp_b200_dev.so timeout 300 python tools/prof_run.py 32768 2 cold 2>&1 | tail -16 >> gpurun_out/l_prof.log
FCCQP_STRUCT_FRONT=$f FCCQP_STRUCT_REFINE=$r timeout 300 python tools/prof_run.py 65536 3 cold 2>&1 | tail -1 >> gpurun_out/l_prof.log
done
done
timeout 900 python tools/struct_debug.py 65536 2>&1 | grep "^time" > gpurun_out/l_time.log
cat gpurun_out/l_t1.log gpurun_out/l_prof.log gpurun_out/l_time.log
