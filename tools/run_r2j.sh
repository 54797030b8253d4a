set -x
for v in "FCCQP_STRUCT_PREFETCH=0" "FCCQP_STRUCT_BULK=0" "FCCQP_STRUCT_PREFETCH=0 FCCQP_STRUCT_BULK=0"; do
echo "== $v" >> gpurun_out/j.log
env $v FCCQP_CTAS_PER_SM=1 FCCQP_STRUCT_REFINE=0 FCCQP_PROFILE=1 FCCQP_LIB=$PWD/fcc_qp_b200/libfccqp_b200_dev.so timeout 300 python tools/prof_run.py 32768 2 cold 2>&1 | tail -13 | head -5 >> gpurun_out/j.log
env $v FCCQP_STRUCT_REFINE=0 timeout 300 python tools/prof_run.py 65536 3 cold 2>&1 | tail -1 >> gpurun_out/j.log
done
cat gpurun_out/j.log
