set -x
for md in same copies; do
for f in "" "staged"; do
echo "== mode=$md FRONT=$f" >> gpurun_out/m_prof.log
FCCQP_STRUCT_FRONT=$f FCCQP_CTAS_PER_SM=1 FCCQP_STRUCT_REFINE=1 FCCQP_PROFILE=1 FCCQP_LIB=$PWD/fcc_qp_b200/libfccqp_b200_dev.so timeout 300 python tools/prof_same.py 32768 0 $md 2>&1 | tail -16 >> gpurun_out/m_prof.log
FCCQP_STRUCT_FRONT=$f FCCQP_STRUCT_REFINE=1 timeout 300 python tools/prof_same.py 65536 0 $md 2>&1 | tail -1 >> gpurun_out/m_prof.log
done
done
cat gpurun_out/m_prof.log
