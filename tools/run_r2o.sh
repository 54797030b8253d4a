set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k paper_settings 2>&1 | grep -E "^E|Error|assert" | head -20 > gpurun_out/o_t1.log
for fo in 1 0; do
for c in 1 4; do
echo "== humanoid FULLOP=$fo CTAS=$c" >> gpurun_out/o_prof.log
FCCQP_STRUCT_FULLOP=$fo FCCQP_CTAS_PER_SM=$c FCCQP_PROFILE=1 FCCQP_LIB=$PWD/fcc_qp_b200/libfccqp_b200_dev.so timeout 300 python tools/prof_shape.py humanoid 16384 2 cold 2>&1 | tail -16 >> gpurun_out/o_prof.log
done
done
cat gpurun_out/o_t1.log gpurun_out/o_prof.log
