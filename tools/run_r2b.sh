set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shared_structure.py -q -x -k "paper_settings" 2>&1 | grep -E "^E|assert|Error" | head -30 > gpurun_out/b1.log
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "full_size_properties or fp32_data_mode_walking_log" 2>&1 | grep -E "^E|assert|Error|^tests" | head -60 > gpurun_out/b2.log
timeout 900 python -m pytest tests/test_gpu_shared_structure.py -q 2>&1 | grep -E "^E|assert|Error|^tests" | head -60 > gpurun_out/b3.log
FCCQP_FULL_INVERSE_AT=100000 timeout 600 python tools/struct_debug2.py > gpurun_out/b4.log 2>&1
timeout 600 python tools/struct_debug2.py > gpurun_out/b5.log 2>&1
cat gpurun_out/b1.log gpurun_out/b2.log gpurun_out/b3.log gpurun_out/b4.log gpurun_out/b5.log
