# round 2, first GPU pass: parity of the structure-exploiting kernel, then timings
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests/test_gpu_structure.py -x -q 2>&1 | tail -30 > gpurun_out/a_pytest_struct.log
timeout 900 python tools/struct_debug.py 65536 > gpurun_out/a_struct_debug.log 2>&1
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python tools/sanitize_run.py walking 2>&1 | grep -v "^$" | tail -12 > gpurun_out/a_memcheck.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/a_pytest_all.log
cat gpurun_out/a_pytest_struct.log gpurun_out/a_struct_debug.log gpurun_out/a_memcheck.log gpurun_out/a_pytest_all.log
