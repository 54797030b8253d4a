# GPU check: parity tests + bench line + reference arm
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
nproc; lscpu | grep "Model name"
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/c_pytest.log
python bench.py > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c_bench_ref.json 2>&1
cat gpurun_out/c_pytest.log gpurun_out/c_bench.json gpurun_out/c_bench_ref.json; tail -3 gpurun_out/c_bench.err
