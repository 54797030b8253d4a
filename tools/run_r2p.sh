set -x
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi_device.py -x -q 2>&1 | tail -15 > gpurun_out/p_t1.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/p_t2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/p_bench2.json 2> gpurun_out/p_bench2.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/p_bench1.json 2> gpurun_out/p_bench1.err
tail -5 gpurun_out/p_bench1.err gpurun_out/p_bench2.err
cat gpurun_out/p_t1.log gpurun_out/p_t2.log gpurun_out/p_bench2.json gpurun_out/p_bench1.json
