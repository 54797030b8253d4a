# H2D rate per GPU under 1, 2, 4 (8) concurrent copiers, then the bench under torchrun on all GPUs of the box
G=${1:-4}
nproc; nvidia-smi -L | wc -l
for n in 1 2 4 8; do
  [ $n -le $G ] || continue
  for i in $(seq 0 $((n-1))); do python tools/h2d_concurrent.py $i $n & done; wait
done 2>&1 | grep concurrent | sort > gpurun_out/h2d_scaling.log
cat gpurun_out/h2d_scaling.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $G --steps 5 --warmup 3 > gpurun_out/bench_${G}gpu.json 2> gpurun_out/bench_${G}gpu.err
cat gpurun_out/bench_${G}gpu.json | cut -c1-1600
python -m pytest tests/test_gpu_multi_device.py -m gpu -q 2>&1 | tail -2
