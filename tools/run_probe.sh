set -x
export DEV=$PWD/fcc_qp_b200/libfccqp_b200_dev.so
# 1. uncontended trace of one QP
FCCQP_LIB=$DEV FCCQP_TRACE=gpurun_out/trace_b1.txt python tools/prof_run.py 1 1 > gpurun_out/probe_trace.log 2>&1
FCCQP_LIB=$DEV FCCQP_TRACE=gpurun_out/trace_b64k.txt python tools/prof_run.py 65536 1 >> gpurun_out/probe_trace.log 2>&1
# 2. parity without the refinement step
FCCQP_PRESOLVE_REFINE=0 python tools/gpu_check.py --quick > gpurun_out/probe_norefine.log 2>&1
FCCQP_PRESOLVE_REFINE=0 python -m pytest tests -m gpu -q 2>&1 | tail -15 >> gpurun_out/probe_norefine.log
# 3. raw H2D bandwidth
python - > gpurun_out/probe_h2d.log 2>&1 <<'PY'
import torch, time
for mb in (64, 256, 1024, 3072):
    h = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    d = torch.empty(mb << 20, dtype=torch.uint8, device="cuda")
    d.copy_(h, non_blocking=True); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    print(f"H2D {mb} MiB: {mb/1024/dt:.1f} GiB/s")
    t0 = time.perf_counter()
    for _ in range(3): h.copy_(d, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    print(f"D2H {mb} MiB: {mb/1024/dt:.1f} GiB/s")
import subprocess
print(subprocess.run("nvidia-smi -q | grep -A8 'GPU Link Info'; lscpu | head -20; numactl -H 2>/dev/null | head", shell=True, capture_output=True, text=True).stdout)
PY
# 4. ubench
cd tools/ubench && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_ubench fp64_ubench.cu && ./fp64_ubench > ../../gpurun_out/probe_ubench.log 2>&1
