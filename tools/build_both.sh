# production + developer (FCCQP_DEV instrumentation) builds of the CUDA library
set -e
cd "$(dirname "$0")/.."
python -m fcc_qp_b200.build 2>&1 | grep -E "error|^\+" || true
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -DFCCQP_DEV \
  -o fcc_qp_b200/libfccqp_b200_dev.so fcc_qp_b200/csrc/fccqp_capi.cu 2>&1 | grep -E "error" && exit 1 || true
ls -la fcc_qp_b200/*.so
