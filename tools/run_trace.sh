# uncontended and contended event traces of one QP (developer build), plus the FP64 microbenchmarks
export DEV=$PWD/fcc_qp_b200/libfccqp_b200_dev.so
FCCQP_LIB=$DEV FCCQP_TRACE=gpurun_out/trace_b1.txt python tools/prof_run.py 1 1 > gpurun_out/probe_trace.log 2>&1
FCCQP_LIB=$DEV FCCQP_TRACE=gpurun_out/trace_b64k.txt python tools/prof_run.py 65536 1 >> gpurun_out/probe_trace.log 2>&1
cd tools/ubench && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_ubench fp64_ubench.cu && ./fp64_ubench > ../../gpurun_out/probe_ubench.log 2>&1
