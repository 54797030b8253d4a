# A/B of two builds of the library on the same box: FCCQP_LIB=libfccqp_b200_prev.so against the in-tree one, interleaved
set -x
P=$PWD/fcc_qp_b200/libfccqp_b200_prev.so
for r in 1 2; do
  for lib in prev new; do
    if [ $lib = prev ]; then export FCCQP_LIB=$P; else unset FCCQP_LIB; fi
    echo "== $lib" >> gpurun_out/ab.log
    timeout 300 python tools/prof_run.py 65536 4 cold 2>&1 | tail -1 >> gpurun_out/ab.log
    timeout 300 python tools/prof_shape.py humanoid 32768 3 cold 2>&1 | tail -1 >> gpurun_out/ab.log
    timeout 300 python tools/prof_shape.py multicontact 16384 3 cold 2>&1 | tail -1 >> gpurun_out/ab.log
    timeout 300 python tools/prof_shape.py quadruped 65536 3 cold 2>&1 | tail -1 >> gpurun_out/ab.log
  done
done
cat gpurun_out/ab.log
