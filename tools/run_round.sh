set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r1_pytest.log
python bench.py > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1_bench_ref.json 2>&1
FCCQP_LIB=$PWD/fcc_qp_b200/libfccqp_b200_dev.so FCCQP_PROFILE=1 python tools/prof_run.py 65536 2 > gpurun_out/r1_phase.log 2>&1
FCCQP_LIB=$PWD/fcc_qp_b200/libfccqp_b200_dev.so FCCQP_PROFILE=1 python tools/prof_run.py 65536 2 warm > gpurun_out/r1_phase_warm.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r1_ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fccqp_solve -s 1 -c 1 -o gpurun_out/r1_prof python tools/prof_run.py 16384 2 > gpurun_out/r1_ncu_full.log 2>&1
cat gpurun_out/r1_pytest.log gpurun_out/r1_bench.json gpurun_out/r1_phase.log
