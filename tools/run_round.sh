# Round evidence: parity tests, both bench arms, ncu launch list of the bench command, ncu --set full of one launch
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
nproc; lscpu | grep "Model name"
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r_pytest.log
python bench.py > gpurun_out/r_bench.json 2> gpurun_out/r_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r_bench_ref.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r_ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fccqp_solve -s 1 -c 1 -f -o gpurun_out/r_prof python tools/prof_run.py 16384 2 > gpurun_out/r_ncu_full.log 2>&1
ncu -i gpurun_out/r_prof.ncu-rep --page raw --csv > gpurun_out/r_prof_raw.csv 2>/dev/null
ncu -i gpurun_out/r_prof.ncu-rep --page source --csv --print-source sass,cuda > gpurun_out/r_prof_src.csv 2>/dev/null
rm -f gpurun_out/r_prof.ncu-rep
cat gpurun_out/r_pytest.log gpurun_out/r_bench.json gpurun_out/r_bench_ref.json
