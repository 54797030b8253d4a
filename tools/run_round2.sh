# Round-2 evidence in one gpurun call: parity tests, both bench arms, ncu launch list of the bench command, ncu --set full of
# the structure-exploiting kernel on three shapes and of the warp kernel, DRAM traffic at the bench batch, small-QP rates.
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
nproc; lscpu | grep "Model name"
python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r2_pytest.log
python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/r2_ncu_b.log 2>&1
prof() {  # tag, kernel regex, command...
  tag=$1; k=$2; shift; shift
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/$tag "$@" > gpurun_out/${tag}_ncu.log 2>&1
  ncu -i gpurun_out/$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
  ncu -i gpurun_out/$tag.ncu-rep --page source --csv --print-source sass,cuda > gpurun_out/${tag}_src.csv 2>/dev/null
  rm -f gpurun_out/$tag.ncu-rep
}
prof r2_cassie fccqp_struct_kernel python tools/prof_run.py 16384 2
prof r2_humanoid fccqp_struct_kernel python tools/prof_shape.py humanoid 16384 2
prof r2_multicontact fccqp_struct_kernel python tools/prof_shape.py multicontact 8192 2
prof r2_warp fccqp_warp_kernel python tools/bench_small.py 65536
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:fccqp_struct_kernel -s 1 -c 1 --csv --log-file gpurun_out/r2_dram65536.csv python tools/prof_run.py 65536 2 > /dev/null 2>&1
python tools/bench_small.py > gpurun_out/r2_small.jsonl 2>/dev/null
FCCQP_NO_WARP=1 python tools/bench_small.py >> gpurun_out/r2_small.jsonl 2>/dev/null
cat gpurun_out/r2_pytest.log gpurun_out/r2_bench.json gpurun_out/r2_bench_ref.json; tail -3 gpurun_out/r2_dram65536.csv
# processing order from the previous solve (FCCQP_SCHEDULE_LPT) against the default order, same process conditions
python tools/prof_run.py 65536 5 cold 2>&1 | tail -3 > gpurun_out/r2_lpt.log
LPT=1 python tools/prof_run.py 65536 5 cold 2>&1 | tail -3 >> gpurun_out/r2_lpt.log
cat gpurun_out/r2_lpt.log
