# ncu --set full (source import) of one batched launch + per-line export; see profiles/
ncu --set full --clock-control none --import-source on -k regex:fccqp_solve -s 1 -c 1 -f -o gpurun_out/prof python tools/prof_run.py 16384 2 > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/prof.ncu-rep --page source --csv --print-source sass,cuda > gpurun_out/prof_src.csv 2>/dev/null
ncu -i gpurun_out/prof.ncu-rep --page raw --csv > gpurun_out/prof_raw.csv 2>/dev/null
tail -3 gpurun_out/ncu_full.log
