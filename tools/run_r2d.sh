set -x
FCCQP_STRUCT_REFINE=${REFINE:-0} ncu --set full --clock-control none --import-source on -k regex:fccqp_struct_kernel -s 1 -c 1 -f -o gpurun_out/d_prof python tools/prof_run.py 16384 2 > gpurun_out/d_ncu.log 2>&1
ncu -i gpurun_out/d_prof.ncu-rep --page raw --csv > gpurun_out/d_prof_raw.csv 2>/dev/null
ncu -i gpurun_out/d_prof.ncu-rep --page source --csv --print-source sass,cuda > gpurun_out/d_prof_src.csv 2>/dev/null
ncu -i gpurun_out/d_prof.ncu-rep --page details > gpurun_out/d_prof_details.txt 2>/dev/null
rm -f gpurun_out/d_prof.ncu-rep
tail -5 gpurun_out/d_ncu.log
