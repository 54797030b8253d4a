set -x
FCCQP_CTAS_PER_SM=1 FCCQP_STRUCT_REFINE=0 FCCQP_PROFILE=1 FCCQP_LIB=$PWD/fcc_qp_b200/libfccqp_b200_dev.so timeout 300 python tools/prof_run.py 32768 2 cold 2>&1 | tail -13 > gpurun_out/i_prof_c1.log
timeout 900 python tools/struct_debug.py 65536 2>&1 | grep "^time" > gpurun_out/i_time.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/i_t1.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 8 python tools/race_run.py log 2>&1 | grep -v "^$" | tail -6 > gpurun_out/i_race.log
cat gpurun_out/i_prof_c1.log gpurun_out/i_time.log gpurun_out/i_t1.log gpurun_out/i_race.log
