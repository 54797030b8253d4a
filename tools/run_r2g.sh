set -x
timeout 300 python tools/race_run.py log > gpurun_out/g1.log 2>&1
timeout 300 python tools/race_run.py multicontact >> gpurun_out/g1.log 2>&1
timeout 600 compute-sanitizer --tool racecheck --print-limit 8 python tools/race_run.py log 2>&1 | grep -v "^$" | tail -40 > gpurun_out/g2.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 8 python tools/race_run.py multicontact 2>&1 | grep -v "^$" | tail -30 > gpurun_out/g3.log
timeout 600 python tools/occ_sweep.py 32768 > gpurun_out/e_sweep.log 2>&1
cat gpurun_out/g1.log gpurun_out/g2.log gpurun_out/g3.log gpurun_out/e_sweep.log
