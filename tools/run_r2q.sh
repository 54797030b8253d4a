# ncu --set full of the structure-exploiting kernel on three shapes + phase profile at full occupancy
set -x
prof() {  # tag, command...
  tag=$1; shift
  ncu --set full --clock-control none --import-source on -k regex:fccqp_struct_kernel -s 1 -c 1 -f -o gpurun_out/$tag "$@" > gpurun_out/${tag}_ncu.log 2>&1
  ncu -i gpurun_out/$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
  ncu -i gpurun_out/$tag.ncu-rep --page source --csv --print-source sass,cuda > gpurun_out/${tag}_src.csv 2>/dev/null
  ncu -i gpurun_out/$tag.ncu-rep --page details > gpurun_out/${tag}_details.txt 2>/dev/null
  rm -f gpurun_out/$tag.ncu-rep
}
prof q_cassie python tools/prof_run.py 16384 2
prof q_humanoid python tools/prof_shape.py humanoid 16384 2
prof q_multicontact python tools/prof_shape.py multicontact 8192 2
for c in 1 5; do
FCCQP_CTAS_PER_SM=$c FCCQP_PROFILE=1 FCCQP_LIB=$PWD/fcc_qp_b200/libfccqp_b200_dev.so timeout 300 python tools/prof_run.py 65536 2 cold 2>&1 | tail -18 > gpurun_out/q_phase_c$c.log
done
FCCQP_PROFILE=1 FCCQP_LIB=$PWD/fcc_qp_b200/libfccqp_b200_dev.so timeout 300 python tools/prof_shape.py multicontact 16384 2 cold 2>&1 | tail -18 > gpurun_out/q_phase_mc.log
cat gpurun_out/q_phase_c1.log gpurun_out/q_phase_c5.log gpurun_out/q_phase_mc.log
