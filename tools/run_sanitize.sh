# compute-sanitizer passes over small batches (Cassie log: 128-thread kernel; multicontact: 256-thread kernel)
for tool in memcheck racecheck synccheck; do
  for shape in walking multicontact quadruped odd; do
    echo "== $tool $shape"
    timeout 600 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_run.py $shape 2>&1 | grep -v "^$" | tail -6
  done
done
# warp-per-QP kernel: "odd" above (n + m = 19); the CTA kernels on the same shape; adaptive rho on both mappings
for tool in memcheck racecheck; do
  echo "== $tool odd FCCQP_NO_WARP=1"
  FCCQP_NO_WARP=1 timeout 600 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_run.py odd 2>&1 | grep -v "^$" | tail -4
  for shape in odd walking; do
    echo "== $tool $shape adapt_rho_interval=5"
    SAN_ADAPT=5 timeout 600 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_run.py $shape 2>&1 | grep -v "^$" | tail -4
  done
done
