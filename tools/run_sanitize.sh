# compute-sanitizer passes over small batches (Cassie log: 128-thread kernel; multicontact: 256-thread kernel)
for tool in memcheck racecheck synccheck; do
  for shape in walking multicontact quadruped odd; do
    echo "== $tool $shape"
    timeout 600 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_run.py $shape 2>&1 | grep -v "^$" | tail -6
  done
done
