# compute-sanitizer over the solution polish (prepare / finish kernels + the equality-constrained inner solve)
for tool in memcheck racecheck; do
  for shape in walking odd humanoid; do
    echo "== $tool $shape (polish)"
    SAN_POLISH=1 timeout 600 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_run.py $shape 2>&1 | grep -v "^$" | tail -4
  done
done
