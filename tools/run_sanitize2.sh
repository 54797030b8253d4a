# compute-sanitizer over what round 2 added last: the compact operator loop of the structure kernel (multicontact: 128 threads,
# humanoid: 96 threads; both sets hold QPs that run to max_iter) and the FP32 instance of the warp kernel ("odd")
for tool in memcheck racecheck synccheck; do
  for shape in multicontact humanoid odd; do
    echo "== $tool $shape"
    timeout 600 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_run.py $shape 2>&1 | grep -v "^$" | tail -7
  done
done
