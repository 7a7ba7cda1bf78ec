set -x
for cfg in "21 1024 1 epi_assist=0" "21 1024 1" "31 1024 1 splitk=3"; do
  timeout 250 compute-sanitizer --tool synccheck --print-limit 3 python scripts/run_one.py $cfg 2>&1 | grep -E "ERROR SUMMARY|Barrier error|done|trap|at ftsgemm" | head -4
done
