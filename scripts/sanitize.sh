set -x
for cfg in "31 1024 1" "31 1024 1 enc_mode=3" "31 1024 1 enc_mode=2" "31 1024 1 splitk=3" "16 768 1 enc_mode=3" "21 1024 1 splitk=2" "11 512 1 enc_mode=2" "15 1024 1 enc_mode=3"; do
  timeout 250 compute-sanitizer --tool memcheck --print-limit 3 --error-exitcode 7 python scripts/run_one.py $cfg 2>&1 | grep -E "ERROR SUMMARY|Invalid|Error|done|trap" | head -5
done
