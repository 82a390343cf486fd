#!/bin/bash
# compute-sanitizer over the small-shape kernel cases (tools/sanitizer_cases.py); logs -> gpurun_out/sanitizer_<tool>.log
# usage: tools/run_sanitizers.sh [tool ...]   (default: memcheck racecheck synccheck)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/sanitizer_cases.py > gpurun_out/sanitizer_plain.log 2>&1 || { echo "cases fail WITHOUT sanitizer"; tail -5 gpurun_out/sanitizer_plain.log; exit 1; }
tools=${*:-memcheck racecheck synccheck}
rc=0
for t in $tools; do
  timeout "${LLB_SANITIZER_TIMEOUT:-900}" compute-sanitizer --tool "$t" --print-limit 30 --error-exitcode 9 \
    --kernel-name-exclude kernel_substring=at::native --log-file "gpurun_out/sanitizer_$t.raw" \
    python tools/sanitizer_cases.py > "gpurun_out/sanitizer_$t.log" 2>&1
  code=$?
  { echo "== compute-sanitizer --tool $t: exit $code"; grep -c "^========= .*\(Error\|Hazard\|Race\|Invalid\)" "gpurun_out/sanitizer_$t.raw" | sed 's/^/error lines: /'; \
    grep "ERROR SUMMARY\|RACECHECK SUMMARY\|LEAK SUMMARY" "gpurun_out/sanitizer_$t.raw"; tail -3 "gpurun_out/sanitizer_$t.log"; } | tee -a gpurun_out/sanitizer_summary.txt
  head -c 200000 "gpurun_out/sanitizer_$t.raw" > "gpurun_out/sanitizer_$t.txt"; rm -f "gpurun_out/sanitizer_$t.raw"
  [ $code -ne 0 ] && rc=$code
done
exit $rc
