#!/bin/sh
# run on the GPU box: every kernel of libp2v.so under compute-sanitizer (memcheck, then initcheck and racecheck on the quick set).
# Output: gpurun_out/sanitize_<tool>.log; exit code != 0 if a tool reports an error.
mkdir -p gpurun_out
rc=0
python tools/sanitize_run.py quick > gpurun_out/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/sanitize_plain.log; exit 1; }
tail -1 gpurun_out/sanitize_plain.log
for tool in memcheck ${SAN_TOOLS:-initcheck racecheck}; do
  timeout ${SAN_TIMEOUT:-420} compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 20 python tools/sanitize_run.py quick > gpurun_out/sanitize_$tool.log 2>&1
  r=$?
  echo "== $tool: exit $r"; grep -E "ERROR SUMMARY|sanitize_run ok|Error|error" gpurun_out/sanitize_$tool.log | head -12
  [ $r -ne 0 ] && rc=$r
done
exit $rc
