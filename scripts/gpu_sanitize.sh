#!/bin/bash
# compute-sanitizer over a small-shape pass of every kernel family (scripts/sanitize_target.py).
# Each tool/section runs under its own timeout; summaries land in gpurun_out/sanitize_*.txt.
mkdir -p gpurun_out
timeout 120 python scripts/sanitize_target.py all > gpurun_out/sanitize_plain.txt 2>&1; echo "plain exit $?"; tail -3 gpurun_out/sanitize_plain.txt
SECTIONS=${SECTIONS:-"rr_host rr_device rd_host rd_device rr_rk4 rr_em rrt_exact rrt_fixed rrt_hutch tc_tile staged"}
for tool in ${TOOLS:-memcheck synccheck racecheck}; do
  out=gpurun_out/sanitize_$tool.txt; : > $out
  for sec in $SECTIONS; do
    echo "=== $tool $sec" >> $out
    timeout ${SAN_TIMEOUT:-240} compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_target.py $sec > gpurun_out/_san.tmp 2>&1
    rc=$?
    grep -E "^=========|^ok|^done|Error|error|hazard|Hazard|RACECHECK|ERROR SUMMARY" gpurun_out/_san.tmp | grep -v "^========= COMPUTE-SANITIZER$" | head -40 >> $out
    echo "exit $rc" >> $out
  done
  echo "== $tool"; grep -E "^=== |ERROR SUMMARY|exit |hazard" $out | paste - - - | head -40
done
rm -f gpurun_out/_san.tmp
