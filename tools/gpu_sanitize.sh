#!/bin/bash
# compute-sanitizer over the wide path (tools/sanitize_wide.py), once with long rows in the batch and once with labels only.
mkdir -p gpurun_out
for short in 0 1; do
for tool in memcheck synccheck racecheck; do
  echo "== $tool (OWC_SANITIZE_SHORT=$short)"
  OWC_SANITIZE_SHORT=$short timeout 900 compute-sanitizer --tool $tool python tools/sanitize_wide.py 2>&1 | grep -v "^=========  *at\|^=========  *by\|Host Frame\|Device Frame" | tail -25
done
done > gpurun_out/r2_sanitize_wide.log 2>&1
grep "^==\|ERROR SUMMARY\|RACECHECK SUMMARY\|^ok\|Race reported" gpurun_out/r2_sanitize_wide.log | cut -c1-200
