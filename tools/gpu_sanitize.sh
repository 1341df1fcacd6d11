#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  echo "== $tool" 
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_wide.py 2>&1 | grep -v "^=========  *at\|^=========  *by\|Host Frame\|Device Frame" | tail -25
done > gpurun_out/r2_sanitize_wide.log 2>&1
tail -60 gpurun_out/r2_sanitize_wide.log
