#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_target.py > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|W256|small|W300|cell|native|CLSTM" gpurun_out/r2_sanitizer_$tool.log | tail -14
done
