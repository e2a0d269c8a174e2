#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=${1:-r2}
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_target.py > gpurun_out/${R}_sanitizer_$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|W256|small|W300|cell|native|CLSTM|pair" gpurun_out/${R}_sanitizer_$tool.log | tail -16
done
