#!/bin/bash
# compute-sanitizer sweep of one small pair through the whole path (run on the GPU box):
#   bash tools/sanitize.sh        # memcheck (even and odd widths), racecheck, synccheck, initcheck
cd "$(dirname "$0")/.."
for spec in "memcheck 3 64 48" "memcheck 3 50 37" "racecheck 3 96 72" "synccheck 3 64 48" "initcheck 2 64 48"; do
  set -- $spec
  echo "=== $1 (L=$2, lowest $3x$4)"
  compute-sanitizer --tool "$1" --print-limit 6 python tools/prof_pair.py "$2" "$3" "$4" 1 2>&1 | grep -E "SUMMARY|Error|points" | tail -6
done
# the sink filter (sink.cu) on a small surface cloud with isolated points (ring widening) and on a sparse cloud
for tool in memcheck racecheck initcheck; do
  echo "=== $tool sink (surface 120x120, spacing 0.4)"
  compute-sanitizer --tool "$tool" --print-limit 6 python tools/prof_sink.py 120 0.4 2>&1 | grep -E "SUMMARY|Error|points" | tail -4
done
echo "=== memcheck sink (sparse: spacing 6, most rings widen)"
compute-sanitizer --tool memcheck --print-limit 6 python tools/prof_sink.py 40 6.0 2>&1 | grep -E "SUMMARY|Error|points" | tail -4
