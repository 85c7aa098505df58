#!/bin/bash
# A/B timing of decode-kernel builds on one box: every speaksense_b200/lib/variants/*.so in turn (no profiling counters)
# usage: tools/ab.sh [steps] [reps]
cd "$(dirname "$0")/.."
MAIN=speaksense_b200/lib/libspeaksense_whisper.so
cp $MAIN /tmp/main.so
for rep in $(seq 1 ${2:-2}); do
for v in speaksense_b200/lib/variants/*.so; do
  cp $v $MAIN
  echo "== $(basename $v) $(SS_NO_PROF=1 python tools/mega_prof.py large-v3 ${1:-64} 2>&1 | grep -E "ms/step|result-hash" | tail -4 | tr '\n' ' ')"
done
done
cp /tmp/main.so $MAIN
