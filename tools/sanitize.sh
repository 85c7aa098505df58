#!/bin/bash
# compute-sanitizer over every kernel family on the smallest model shapes (tools/sanitize_run.py); run on a GPU box:
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/sanitize.sh r2'
# memcheck: out-of-bounds / misaligned accesses of global, shared and local memory, leaks of device allocations are not checked.
# The plain run first gives the digests the sanitized run has to reproduce.
cd "$(dirname "$0")/.."
TAG=${1:-check}; OUT=gpurun_out/${TAG}_sanitizer.txt
mkdir -p gpurun_out
{
echo "== plain run"; timeout 300 python tools/sanitize_run.py 2>&1 | tail -12
echo "== compute-sanitizer --tool memcheck"
timeout 800 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python tools/sanitize_run.py 2>&1 | grep -v "^$" | tail -30
echo "memcheck exit code: ${PIPESTATUS[0]}"
# racecheck is opt-in (second argument "race"): it needs > 10 minutes inside the persistent decode kernel and its only finding is the
# TMA-write / ldmatrix-read pair ordered by mbarriers, which the tool does not model (profiles/r2_sanitizer.txt)
for TOOL in synccheck $([ "$2" = race ] && echo racecheck); do
  echo "== compute-sanitizer --tool $TOOL"
  timeout 800 compute-sanitizer --tool $TOOL --error-exitcode 7 --print-limit 12 python tools/sanitize_run.py 2>&1 | grep -v "^$" | tail -40
  echo "$TOOL exit code: ${PIPESTATUS[0]}"
done
} > $OUT 2>&1
cat $OUT
