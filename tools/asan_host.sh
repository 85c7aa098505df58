#!/bin/bash
# Host side of the library (model loader, C ABI, engine / batch engine logic) under AddressSanitizer + UBSan, CPU only:
#   bash tools/asan_host.sh            -> profiles/r2_asan_host.txt
# The .cc files are recompiled with -fsanitize=address,undefined, linked with the regular kernel objects into /tmp/ss_asan/, and the
# CPU test files that drive the loader / ABI error paths (corrupt and truncated model files, quantised files, header checks, text
# rules) run against that build with libasan preloaded into python.
set -e
cd "$(dirname "$0")/.."
python speaksense_b200/build.py > /dev/null
mkdir -p /tmp/ss_asan
OBJS=""
for f in api.cc engine.cc engine_batch.cc model.cc; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O1 -g -std=c++17 -Xcompiler -fPIC,-fsanitize=address,-fsanitize=undefined,-fno-omit-frame-pointer \
       --cudart static -x cu -Ispeaksense_b200/csrc -c speaksense_b200/csrc/$f -o /tmp/ss_asan/$f.o 2>&1 | grep -v deprecated || true
  OBJS="$OBJS /tmp/ss_asan/$f.o"
done
KOBJS=$(ls speaksense_b200/build/*.cu.o)
nvcc -shared -o /tmp/ss_asan/libspeaksense_whisper_asan.so $OBJS $KOBJS --cudart static -ldl -lpthread -Xcompiler -fsanitize=address,-fsanitize=undefined 2>&1 | grep -v deprecated || true
ASAN=$(gcc -print-file-name=libasan.so); UBSAN=$(gcc -print-file-name=libubsan.so)
OUT=profiles/r2_asan_host.txt
{
echo "host .cc files built with -fsanitize=address,undefined (tools/asan_host.sh), CPU-only tests against that build:"
LD_PRELOAD="$ASAN $UBSAN" ASAN_OPTIONS=detect_leaks=0:abort_on_error=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 \
  SS_LIB_PATH=/tmp/ss_asan/libspeaksense_whisper_asan.so \
  python -m pytest tests/test_abi.py tests/test_quantized.py tests/test_synth.py -x -q -m "not gpu" -p no:cacheprovider 2>&1 | tail -15
} > $OUT 2>&1
cat $OUT
