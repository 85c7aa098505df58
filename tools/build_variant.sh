#!/bin/bash
# A/B variants of one kernel file: tools/build_variant.sh <name> [source.cu] [extra nvcc flags...]
#   links a copy of the library with that file's object replaced into speaksense_b200/lib/variants/<name>.so (decode kernel: timed by tools/ab.sh)
set -e
cd "$(dirname "$0")/.."
NAME=$1; SRC=${2:-speaksense_b200/csrc/decoder_mega.cu}; shift; shift || true
python speaksense_b200/build.py > /dev/null
mkdir -p speaksense_b200/lib/variants /tmp/ssv
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function --cudart static -x cu \
     -Ispeaksense_b200/csrc "$@" -c $SRC -o /tmp/ssv/$NAME.o 2>&1 | grep -v deprecated || true
OBJS=$(ls speaksense_b200/build/*.o | grep -v "/$(basename $SRC).o")
nvcc -shared -o speaksense_b200/lib/variants/$NAME.so $OBJS /tmp/ssv/$NAME.o --cudart static -ldl -lpthread 2>&1 | grep -v deprecated || true
ls -la speaksense_b200/lib/variants/$NAME.so
