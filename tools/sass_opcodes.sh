#!/bin/bash
# Opcode histogram of the built objects (evidence of the Blackwell-native path: UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st,
# UTMALDG/UTMASTG/UBLKCP = TMA, HMMA = legacy mma.sync).   tools/sass_opcodes.sh > profiles/sass_opcodes.txt
cd "$(dirname "$0")/.."
python speaksense_b200/build.py > /dev/null
echo "cuobjdump -sass speaksense_b200/build/*.o | grep -oE '<mnemonic>' | sort | uniq -c   ($(date -u +%Y-%m-%dT%H:%MZ), nvcc $(nvcc --version | grep -oE 'V[0-9]+\.[0-9]+\.[0-9]+' | head -1))"
for o in speaksense_b200/build/*.cu.o; do
  echo "== $(basename $o)  (source sha256[:16] $(sha256sum speaksense_b200/csrc/$(basename $o .o) | cut -c1-16))"
  cuobjdump -sass $o | grep -oE '\b(UTC[A-Z]*MMA[A-Z0-9_.]*|UTCBAR[A-Z0-9_.]*|UTCATOMSWS[A-Z0-9_.]*|LDTM[A-Z0-9_.]*|STTM[A-Z0-9_.]*|UTMALDG[A-Z0-9_.]*|UTMASTG[A-Z0-9_.]*|UBLKCP[A-Z0-9_.]*|HMMA[A-Z0-9_.]*|LDSM[A-Z0-9_.]*|LDGSTS[A-Z0-9_.]*|SYNCS[A-Z0-9_.]*|UCGABAR[A-Z0-9_.]*|REDUX[A-Z0-9_.]*|MUFU\.[A-Z0-9_.]*)' | sort | uniq -c | sort -rn | awk '{printf "  %7d %s\n", $1, $2}'
done
