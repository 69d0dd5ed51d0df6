#!/bin/bash
# Per-object counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md):
# UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA load/store, UTCBAR = tcgen05.commit,
# HMMA would be the legacy mma.sync path (expected: 0).  Usage: scripts/sass_summary.sh > profiles/rN_sass_summary.txt
HERE="$(cd "$(dirname "$0")/.." && pwd)"
printf "%-26s %8s %8s %8s %8s %8s %8s %8s\n" object UTCxMMA UTCBAR LDTM STTM UTMALDG UTMASTG HMMA
for o in "$HERE"/elektronn2_b200/lib/*.o; do
  sass=$(cuobjdump -sass "$o" 2>/dev/null)
  c() { echo "$sass" | grep -c -E "$1"; }
  printf "%-26s %8d %8d %8d %8d %8d %8d %8d\n" "$(basename "$o")" "$(c 'UTC[A-Z]*MMA')" "$(c 'UTCBAR')" "$(c 'LDTM')" "$(c 'STTM')" "$(c 'UTMALDG')" "$(c 'UTMASTG')" "$(c '\bHMMA')"
done
echo "# built from elektronn2_b200/csrc at $(git -C "$HERE" rev-parse --short HEAD 2>/dev/null) with: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3"
