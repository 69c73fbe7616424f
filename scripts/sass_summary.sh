#!/bin/bash
# Blackwell-specific SASS mnemonics per object file of libscda_b200.so (cuobjdump -sass | grep -c):
# UTCHMMA = tcgen05.mma, UTMALDG/UTMASTG = TMA load/store, LDTM/STTM = tcgen05.ld/st (TMEM),
# UTCBAR = tcgen05.commit, SYNCS = mbarrier, ELECT = elect.sync, UBLKCP = cp.async.bulk
# usage: scripts/sass_summary.sh > profiles/rN_sass_summary.txt
cd "$(dirname "$0")/../scda_b200/csrc/build" || exit 1
printf "%-18s %8s %8s %8s %8s %8s %8s %8s %8s %8s %8s %8s\n" object UTCHMMA .2CTA UTMALDG UTMASTG LDTM UTCBAR SYNCS ELECT UBLKCP LDG.128 STG.128
for o in *.o; do
  s=$(cuobjdump -sass "$o" 2>/dev/null)
  c() { echo "$s" | grep -c -- "$1"; }
  printf "%-18s %8d %8d %8d %8d %8d %8d %8d %8d %8d %8d %8d\n" "${o%.o}.cu" "$(c UTCHMMA)" "$(c 'UTCHMMA.2CTA')" "$(c UTMALDG)" "$(c UTMASTG)" "$(c LDTM)" "$(c UTCBAR)" "$(c SYNCS)" "$(c ELECT)" "$(c UBLKCP)" "$(c 'LDG.E.128')" "$(c 'STG.E.128')"
done
echo
echo "# nvcc $(nvcc --version | tail -2 | head -1); flags: -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo"
echo "# commit $(git -C ../../.. rev-parse --short HEAD)"
