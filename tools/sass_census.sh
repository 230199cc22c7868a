#!/bin/bash
# SASS opcode census of the shipped library, per kernel: the Blackwell-native mnemonics (B200_PROFILING.md)
#   UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor loads/stores, UBLKCP = cp.async.bulk,
#   LDGSTS = cp.async, HMMA = legacy mma.sync (must be 0), SYNCS = mbarrier, FMNMX3/FMUL2/FFMA2 = 3-input max / packed fp32x2
# usage: tools/sass_census.sh [lib.so] > profiles/rN_sass_census.txt
LIB=${1:-dcfa-yolo_b200/lib/libdcfa_b200.so}
cuobjdump -sass "$LIB" | awk '
  /Function :/ { fn=$3; order[++n]=fn; next }
  fn != "" {
    if ($0 ~ /UTC[A-Z]*MMA/) c[fn,"UTCMMA"]++
    if ($0 ~ /LDTM/) c[fn,"LDTM"]++
    if ($0 ~ /UTMALDG/) c[fn,"UTMALDG"]++
    if ($0 ~ /UTMASTG/) c[fn,"UTMASTG"]++
    if ($0 ~ /UBLKCP/) c[fn,"UBLKCP"]++
    if ($0 ~ /LDGSTS/) c[fn,"LDGSTS"]++
    if ($0 ~ / HMMA/) c[fn,"HMMA"]++
    if ($0 ~ /SYNCS/) c[fn,"SYNCS"]++
    if ($0 ~ /FFMA2|FMUL2|FADD2/) c[fn,"F32X2"]++
    if ($0 ~ /FMNMX3/) c[fn,"FMNMX3"]++
    if ($0 ~ /UCGABAR|CGABAR/) c[fn,"CGABAR"]++
    if ($0 ~ /^ +\/\*[0-9a-f]+\*\/ /) c[fn,"total"]++
  }
  END {
    printf "%-60s %7s %6s %5s %7s %7s %6s %6s %5s %5s %6s %6s\n", "kernel", "instrs", "UTCMMA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "HMMA", "SYNCS", "F32X2", "CGABAR"
    for (i=1;i<=n;i++) { f=order[i]; name=f; gsub(/^_ZN4dcfa[0-9]*_GLOBAL__N__[0-9a-f_]*/,"",name);
      printf "%-60s %7d %6d %5d %7d %7d %6d %6d %5d %5d %6d %6d\n", substr(name,1,60), c[f,"total"], c[f,"UTCMMA"], c[f,"LDTM"], c[f,"UTMALDG"], c[f,"UTMASTG"], c[f,"UBLKCP"], c[f,"LDGSTS"], c[f,"HMMA"], c[f,"SYNCS"], c[f,"F32X2"], c[f,"CGABAR"] }
  }'
