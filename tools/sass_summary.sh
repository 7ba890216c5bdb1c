#!/bin/bash
# SASS evidence of the Blackwell-native paths in the built library: tcgen05 MMA / TMEM loads / TMA loads and stores /
# bulk copies per kernel.  usage: tools/sass_summary.sh > profiles/sass_summary_<tag>.txt
cd "$(dirname "$0")/.."
LIB=monocular-depth-estimation-cil_b200/libdepth_b200.so
echo "cuobjdump -sass $LIB  (build $(cat monocular-depth-estimation-cil_b200/build_id.txt), $(date -u +%F))"
echo "whole library:"
cuobjdump -sass $LIB | grep -oE "UTCHMMA|UTMALDG|UTMASTG|UTMAPF|LDTM|STTM|UBLKCP|UBLKPF|UTCBAR|HMMA|HGMMA|SYNCS\.[A-Z.]*|ELECT|FFMA2|F2FP\.[A-Z0-9.]*RELU[A-Z0-9._]*" | sort | uniq -c | sort -rn
echo
echo "per kernel (kernels that contain tcgen05 / TMA / bulk-copy instructions):"
cuobjdump -sass $LIB | awk '
/Function :/ {name=$3}
/UTCHMMA/ {mma[name]++} /LDTM/ {ldtm[name]++} /UTMALDG/ {ldg[name]++} /UTMASTG/ {stg[name]++} /UBLKCP/ {blk[name]++} /SETMAXREG|USETMAXREG/ {smr[name]++}
END {for (n in mma) seen[n]=1; for (n in ldg) seen[n]=1; for (n in blk) seen[n]=1;
     for (n in seen) printf "%-110s UTCHMMA %3d  LDTM %3d  UTMALDG %3d  UTMASTG %3d  UBLKCP %3d  SETMAXREG %2d\n", substr(n,1,110), mma[n], ldtm[n], ldg[n], stg[n], blk[n], smr[n]}' | sort
