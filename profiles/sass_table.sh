#!/bin/bash
# usage: profiles/sass_table.sh <tag>  -> profiles/sass_<tag>.txt: which Blackwell / tensor instructions each object really contains
cd "$(dirname "$0")/../featuredetection_b200/csrc/_build"
out=../../../profiles/sass_$1.txt
{
echo "# cuobjdump -sass <object> | grep -c <mnemonic>  (nvcc 12.9 -gencode arch=compute_100a,code=sm_100a; commit $(git rev-parse --short HEAD))"
echo "# UTCIMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (TMA), UBLKCP = cp.async.bulk, SYNCS = mbarrier, IMMA = mma.sync int8, DMMA = mma.sync f64, LDSM = ldmatrix"
printf "%-22s %8s %6s %8s %7s %6s %6s %6s %6s %6s %7s\n" object UTCIMMA LDTM UTMALDG UBLKCP IMMA DMMA LDSM ATOMS SYNCS IDP.4A
for o in *.cu.o; do s=$(cuobjdump -sass $o 2>/dev/null); printf "%-22s" $o; for m in UTCIMMA LDTM UTMALDG UBLKCP "IMMA\." DMMA LDSM ATOMS SYNCS "IDP.4A"; do printf " %7s" $(echo "$s" | grep -c "$m"); done; echo; done
} > $out
cat $out
