#!/bin/bash
# SASS mnemonic counts of the built objects (cuobjdump -sass): the proof that the tensor-core / TMA /
# bulk-copy instructions the design names are in the binary. usage: scripts/sass_counts.sh > profiles/rNN_sass_counts.txt
cd "$(dirname "$0")/../semadb_b200/lib" || exit 1
echo "cuobjdump -sass semadb_b200/lib/*.o — instruction counts per object ($(/usr/local/cuda/bin/nvcc --version | tail -1))"
for o in flat_tc.o search.o search_l2.o search_bits.o search_pq.o search_pq_fly8.o insert.o quant.o dist.o; do
  [ -f "$o" ] || continue
  echo "== $o"
  cuobjdump -sass "$o" > /tmp/sass_$$.txt
  for m in UTCHMMA UTCHMMA.2CTA UTCBAR LDTM STTM UTMALDG UTMASTG UBLKCP UBLKPF SYNCS LDG.E.ENL2.256 "LDG.E.*128" "LDG.E.NA.128.CONSTANT" "CCTL" "PREFETCH\|CCTL.E.PF" LDS.128 LDS SHFL POPC FFMA HMMA ATOMS "ATOMG\|RED.E" WARPSYNC.COLLECTIVE REDUX BAR.SYNC; do
    c=$(grep -c -E "\b$m" /tmp/sass_$$.txt)
    [ "$c" != "0" ] && printf "  %-28s %8d\n" "$m" "$c"
  done
done
rm -f /tmp/sass_$$.txt
