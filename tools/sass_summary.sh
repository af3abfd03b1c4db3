#!/bin/bash
# Which Blackwell instructions each object of libccx contains (cuobjdump -sass, counted per object):
#   UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor copies, UBLKCP = cp.async.bulk,
#   HMMA = mma.sync, FFMA2 = packed fp32 FMA, SYNCS = mbarrier, UCGABAR = cluster barrier.
# usage: tools/sass_summary.sh > profiles/r02_sass_summary.txt   (after `make`)
printf "%-24s %8s %6s %6s %8s %8s %7s %7s %8s %7s %8s\n" object UTCxMMA LDTM STTM UTMALDG UTMASTG UBLKCP HMMA FFMA2 SYNCS UCGABAR
for o in build/*.o; do
  s=$(cuobjdump -sass "$o" 2>/dev/null)
  c() { echo "$s" | grep -c -E "$1"; }
  printf "%-24s %8d %6d %6d %8d %8d %7d %7d %8d %7d %8d\n" "$(basename $o)" "$(c 'UTC[A-Z]*MMA')" "$(c 'LDTM')" "$(c 'STTM')" \
    "$(c 'UTMALDG')" "$(c 'UTMASTG')" "$(c 'UBLKCP')" "$(c '[^A-Z]HMMA')" "$(c 'FFMA2')" "$(c 'SYNCS')" "$(c 'UCGABAR')"
done
echo
echo "# kernels of lstm_persist.o (the persistent LSTM recurrence) and their tensor-core instructions:"
cuobjdump -sass build/lstm_persist.o 2>/dev/null | awk '/Function :/ {name=$3} /UTC[A-Z]*MMA/ {u[name]++} /[^A-Z]HMMA/ {h[name]++} /LDTM/ {l[name]++} /UBLKCP/ {b[name]++} /UTMALDG/ {t[name]++} END {for (n in h) printf "%s  UTCxMMA=%d LDTM=%d HMMA=%d UBLKCP=%d UTMALDG=%d\n", n, u[n], l[n], h[n], b[n], t[n]}'
