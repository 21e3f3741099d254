#!/bin/bash
# Per-kernel census of the Blackwell-only SASS mnemonics in libirr_b200.so (VERDICT r1 'next' #10): proves the conv
# kernels are tcgen05/TMEM/TMA code and the correlation kernels are TMA-fed SIMT code.
#   scripts/sass_census.sh > profiles/r02_sass_census.txt
SO=${1:-irr_b200/libirr_b200.so}
echo "# SASS census of $SO ($(date -u +%Y-%m-%dT%H:%MZ)); cuobjdump -sass, sm_100a"
echo "# columns: UTCHMMA (tcgen05.mma) | UTCBAR (tcgen05.commit) | STTM (tcgen05.st) | LDTM (tcgen05.ld) | UTMALDG (TMA tensor load) | UBLKCP (bulk copy) | SYNCS (mbarrier) | FFMA | HMMA/IMMA (legacy mma.sync) | kernel"
cuobjdump -sass "$SO" | awk '
  /Function :/ { if (name != "") emit(); name=$3; for (k in c) delete c[k]; next }
  { if ($0 ~ /UTCHMMA/) c["a"]++; if ($0 ~ /UTCBAR/) c["b"]++; if ($0 ~ /STTM/) c["c"]++; if ($0 ~ /LDTM/) c["d"]++;
    if ($0 ~ /UTMALDG/) c["e"]++; if ($0 ~ /UBLKCP/) c["f"]++; if ($0 ~ /SYNCS/) c["g"]++; if ($0 ~ / FFMA/) c["h"]++;
    if ($0 ~ /[^C]HMMA\.|[ \t]IMMA\./) c["i"]++ }
  function emit() { printf "%6d %6d %6d %6d %6d %6d %6d %6d %6d  %s\n", c["a"],c["b"],c["c"],c["d"],c["e"],c["f"],c["g"],c["h"],c["i"], name }
  END { if (name != "") emit() }' | while read -r a b c d e f g h i name; do
    printf "%6s %6s %6s %6s %6s %6s %6s %6s %6s  %s\n" "$a" "$b" "$c" "$d" "$e" "$f" "$g" "$h" "$i" "$(echo "$name" | c++filt | cut -c1-110)"
  done | sort -k10
