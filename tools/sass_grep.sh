#!/bin/bash
# Counts of Blackwell-native SASS mnemonics in libmnf_b200.so (B200_PROFILING.md "What proves a Blackwell-native kernel").
# usage: tools/sass_grep.sh > profiles/r02_sass_grep.txt
LIB="$(dirname "$0")/../torch-mnf_b200/lib/libmnf_b200.so"
SASS=$(mktemp)
cuobjdump -sass "$LIB" > "$SASS"
echo "# cuobjdump -sass torch-mnf_b200/lib/libmnf_b200.so | grep -c <mnemonic>   ($(date -u +%Y-%m-%dT%H:%MZ), $(nvcc --version | tail -1))"
for m in UTCHMMA UTCQMMA UTCMMA UTMALDG UTMASTG UBLKCP LDTM STTM UTCBAR UTCCP "SYNCS" FFMA2 "LDCU" HMMA "MUFU.EX2" "MUFU.LG2" "MUFU.RCP" LDGSTS "MULTIMEM\|multimem" "REDUX"; do
  printf "%-18s %s\n" "$m" "$(grep -c "$m" "$SASS")"
done
echo
echo "# per kernel: tensor-core / TMA / TMEM instructions"
awk '/Function :/ {fn=$3} /UTCHMMA|UTMALDG|UTMASTG|LDTM|STTM|UTCBAR/ {split($0,a," "); for(i in a) if (a[i] ~ /^(UTCHMMA|UTMALDG|UTMASTG|LDTM|STTM|UTCBAR)/) {gsub(/\..*/,"",a[i]); c[fn" "a[i]]++}} END {for (k in c) print c[k], k}' "$SASS" | sort -k2 | c++filt | cut -c1-200
rm -f "$SASS"
