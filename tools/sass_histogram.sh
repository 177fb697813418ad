#!/usr/bin/env bash
# Per-kernel histogram of the SASS mnemonics that prove (or disprove) a Blackwell-native path
# (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UTMAREDG; legacy HMMA =
# mma.sync). Usage: tools/sass_histogram.sh [lib.so] > profiles/r2_sass_histogram.txt
LIB="${1:-clip_fsar_b200/libfsar_sm100.so}"
echo "# cuobjdump -sass $LIB  ($(date -u +%F))  -- occurrences per kernel"
cuobjdump -sass "$LIB" | awk '
/Function :/ { f=$3; fn[f]=1; next }
{
  if ($0 ~ /UTCHMMA\.2CTA/) c[f,"UTCHMMA.2CTA"]++; else if ($0 ~ /UTCHMMA/) c[f,"UTCHMMA"]++;
  if ($0 ~ /UTMALDG/) c[f,"UTMALDG"]++;
  if ($0 ~ /UTMASTG/) c[f,"UTMASTG"]++;
  if ($0 ~ /UTMAREDG/) c[f,"UTMAREDG"]++;
  if ($0 ~ / LDTM/) c[f,"LDTM"]++;
  if ($0 ~ / STTM/) c[f,"STTM"]++;
  if ($0 ~ /UTCBAR/) c[f,"UTCBAR"]++;
  if ($0 ~ / HMMA\./) c[f,"HMMA(legacy)"]++;
  if ($0 ~ /MUFU\.EX2/) c[f,"MUFU.EX2"]++;
  if ($0 ~ /MUFU\.TANH/) c[f,"MUFU.TANH"]++;
}
END {
  n=split("UTCHMMA.2CTA UTCHMMA UTMALDG UTMASTG UTMAREDG LDTM STTM UTCBAR MUFU.EX2 MUFU.TANH HMMA(legacy)", cols, " ");
  printf "0 %-72s", "kernel"; for (i=1;i<=n;i++) printf " %12s", cols[i]; printf "\n";
  for (f in fn) { printf "1 %-72s", substr(f,1,72); for (i=1;i<=n;i++) printf " %12d", c[f,cols[i]]+0; printf "\n"; t++ }
  printf "2 # %d kernels; total legacy HMMA (mma.sync): ", t; s=0; for (f in fn) s+=c[f,"HMMA(legacy)"]; print s
}' | sort | cut -c3-
