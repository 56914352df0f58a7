#!/bin/bash
# profiles/sass_*.txt: per kernel of libgandtr_b200.so, the counts of the SASS mnemonics that prove the hardware path
# (tcgen05 = UTCHMMA / UTCQMMA..., TMA = UTMALDG / UTMASTG, TMEM = LDTM / STTM, cluster barriers = UTCBAR, packed fp32 =
# FMUL2 / FADD2 / FFMA2, dp4a / dp2a = IDP). usage: tools/sass_excerpt.sh > profiles/sass_r2.txt
SO=${1:-gandtr_b200/libgandtr_b200.so}
cuobjdump -sass "$SO" | awk '
/Function :/ { fn=$3 }
{
  if (match($0, /(UTC[A-Z]*MMA[A-Za-z0-9_.]*|UTMALDG[A-Za-z0-9_.]*|UTMASTG[A-Za-z0-9_.]*|UTMAPF[A-Za-z0-9_.]*|LDTM[A-Za-z0-9_.]*|STTM[A-Za-z0-9_.]*|UTCBAR[A-Za-z0-9_.]*|UTCATOMSWS[A-Za-z0-9_.]*|FMUL2|FADD2|FFMA2|IDP[A-Za-z0-9_.]*|SYNCS[A-Za-z0-9_.]*|TLD[A-Za-z0-9_.]*|ATOMS[A-Za-z0-9_.]*|REDG[A-Za-z0-9_.]*|RED\.[A-Za-z0-9_.]*)/)) {
    m=substr($0, RSTART, RLENGTH); c[fn "\t" m]++
  }
}
END { for (k in c) print k "\t" c[k] }' | sort | c++filt | awk -F'\t' '{ n=$1; sub(/\(.*/, "", n); printf "%-110s %-28s %5d\n", substr(n,1,110), $2, $3 }'
