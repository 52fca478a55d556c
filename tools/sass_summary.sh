#!/bin/bash
# Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA use (B200_PROFILING.md):
#   UTCHMMA/UTCQMMA (tcgen05.mma), UTCBAR (tcgen05.commit), LDTM/STTM (tcgen05.ld/st), UTMALDG/UTMASTG (TMA tensor),
#   UBLKCP (cp.async.bulk), SYNCS (mbarrier), STAS (st.async); a .2CTA suffix marks the cta_group::2 forms
# Usage: tools/sass_summary.sh > profiles/r02_sass_summary.txt
cd "$(dirname "$0")/.."
for lib in kaldi_ctc_b200/libb200rnn.so kaldi_ctc_b200/libb200ctc.so; do
  echo "== $lib ($(stat -c %s $lib) bytes, $(date -u -r $lib +%FT%TZ))"
  cuobjdump -sass $lib | awk '
    /Function :/ { fn=$3 }
    { for (i=1;i<=NF;i++) { t=$i; two=(t ~ /2CTA/); sub(/\..*/,"",t);
        if (t ~ /^(UTCHMMA|UTCQMMA|UTCBAR|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|UBLKRED|SYNCS|STAS|UTCCP)$/) c[fn" "t (two ? ".2CTA" : "")]++ } }
    END { for (k in c) print k, c[k] }' | sort | c++filt | awk '{n=$NF; m=$(NF-1); $NF=""; $(NF-1)=""; printf "%-8s %5d  %s\n", m, n, $0}' | sed 's/(b200::RecArgs.*//; s/(CUtensorMap.*//; s/((anonymous namespace)::CtcDev.*//' 
done
