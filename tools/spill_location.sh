#!/bin/bash
# Where do the register spills ptxas reports (csrc/*.ptxas.log) sit?  For every kernel of libb200rnn.so that has
# local-memory instructions (STL / LDL), prints how many of them are inside a per-time-step loop -- i.e. between the
# target and the source of a backward branch whose range also holds a tcgen05.mma (UTCHMMA) or a tcgen05.ld (LDTM) --
# and how many are in code that runs once per launch (the prologue that converts the recurrent weights to BF16 and
# stores them to tensor memory with its own small loop, the set-up of the epilogue warps, the bias sums after the
# last time step).  Loop back-edges are the conditional backward branches (the unconditional ones return from the
# BRA.DIV handlers at the end of a function).
# Usage: tools/spill_location.sh > profiles/r02_spill_location.txt
cd "$(dirname "$0")/.."
cuobjdump -sass kaldi_ctc_b200/libb200rnn.so | awk '
  function hex(s,   i, c, v, d) { v = 0; s = tolower(s); for (i = 1; i <= length(s); i++) { c = substr(s, i, 1); d = index("0123456789abcdef", c) - 1; if (d < 0) break; v = v * 16 + d }; return v }
  function flush(   i, j, inl, once) {
    if (fn == "" || nl == 0) return
    for (j = 1; j <= nb; j++) { steploop[j] = 0; for (i = 1; i <= nm; i++) if (bt[j] <= mm[i] && mm[i] <= bs[j]) { steploop[j] = 1; break } }
    inl = 0; once = 0
    for (i = 1; i <= nl; i++) {
      hit = 0
      for (j = 1; j <= nb; j++) if (bt[j] <= loc[i] && loc[i] <= bs[j] && steploop[j]) { hit = 1; break }
      if (hit) inl++; else once++
    }
    printf "run once per launch %3d   inside a per-time-step loop %3d   %s\n", once, inl, fn
  }
  /Function :/ { flush(); fn=$3; nl=0; nb=0; nm=0 }
  match($0, /\/\*[0-9a-f]+\*\//) {
    addr = hex(substr($0, RSTART + 2, RLENGTH - 4))
    if ($0 ~ / STL| LDL/) loc[++nl] = addr
    if ($0 ~ /UTCHMMA|UTCQMMA|LDTM/) mm[++nm] = addr
    if ($0 ~ /@!?U?P[0-9] +BRA/ && match($0, /0x[0-9a-f]+ *;/)) { t = hex(substr($0, RSTART + 2)); if (t <= addr) { nb++; bt[nb] = t; bs[nb] = addr } }
  }
  END { flush() }' | c++filt | sed 's/(b200::RecArgs.*//; s/b200::(anonymous namespace):://'
