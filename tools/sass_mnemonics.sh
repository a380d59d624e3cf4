#!/bin/bash
# Per-kernel counts of the SASS mnemonics that prove (or disprove) a Blackwell-native kernel
# (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG, mma.sync -> HMMA).
SO=${1:-oobleck_b200/liboobleck_b200.so}
echo "# cuobjdump -sass $SO | per-kernel mnemonic counts   ($(date -u +%F))"
cuobjdump -sass "$SO" 2>/dev/null | awk '
/Function :/ {fn=$3}
{for(i=1;i<=NF;i++) if ($i ~ /^(UTCHMMA|UTCQMMA|UTCBAR|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|HMMA|LDGSTS|SYNCS|MUFU\.EX2)/) {k=$i; sub(/\..*/,"",k); c[fn" "k]++; fns[fn]=1}}
END {for (f in fns) {printf "%s :", f; n=split("UTCHMMA LDTM STTM UTMALDG UTMASTG HMMA LDGSTS MUFU",ks," "); for(j=1;j<=n;j++) if (c[f" "ks[j]]>0) printf " %s=%d", ks[j], c[f" "ks[j]]; printf "\n"}}' | c++filt | sed 's/(anonymous namespace):://; s/(CUtensorMap_st.*) :/(...) :/; s/(float.*) :/(...) :/; s/(long.*) :/(...) :/; s/(unsigned.*) :/(...) :/; s/(char.*) :/(...) :/; s/(int.*) :/(...) :/' | sort
