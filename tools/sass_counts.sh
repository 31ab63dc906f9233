#!/usr/bin/env bash
# SASS opcode histogram of every kernel in libfa2_b200.so (runs on a CPU box: cuobjdump only reads the cubin).
# The Blackwell-native evidence: UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UTMAREDG = TMA
# tensor load / store / reduce, UTCBAR = tcgen05.commit; HMMA (legacy mma.sync) must be absent.
set -euo pipefail
LIB=${1:-$(dirname "$0")/../cuda-flash-attention_b200/libfa2_b200.so}
cuobjdump -sass "$LIB" | awk '
  /Function :/ { fn=$3; next }
  /^[[:space:]]+\/\*[0-9a-f]{4}\*\// {
     op=$2; if (op ~ /^@/) op=$3; sub(/;$/, "", op);
     n=split(op, parts, "."); base=parts[1];
     key=base; if (base ~ /^(UTC|UTMA|LDTM|STTM|UBLKCP|SYNCS|MUFU|HMMA|FFMA2|FADD2|FMUL2|ELECT|USETMAXREG|REDG|RED)/) key=op;
     cnt[fn SUBSEP key]++; tot[fn]++ }
  END {
     for (k in cnt) { split(k, a, SUBSEP); print a[1] "\t" cnt[k] "\t" a[2] }
  }' | sort -k1,1 -k2,2nr | awk -F'\t' '
  { if ($1 != last) { if (last != "") print ""; cmd="c++filt " $1; cmd | getline name; close(cmd); print "== " name; last=$1; shown=0 }
    if ($3 ~ /^(UTC|UTMA|LDTM|STTM|UBLKCP|SYNCS|MUFU|HMMA|FFMA2|FADD2|FMUL2|ELECT|USETMAXREG|REDG|RED)/ || shown < 12) { printf "  %6d  %s\n", $2, $3; shown++ } }'
