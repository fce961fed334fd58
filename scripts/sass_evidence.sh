#!/usr/bin/env bash
# Per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths -> profiles/sass_mnemonics.txt
cd "$(dirname "$0")/.."
out=profiles/sass_mnemonics.txt
{
  echo "# SASS mnemonics per kernel (cuobjdump -sass mdi_llm_b200/ops/_mdi_ops.so) — proof of the Blackwell-native paths"
  echo "# UTCHMMA = tcgen05.mma kind::f16, UTCQMMA = tcgen05.mma kind::f8f6f4 (fp8), LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (TMA), UBLKCP = cp.async.bulk, UTMAPF/UBLKPF = bulk prefetch,"
  echo "# UTCBAR = tcgen05.commit, UTCATOMSWS = TMEM alloc/dealloc, SYNCS = mbarrier, MEMBAR.SC.SYS + ATOMG = hop ticket/flag release"
  cuobjdump -sass mdi_llm_b200/ops/_mdi_ops.so | awk '
    /Function :/ { fn=$3 }
    { for (i=1;i<=NF;i++) if ($i ~ /^(UTCHMMA|UTCQMMA|FFMA2|FMUL2|LDTM|UTMALDG|UBLKCP|UBLKPF|UTMAPF|UTCBAR|UTCATOMSWS|SYNCS|MEMBAR\.SC\.SYS|ATOMG|LDG\.E\.EF\.128|ERRBAR|ACQBULK|UTMACCTL)/) { gsub(/;$/,"",$i); c[fn" "$i]++ } }
    END { for (k in c) print c[k], k }' | sort -k2,2 -k1,1nr
} > $out
wc -l $out
