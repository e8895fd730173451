#!/bin/bash
# SASS evidence that the hot path is tcgen05 / TMEM / TMA code: mnemonic counts of the built library.
# usage: bash experiments/sass_summary.sh > profiles/r02/r02_sass_summary.txt
LIB=pytorch-unet_b200/b200unet/libb200unet.so
T=$(mktemp)
cuobjdump -sass $LIB > $T
echo "library: $LIB ($(stat -c %s $LIB) bytes), $(cuobjdump -lelf $LIB | grep -c sm_100a) sm_100a ELF image(s)"
count() { printf "%-34s %6d   %s\n" "$1" "$(grep -cE "$2" $T)" "$3"; }
count "UTCHMMA (all)"            "UTCHMMA"                 "tcgen05.mma"
count "UTCHMMA.2CTA"             "UTCHMMA\.2CTA"           "tcgen05.mma.cta_group::2 (CTA pairs)"
count "UTCHMMA ... A_KEEP/A_REUSE" "UTCHMMA.*A_(KEEP|REUSE)" "tcgen05.mma .collector::a::fill/use/lastuse"
count "UTMALDG"                  "UTMALDG"                 "cp.async.bulk.tensor (TMA loads)"
count "UTMALDG ... MULTICAST/2CTA" "UTMALDG.*(MULTICAST|2CTA)" "TMA loads completing on the pair leader's barrier"
count "LDTM"                     "LDTM"                    "tcgen05.ld (TMEM -> registers)"
count "STTM"                     "STTM"                    "tcgen05.st (bias pre-load into TMEM)"
count "UTCBAR"                   "UTCBAR"                  "tcgen05.commit"
count "UTCBAR.2CTA.MULTICAST"    "UTCBAR.*2CTA.*MULTICAST" "tcgen05.commit multicast to both CTAs of a pair"
count "SYNCS.*TRYWAIT"           "SYNCS.*TRYWAIT"          "mbarrier.try_wait"
count "HMMA (mma.sync, first layer)" "[^C]HMMA\.16816"     "mma.sync.m16n8k16 bf16 (first-layer forward)"
count "LDGSTS"                   "LDGSTS"                  "cp.async (head kernels' shared-memory ring)"
count "FFMA2"                    "FFMA2"                   "packed fp32x2 FMA (head / first-layer CUDA-core kernels)"
echo
echo "kernels containing UTCHMMA:"
awk '/Function :/ {f=$3} /UTCHMMA/ {c[f]++} END {for (k in c) printf "  %5d  %s\n", c[k], k}' $T | sort -k2 | c++filt | cut -c1-150
rm -f $T
