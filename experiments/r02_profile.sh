#!/bin/bash
# Round-2 profiling pass (run under gpurun): launch list with DRAM bytes + full ncu captures of the kernels VERDICT r01 names.
# usage: bash experiments/r02_profile.sh <tag>
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --batch 8"
$CMD > $OUT/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -20 $OUT/${TAG}_plain.log; exit 1; }
cap() {  # name regex skip count
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 \
      -f -o $OUT/${TAG}_$1 $CMD > $OUT/${TAG}_$1.log 2>&1
  echo "ncu $1 rc=$?"
}
WHAT=${2:-conv64 convt head smallc}
for w in $WHAT; do
  case $w in
    conv64) cap conv64 '.*umma_conv_kernel<\(int\)4, \(int\)64, \(int\)2>.*' 15 5 ;;
    convt)  cap convt  '.*umma_conv_kernel<\(int\)2, \(int\)128, \(int\)1>.*' 9 3 ;;
    head)   cap head   '.*head_kernel.*' 6 2 ;;
    smallc) cap smallc '.*smallc_fwd_kernel.*' 3 1 ;;
    poolbwd) cap poolbwd '.*maxpool_bwd_kernel.*' 12 2 ;;
  esac
done
