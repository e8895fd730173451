#!/bin/bash
# Round-2 profiling pass #2 (run under gpurun): cfg-3 launch list with DRAM bytes at the full batch, then ncu --set full
# captures of EVERY tcgen05 conv launch of one step (so the ConvTranspose / 64-channel launches can be picked by shape),
# of the wgrad launches and of the HBM-side kernels.  usage: bash experiments/r02_profile2.sh <tag>
set -u
TAG=${1:-r02h}
OUT=gpurun_out
mkdir -p $OUT
BASE="python bench.py --steps 1 --warmup 1 --no-cpu --no-cudnn"
$BASE > $OUT/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -20 $OUT/${TAG}_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv \
    --log-file $OUT/${TAG}_launches_cfg3.csv $BASE > $OUT/${TAG}_launches.log 2>&1
echo "launch list rc=$?"
CMD="$BASE --batch 8"
cap() {  # name regex skip count
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 \
      -f -o $OUT/${TAG}_$1 $CMD > $OUT/${TAG}_$1.log 2>&1
  echo "ncu $1 rc=$?"
}
cap conv  '.*umma_conv_kernel.*' 0 90
cap wgrad '.*wgrad_umma_kernel.*' 0 44
cap hbm   '.*(head_dense|head_kernel|smallc|maxpool|im2col).*' 0 16
ls -la $OUT/${TAG}_*
