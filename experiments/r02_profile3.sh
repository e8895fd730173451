#!/bin/bash
# Round-2 profiling pass (run under gpurun).  The .ncu-rep files are summarised ON THE BOX (raw metrics + per-instruction stall
# reports as text) and deleted, because gpurun only brings back 64 MiB.   usage: bash experiments/r02_profile3.sh <tag>
set -u
TAG=${1:-r02i}
OUT=gpurun_out
mkdir -p $OUT
BASE="python bench.py --steps 1 --warmup 1 --no-cpu --no-cudnn"
$BASE > $OUT/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -20 $OUT/${TAG}_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv \
    --log-file $OUT/${TAG}_launches_cfg3.csv $BASE > $OUT/${TAG}_launches.log 2>&1
echo "launch list rc=$?"
gzip -f $OUT/${TAG}_launches_cfg3.csv
cap() {  # name regex count "launches to stall-report"
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -c $3 \
      -f -o /tmp/${TAG}_$1 $BASE > $OUT/${TAG}_$1.log 2>&1
  echo "ncu $1 rc=$?"
  python experiments/ncu_summary.py /tmp/${TAG}_$1.ncu-rep > $OUT/${TAG}_$1_metrics.txt 2>&1
  ncu -i /tmp/${TAG}_$1.ncu-rep --page raw --csv 2>/dev/null | gzip > $OUT/${TAG}_$1_raw.csv.gz
  for l in $4; do
    python experiments/stall_report2.py /tmp/${TAG}_$1.ncu-rep $l 60 > $OUT/${TAG}_$1_stalls_$l.txt 2>&1
  done
  rm -f /tmp/${TAG}_$1.ncu-rep
}
cap conv  '.*umma_conv_kernel.*' 26 "0 1 9 18 20 21 22 23"
cap wgrad '.*wgrad_umma_kernel.*' 4 "0 1 2"
cap hbm   ".*(head_pix|smallc_mma|im2col3x3_c1|maxpool_bwd|maxpool_fwd).*" 8 "0 1 2 3 4"
du -sh $OUT
