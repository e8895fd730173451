#!/bin/bash
# ncu captures of the round-2 HBM-side kernels (first-layer forward, head, im2col); summarised on the box.
set -u
TAG=${1:-r02l}
OUT=gpurun_out
BASE="python bench.py --steps 1 --warmup 1 --no-cpu --no-cudnn"
cap() {
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -c $3 \
      -f -o /tmp/${TAG}_$1 $BASE > $OUT/${TAG}_$1.log 2>&1
  echo "ncu $1 rc=$?"
  python experiments/ncu_summary.py /tmp/${TAG}_$1.ncu-rep > $OUT/${TAG}_$1_metrics.txt 2>&1
  ncu -i /tmp/${TAG}_$1.ncu-rep --page raw --csv 2>/dev/null | gzip > $OUT/${TAG}_$1_raw.csv.gz
  for l in $4; do
    python experiments/stall_report2.py /tmp/${TAG}_$1.ncu-rep $l 40 > $OUT/${TAG}_$1_stalls_$l.txt 2>&1
  done
  rm -f /tmp/${TAG}_$1.ncu-rep
}
cap hbm '.*(head_pix|smallc_fwd|im2col3x3_c1).*' 4 "0 1 2 3"
