#!/bin/bash
# Runs the UMMA descriptor probe over a grid of configurations; each in its own process so a fault is isolated.
cd "$(dirname "$0")"
out=../gpurun_out/probe.log
: > $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv >> $out 2>&1
run() { timeout 30 ./umma_probe "$@" >> $out 2>&1 || echo "cfg $* : exit $?" >> $out; }
run 0 0 0 0 0
for bo in 0 1; do
  for s in 1 2 3 5 8 9 18; do run 0 0 $s 0 $bo; done
  run 0 0 0 3 $bo
  run 0 0 19 0 $bo
done
run 1 1 0 0 0
run 0 1 0 0 0
run 1 0 0 0 0
for bo in 0 1; do
  for s in 1 3 8 9 18; do run 1 1 0 $s $bo; done
  run 1 1 3 3 $bo
  run 1 1 5 0 $bo
done
cat $out
