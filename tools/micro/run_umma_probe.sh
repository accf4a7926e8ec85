#!/bin/bash
# Runs every umma_probe variant in its own process (a faulting variant must not take the others with it).
out=gpurun_out/umma_probe.txt
mkdir -p gpurun_out
: > $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv >> $out 2>&1
for a_mn in 0 1; do for opt in 0 1; do for swap in 0 1; do
  timeout 60 tools/micro/umma_probe layout $a_mn $opt $swap >> $out 2>&1 || echo "layout $a_mn $opt $swap: exit $?" >> $out
done; done; done
timeout 60 tools/micro/umma_probe split 0 >> $out 2>&1 || echo "split 0: exit $?" >> $out
timeout 60 tools/micro/umma_probe split 1 >> $out 2>&1 || echo "split 1: exit $?" >> $out
timeout 120 tools/micro/umma_probe time >> $out 2>&1 || echo "time: exit $?" >> $out
cat $out
