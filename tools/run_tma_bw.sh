#!/bin/bash
# TMA delivery-rate experiments (tools/umma_probe.cu, tma_bw): mode rows row_bytes pitch depth swizzle waitmode
cd "$(dirname "$0")/.."
P=tools/umma_probe.bin
LOG=gpurun_out/tma_bw.log
: > $LOG
run() { timeout 60 $P tma_bw "$@" >> $LOG 2>&1; }
run 1 256 128 128 4 0 0
run 2 256 128 128 4 0 0
run 4 256 128 128 4 0 0
run 8 256 128 128 4 0 0
run 1 256 128 128 1 0 0
run 1 16 128 128 4 0 0
run 1 16 128 128 8 0 0
run 1 1024 128 128 1 0 0
run 4 1024 128 128 1 0 0
cat $LOG
