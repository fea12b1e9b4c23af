#!/bin/bash
# launch list only (per-launch gpu__time_duration of this repo's kernels + the two library primitives between them)
set -x
mkdir -p gpurun_out
W="--reads 200000 --ref-bp 300000000 --contigs 32 --sketch 220 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_sketch|k_l1_|k_l2_|k_publish|k_zero_words|k_set_u32|DeviceScan|DeviceRadixSort' -c 400 --csv \
    --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 1 $W > gpurun_out/launches_r1.log 2>&1
tail -2 gpurun_out/launches_r1.log | cut -c1-600
