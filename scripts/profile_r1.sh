#!/bin/bash
# Round-1 profiling recipe (run under gpurun): launch list + one --set full capture of each hot-path kernel,
# on a reduced copy of the bench workload (200k reads vs 300 Mbp) so that ncu's ~40 replays per launch stay short.
set -x
mkdir -p gpurun_out
W="--reads 200000 --ref-bp 300000000 --contigs 32 --sketch 220 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_sketch|k_l1_|k_l2_|k_publish|k_zero_words|k_set_u32|DeviceScan|DeviceRadixSort' -c 400 --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --steps 1 --warmup 1 $W > gpurun_out/launches_r1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_sketch|k_l1_probe|k_l1_warp|k_l1_cta|k_l2_ranges|k_l2_prep|k_l2_scan' -s 7 -c 7 \
    -o gpurun_out/prof_r1 python bench.py --steps 1 --warmup 1 $W > gpurun_out/prof_r1.log 2>&1
ls -la gpurun_out
