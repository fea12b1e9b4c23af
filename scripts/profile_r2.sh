#!/bin/bash
# Round-2 profiling recipe (run under gpurun): launch list + one --set full capture of each hot-path kernel, on a
# reduced copy of the bench workload (200k reads vs 300 Mbp) so that ncu's ~40 replays per launch stay short; then the
# index builder's window-scan kernel on the same reference. The reports are summarised on the box (raw metrics as CSV,
# hot source lines as text); only the mapping report itself is kept (gpurun_out/ is limited to 64 MiB).
set -x
mkdir -p gpurun_out
W="--config 2 --reads 200000 --ref-bp 300000000 --contigs 32 --sketch 220 --no-cpu-baseline"
K='k_pack_bases|k_sketch|k_l1_|k_l2_|k_publish|k_zero_words|k_set_u32|DeviceScan|DeviceRadixSort|k_window|k_resolve|k_patch|k_gather|k_scatter|k_lookup|k_mark|k_keep'
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -c 600 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 1 --warmup 1 $W > gpurun_out/launches_r2.log 2>&1
# the mapping kernels of the second (timed) step: 9 launches per step (pack, sketch, sketch_table, probe, warp, cta, ranges, prep, scan)
ncu --set full --clock-control none --import-source on -k regex:'k_pack_bases|k_sketch|k_l1_probe|k_l1_warp|k_l1_cta|k_l2_ranges|k_l2_prep|k_l2_scan' -s 9 -c 9 \
    -o gpurun_out/prof_r2 python bench.py --steps 1 --warmup 1 $W > gpurun_out/prof_r2.log 2>&1
ncu -i gpurun_out/prof_r2.ncu-rep --page raw --csv > gpurun_out/prof_r2_raw.csv 2>/dev/null
for k in k_sketch k_l1_warp k_l1_probe k_l2_prep k_l2_scan; do
  python scripts/ncu_lines.py gpurun_out/prof_r2.ncu-rep "$k" 40 > gpurun_out/lines_r2_$k.txt 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:'k_window_scan' -c 1 \
    -o gpurun_out/prof_r2_index python bench.py --steps 1 --warmup 1 $W > gpurun_out/prof_r2_index.log 2>&1
ncu -i gpurun_out/prof_r2_index.ncu-rep --page raw --csv > gpurun_out/prof_r2_index_raw.csv 2>/dev/null
python scripts/ncu_lines.py gpurun_out/prof_r2_index.ncu-rep k_window_scan 50 > gpurun_out/lines_r2_k_window_scan.txt 2>&1
rm -f gpurun_out/prof_r2_index.ncu-rep
ls -la gpurun_out
