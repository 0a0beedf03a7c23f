#!/bin/bash
# Round-2 profiler passes (run on the GPU box through gpurun; outputs under gpurun_out/, <= 64 MiB in total).  Numbers printed
# by runs under ncu are never bench values: only the launch lists and counter captures are kept.
set -u
OUT=gpurun_out
M="gpu__time_duration.sum"
# 1. launch list of the headline step (windowed pairs): the run's launches after warm-up
ncu --metrics $M --clock-control none --csv --log-file $OUT/launches_r02_pairs.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --api-pairs 0 --seqs 0 --sustain-s 0.05 > $OUT/ncu_pairs.log 2>&1
# 2. launch list of sequence steps (config D shape, 8 sequences, windowed + fast selection), plain launches so that every kernel shows by name
KLT_B200_NO_GRAPH=1 ncu --metrics $M --clock-control none --csv --log-file $OUT/launches_r02_sequence.csv \
    python tools/seq_probe.py --batches 8 --frames 3 --modes fast > $OUT/ncu_seq.log 2>&1
# 3. full counter sets: one steady-state sequence step (10 kernels) and one pair step (7 kernels)
KLT_B200_NO_GRAPH=1 ncu --set full --clock-control none -k regex:'eigen_fast|select_|premark|lk_windowed|stream_' -s 70 -c 10 \
    -o $OUT/prof_sequence_r02 python tools/seq_probe.py --batches 8 --frames 4 --modes fast > $OUT/ncu_seq_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'stream_|lk_' -s 42 -c 7 \
    -o $OUT/prof_pairs_r02 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --api-pairs 0 --seqs 0 --sustain-s 0.01 > $OUT/ncu_pairs_full.log 2>&1
ls -la $OUT/*.ncu-rep
