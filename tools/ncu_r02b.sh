#!/bin/bash
# Round-2 (second half) profiler passes of the final build: fused level-0/1 kernel, packed decimation.  Run on the GPU box through
# gpurun; outputs under gpurun_out/.  Numbers printed by runs under ncu are never bench values.
set -u
OUT=gpurun_out
M="gpu__time_duration.sum"
# 1. launch list of the headline step (windowed pairs)
ncu --metrics $M --clock-control none --csv --log-file $OUT/launches_r02b_pairs.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --api-pairs 0 --seqs 0 --sustain-s 0.05 > $OUT/ncu_pairs_b.log 2>&1
# 2. full counter set of one pair step (5 kernels: 2 x stream_level01, 2 x stream_down2p, lk_windowed)
ncu --set full --clock-control none --import-source on -k regex:'stream_|lk_' -s 30 -c 5 \
    -o $OUT/prof_pairs_r02b python bench.py --steps 1 --warmup 3 --no-cpu-baseline --api-pairs 0 --seqs 0 --sustain-s 0.01 > $OUT/ncu_pairs_full_b.log 2>&1
# 3. launch list of sequence steps (plain launches so that every kernel shows by name)
KLT_B200_NO_GRAPH=1 ncu --metrics $M --clock-control none --csv --log-file $OUT/launches_r02b_sequence.csv \
    python tools/seq_probe.py --batches 8 --frames 3 --modes fast > $OUT/ncu_seq_b.log 2>&1
ls -la $OUT/*.ncu-rep
