#!/bin/bash
# ncu evidence per size bucket (north_star): DRAM bytes / throughput, shared-memory wavefront
# utilisation, FP64 pipe utilisation, warp occupancy, issue utilisation -- for every kernel of one
# run on a config-4 cohort slab (4 samples) and one on config 3 (scaled: 2 giant loci).
set -e
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,launch__registers_per_thread,launch__shared_mem_per_block_dynamic
python tools/quick_gpu.py 1.0 4 > gpurun_out/buckets_plain1.log 2>&1 &&
ncu --metrics $M --clock-control none -s 12 -c 12 --csv --log-file gpurun_out/buckets_config4.csv python tools/quick_gpu.py 1.0 4 > gpurun_out/buckets_ncu1.log 2>&1
python tools/quick_gpu.py 0.04 1 config3 300 > gpurun_out/buckets_plain2.log 2>&1 &&
ncu --metrics $M --clock-control none -s 12 -c 12 --csv --log-file gpurun_out/buckets_config3.csv python tools/quick_gpu.py 0.04 1 config3 300 > gpurun_out/buckets_ncu2.log 2>&1
