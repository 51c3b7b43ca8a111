#!/bin/bash
# ncu --set full capture of the score kernel of the bench workload (one launch, after warm-up)
TAG=${1:-r1k}
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:score_runs_kernel -s 4 -c 1 -f -o gpurun_out/${TAG}_score_runs python bench.py --steps 1 --warmup 3 --no-extras --no-graph > gpurun_out/${TAG}_ncu_score_runs.log 2>&1
ncu -i gpurun_out/${TAG}_score_runs.ncu-rep --page raw --csv > gpurun_out/${TAG}_score_runs_ncu_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_score_runs.ncu-rep --page details 2>/dev/null | grep -E "Duration|DRAM Throughput|Memory Throughput|L2 Hit|Achieved Occupancy|Theoretical Occ|Registers|Mem Busy|Max Bandwidth|Stall|Warp Cycles|Issue" | head -40
