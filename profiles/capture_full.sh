#!/bin/bash
# ncu --set full captures of the three dominant kernels (run on the GPU box through gpurun; one GPU).
# Outputs gpurun_out/<tag>_{score,pileup_uncapped,hamming}.ncu-rep + raw CSV pages; profiles/extract_traffic.py turns the
# CSVs into profiles/traffic.json (dram bytes per launch), which bench.py reports as roofline.traffic.
TAG=${1:-r1h}
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:score_kernel -s 4 -c 1 -f -o gpurun_out/${TAG}_score python bench.py --steps 1 --warmup 3 --no-extras --no-graph > gpurun_out/${TAG}_ncu_score.log 2>&1
$NCU -k regex:pileup_bitsliced -s 4 -c 1 -f -o gpurun_out/${TAG}_pileup_uncapped python bench.py --steps 1 --warmup 3 --no-extras --no-graph --max-depth 0 > gpurun_out/${TAG}_ncu_pileup.log 2>&1
$NCU -k regex:hamming_min -c 9 -f -o gpurun_out/${TAG}_hamming python bench.py --only-hamming > gpurun_out/${TAG}_ncu_hamming.log 2>&1
for k in score pileup_uncapped hamming; do
  ncu -i gpurun_out/${TAG}_$k.ncu-rep --page raw --csv > gpurun_out/${TAG}_${k}_ncu_raw.csv 2>/dev/null
done
ls -la gpurun_out
