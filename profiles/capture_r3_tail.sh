#!/bin/bash
# Round-2 (third session): full-set ncu capture of the three tail kernels of one serial pass (selection, depth-capped pileup, consensus), with source
T=${1:-r3a}
ncu --set full --clock-control none --import-source on -k regex:"sel_locus|pileup_bitsliced|consensus_kernel" -s 9 -c 3 -f -o gpurun_out/${T}_tail \
  python bench.py --steps 1 --warmup 3 --no-extras --no-parity-check --ingest-reads 0 --no-graph --lanes 1 > gpurun_out/${T}_ncu_tail.log 2>&1
ncu -i gpurun_out/${T}_tail.ncu-rep --page details > gpurun_out/${T}_tail_details.txt 2>/dev/null
ncu -i gpurun_out/${T}_tail.ncu-rep --page source --csv > gpurun_out/${T}_tail_source.csv 2>/dev/null
tail -5 gpurun_out/${T}_ncu_tail.log; ls -la gpurun_out | tail
