#!/bin/bash
# ncu evidence of the final tree of round 2 (one GPU, under gpurun).  Nothing printed by a bench run under ncu is a bench value.
T=${1:-r3o}
# 1. every launch of the timed steps of the default bench (serial eager passes so that each kernel is listed), cold-cache and serialised:
#    compare SHARES with kernel_ms_per_step of the bench line, not absolutes
MMLST_CUDA_PROFILER=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv \
  python bench.py --steps 4 --warmup 3 --no-extras --no-parity-check --ingest-reads 0 --no-graph --lanes 1 > gpurun_out/${T}_launches_bench.log 2>&1
# 2. the dominant kernel of the pass: score, form 6 (ring with pairs of chunks reduced together, stream read evict-first)
ncu --set full --clock-control none --import-source on -k regex:score_runs_ring_pair -s 6 -c 1 -f -o gpurun_out/${T}_score_form6 \
  python bench.py --steps 1 --warmup 3 --no-extras --no-parity-check --ingest-reads 0 --no-graph --lanes 1 > gpurun_out/${T}_ncu_score.log 2>&1
ncu -i gpurun_out/${T}_score_form6.ncu-rep --page raw --csv > gpurun_out/${T}_score_form6_ncu_raw.csv 2>/dev/null
ncu -i gpurun_out/${T}_score_form6.ncu-rep --page details > gpurun_out/${T}_score_form6_details.txt 2>/dev/null
python profiles/summarize_launches.py gpurun_out/${T}_launches.csv > gpurun_out/${T}_launches_summary.txt 2>&1; cat gpurun_out/${T}_launches_summary.txt
grep -E "Duration|DRAM Throughput|Issue Slots Busy|Achieved Occupancy|Memory Throughput" gpurun_out/${T}_score_form6_details.txt | head
python - <<P
import csv
rows = list(csv.reader(open("gpurun_out/${T}_score_form6_ncu_raw.csv")))
h = rows[0]; v = rows[-1]
for k in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"):
    if k in h: print(k, v[h.index(k)], rows[1][h.index(k)])
P
