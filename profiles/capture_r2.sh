#!/bin/bash
# Round-2 ncu evidence (one GPU, under gpurun).  Nothing printed by a bench run under ncu is a bench value.
T=${1:-r2m}
# 1. every launch of the timed steps of the default bench (serial eager passes so that each kernel is listed), cold-cache and serialised:
#    compare SHARES with kernel_ms_per_step of the bench line, not absolutes
MMLST_CUDA_PROFILER=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv \
  python bench.py --steps 4 --warmup 3 --no-extras --no-parity-check --ingest-reads 0 --no-graph --lanes 1 > gpurun_out/${T}_launches_bench.log 2>&1
# 2. the dominant kernel of the pass: score, form 6 (ring with pairs of chunks reduced together)
ncu --set full --clock-control none --import-source on -k regex:score_runs_ring_pair -s 6 -c 1 -f -o gpurun_out/${T}_score_form6 \
  python bench.py --steps 1 --warmup 3 --no-extras --no-parity-check --ingest-reads 0 --no-graph --lanes 1 > gpurun_out/${T}_ncu_score.log 2>&1
# 3. the tensor-core Hamming sweep
ncu --set full --clock-control none --import-source on -k regex:hamming_tc_kernel -s 1 -c 1 -f -o gpurun_out/${T}_hamming_tc \
  python bench.py --only-hamming > gpurun_out/${T}_ncu_hamming_tc.log 2>&1
# 4. launch list of one device ingest (2 M records)
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 120 --csv --log-file gpurun_out/${T}_ingest_launches.csv \
  python bench.py --only-ingest --ingest-reads 500000 > gpurun_out/${T}_ingest_launches_bench.log 2>&1
for k in score_form6 hamming_tc; do
  ncu -i gpurun_out/${T}_$k.ncu-rep --page raw --csv > gpurun_out/${T}_${k}_ncu_raw.csv 2>/dev/null
  ncu -i gpurun_out/${T}_$k.ncu-rep --page details > gpurun_out/${T}_${k}_details.txt 2>/dev/null
done
ls -la gpurun_out | tail -20
