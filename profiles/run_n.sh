#!/bin/bash
# One GPU-box visit (r1n): GPU parity tests, smoke, the default bench line, the reference arm, the ncu launch list of the
# timed steps and one `ncu --set full` capture of the score kernel.  Every step runs under its own timeout and none
# depends on the previous one succeeding.
TAG=${1:-r1n}
O=gpurun_out
mkdir -p $O
timeout 240 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 $O/${TAG}_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/${TAG}_smoke.log
timeout 400 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; rc=$?; echo "bench rc=$rc"; tail -3 $O/${TAG}_bench.err
if [ $rc -ne 0 ]; then
  timeout 300 python bench.py --no-qc > $O/${TAG}_bench_noqc.json 2> $O/${TAG}_bench_noqc.err; echo "bench --no-qc rc=$?"; tail -3 $O/${TAG}_bench_noqc.err
fi
timeout 120 python bench.py --impl reference --steps 5 --warmup 1 > $O/${TAG}_bench_reference_arm.json 2> $O/${TAG}_bench_reference_arm.err; echo "reference arm rc=$?"
MMLST_CUDA_PROFILER=1 timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $O/${TAG}_launches_timed_steps.csv python bench.py --steps 4 --warmup 3 --no-extras --no-graph > $O/${TAG}_launches.log 2>&1
python profiles/summarize_launches.py $O/${TAG}_launches_timed_steps.csv > $O/${TAG}_launches_summary.txt 2>&1
cat $O/${TAG}_launches_summary.txt
timeout 240 ncu --set full --clock-control none --import-source on -k regex:score_runs_kernel -s 4 -c 1 -f -o $O/${TAG}_score_runs_qc \
  python bench.py --steps 1 --warmup 3 --no-extras --no-graph > $O/${TAG}_ncu_score_runs.log 2>&1; echo "ncu full rc=$?"
ncu -i $O/${TAG}_score_runs_qc.ncu-rep --page raw --csv > $O/${TAG}_score_runs_qc_ncu_raw.csv 2>/dev/null
ncu -i $O/${TAG}_score_runs_qc.ncu-rep --page details 2>/dev/null | grep -E "Duration|DRAM Throughput|Memory Throughput|L2 Hit|Achieved Occupancy|Theoretical Occ|Registers|Mem Busy|Max Bandwidth|Stall|Warp Cycles|Issue" | head -40 > $O/${TAG}_score_runs_qc_details.txt
cat $O/${TAG}_score_runs_qc_details.txt | head -20
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $O/${TAG}_nvsmi.txt 2>&1
head -c 1200 $O/${TAG}_bench.json
