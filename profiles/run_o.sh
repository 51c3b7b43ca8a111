#!/bin/bash
# One GPU-box visit (r1o), most important first, every step under its own timeout, none depending on the previous one:
#   1. GPU parity tests with the library's default kernel forms
#   2. the stage-1 / device-pipeline parity tests again under the other two forms of the score kernel (TMA ring = form 2)
#   3. the bench line, score-kernel form chosen by timing all three on the workload (--score-variant auto)
#   4. reference arm, smoke
#   5. ncu launch list of the timed steps and one `ncu --set full` capture of the score kernel in the chosen form
TAG=${1:-r1o}
O=gpurun_out
mkdir -p $O
date +%s > $O/${TAG}_t0.txt
timeout 200 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$? at $(( $(date +%s) - $(cat $O/${TAG}_t0.txt) )) s"
tail -3 $O/${TAG}_pytest_gpu.log
MMLST_TEST_SCORE_VARIANTS=1,2 timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "score or device_pipeline" > $O/${TAG}_pytest_gpu_forms12.log 2>&1; echo "pytest forms 1,2 rc=$?"
tail -3 $O/${TAG}_pytest_gpu_forms12.log
timeout 420 python bench.py --score-variant auto > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; rc=$?; echo "bench rc=$rc at $(( $(date +%s) - $(cat $O/${TAG}_t0.txt) )) s"; tail -3 $O/${TAG}_bench.err
FORM=$(python -c "import json;print(json.load(open('$O/${TAG}_bench.json'))['roofline']['kernel_form'])" 2>/dev/null || echo 0)
echo "score kernel form chosen: $FORM"
python -c "import json;d=json.load(open('$O/${TAG}_bench.json'));print(d['value'],d['ms_per_step'],d['roofline']['frac'],d['roofline']['ms_by_kernel_form'],d['kernel_ms_per_step'])" 2>&1
timeout 100 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference_arm.json 2> $O/${TAG}_bench_reference_arm.err; echo "reference arm rc=$?"
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/${TAG}_smoke.log
MMLST_CUDA_PROFILER=1 timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $O/${TAG}_launches_timed_steps.csv python bench.py --steps 4 --warmup 3 --no-extras --no-graph --score-variant $FORM > $O/${TAG}_launches.log 2>&1
python profiles/summarize_launches.py $O/${TAG}_launches_timed_steps.csv > $O/${TAG}_launches_summary.txt 2>&1
cat $O/${TAG}_launches_summary.txt
timeout 240 ncu --set full --clock-control none --import-source on -k regex:score_runs -s 4 -c 1 -f -o $O/${TAG}_score_runs \
  python bench.py --steps 1 --warmup 3 --no-extras --no-graph --score-variant $FORM > $O/${TAG}_ncu_score_runs.log 2>&1; echo "ncu full rc=$? at $(( $(date +%s) - $(cat $O/${TAG}_t0.txt) )) s"
ncu -i $O/${TAG}_score_runs.ncu-rep --page raw --csv > $O/${TAG}_score_runs_ncu_raw.csv 2>/dev/null
ncu -i $O/${TAG}_score_runs.ncu-rep --page details 2>/dev/null | grep -E "Duration|DRAM Throughput|Memory Throughput|L2 Hit|Achieved Occupancy|Theoretical Occ|Registers|Mem Busy|Max Bandwidth|Stall|Warp Cycles|Issue|Shared Memory Config|Dynamic Shared" | head -40 > $O/${TAG}_score_runs_details.txt
head -24 $O/${TAG}_score_runs_details.txt
timeout 200 python bench.py --lanes 3 --no-extras --score-variant $FORM > $O/${TAG}_bench_lanes3.json 2> $O/${TAG}_bench_lanes3.err; echo "bench lanes3 rc=$?"
python -c "import json;d=json.load(open('$O/${TAG}_bench_lanes3.json'));print('lanes3',d['value'],d['ms_per_step'])" 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $O/${TAG}_nvsmi.txt 2>&1
echo "elapsed $(( $(date +%s) - $(cat $O/${TAG}_t0.txt) )) s"
