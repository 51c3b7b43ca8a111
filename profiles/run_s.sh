#!/bin/bash
# GPU visit r1s (final of round 1): everything under the library defaults (score kernel form 5 = TMA ring, 4 chunks x 2 stages).
#   1. all GPU parity tests (stage-1 tests under forms 5, 0, 1, 2; the cmseq seam on the golden BAMs)
#   2. smoke
#   3. cohort lanes 2 / 3 / 4 (short runs), then the full bench line with the best lane count
#   4. reference arm
#   5. ncu launch list of the timed steps + one `ncu --set full` capture of the score kernel
TAG=${1:-r1s}
O=gpurun_out
mkdir -p $O
T0=$(date +%s)
timeout 300 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$? at $(( $(date +%s) - T0 )) s"
tail -3 $O/${TAG}_pytest_gpu.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/${TAG}_smoke.log
BEST=2; BESTV=0
for L in 2 3 4; do
  timeout 120 python bench.py --lanes $L --no-extras --steps 20 > $O/${TAG}_bench_lanes$L.json 2> $O/${TAG}_bench_lanes$L.err
  V=$(python -c "import json;print(int(json.load(open('$O/${TAG}_bench_lanes$L.json'))['value']))" 2>/dev/null || echo 0)
  echo "lanes $L: $V records/s at $(( $(date +%s) - T0 )) s"
  if [ "$V" -gt "$BESTV" ]; then BESTV=$V; BEST=$L; fi
done
echo "best lanes: $BEST"
timeout 420 python bench.py --lanes $BEST > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$? at $(( $(date +%s) - T0 )) s"; tail -3 $O/${TAG}_bench.err
python -c "import json;d=json.load(open('$O/${TAG}_bench.json'));print(d['value'],d['ms_per_step'],d['lanes'],d['roofline']['frac'],d['roofline']['kernel_form'],d['roofline']['ms_by_kernel_form'],d['kernel_ms_per_step'],d['e2e']['value'])" 2>&1
timeout 100 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference_arm.json 2> $O/${TAG}_bench_reference_arm.err; echo "reference arm rc=$?"
MMLST_CUDA_PROFILER=1 timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $O/${TAG}_launches_timed_steps.csv python bench.py --steps 4 --warmup 3 --no-extras --no-graph --lanes $BEST > $O/${TAG}_launches.log 2>&1
python profiles/summarize_launches.py $O/${TAG}_launches_timed_steps.csv > $O/${TAG}_launches_summary.txt 2>&1
cat $O/${TAG}_launches_summary.txt
timeout 240 ncu --set full --clock-control none --import-source on -k regex:score_runs -s 4 -c 1 -f -o $O/${TAG}_score_ring \
  python bench.py --steps 1 --warmup 3 --no-extras --no-graph > $O/${TAG}_ncu_score_ring.log 2>&1; echo "ncu full rc=$? at $(( $(date +%s) - T0 )) s"
ncu -i $O/${TAG}_score_ring.ncu-rep --page raw --csv > $O/${TAG}_score_ring_ncu_raw.csv 2>/dev/null
ncu -i $O/${TAG}_score_ring.ncu-rep --page details 2>/dev/null | grep -E "Duration|DRAM Throughput|Memory Throughput|L2 Hit|Achieved Occupancy|Theoretical Occ|Registers|Mem Busy|Max Bandwidth|Stall|Warp Cycles|Issue|Shared Memory Config|Dynamic Shared|Block Limit" | head -40 > $O/${TAG}_score_ring_details.txt
head -30 $O/${TAG}_score_ring_details.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $O/${TAG}_nvsmi.txt 2>&1
echo "elapsed $(( $(date +%s) - T0 )) s"
