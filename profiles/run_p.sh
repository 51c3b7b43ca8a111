#!/bin/bash
# GPU visit r1p (diagnostic): the ring forms of the score kernel -- parity of the new stage configurations, time of every
# form at 1/1, 1/2, 1/4 of the stream (fixed cost vs streaming rate), ncu --set full of two ring configurations.
TAG=${1:-r1p}
O=gpurun_out
mkdir -p $O
T0=$(date +%s)
MMLST_TEST_SCORE_VARIANTS=3,4,5 timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "score or device_pipeline" > $O/${TAG}_pytest_gpu_forms345.log 2>&1; echo "pytest forms 3,4,5 rc=$? at $(( $(date +%s) - T0 )) s"
tail -2 $O/${TAG}_pytest_gpu_forms345.log
MMLST_BENCH_SCALING=1 timeout 300 python bench.py --score-variant auto --no-extras > $O/${TAG}_bench_auto.json 2> $O/${TAG}_bench_auto.err; echo "bench rc=$? at $(( $(date +%s) - T0 )) s"
grep "score form" $O/${TAG}_bench_auto.err
python -c "import json;d=json.load(open('$O/${TAG}_bench_auto.json'));print(d['value'],d['ms_per_step'],d['roofline']['frac'],d['roofline']['kernel_form'],d['roofline']['ms_by_kernel_form'])" 2>&1
for F in 2 4; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:score_runs -s 4 -c 1 -f -o $O/${TAG}_score_form$F \
    python bench.py --steps 1 --warmup 3 --no-extras --no-graph --score-variant $F > $O/${TAG}_ncu_form$F.log 2>&1; echo "ncu form $F rc=$? at $(( $(date +%s) - T0 )) s"
  ncu -i $O/${TAG}_score_form$F.ncu-rep --page raw --csv > $O/${TAG}_score_form${F}_ncu_raw.csv 2>/dev/null
  ncu -i $O/${TAG}_score_form$F.ncu-rep --page details 2>/dev/null | grep -E "Duration|DRAM Throughput|Memory Throughput|L2 Hit|Achieved Occupancy|Theoretical Occ|Registers|Mem Busy|Max Bandwidth|Stall|Warp Cycles|Issue|Shared Memory Config|Dynamic Shared|Block Limit" | head -40 > $O/${TAG}_score_form${F}_details.txt
  cat $O/${TAG}_score_form${F}_details.txt
  rm -f $O/${TAG}_score_form$F.ncu-rep
done
echo "elapsed $(( $(date +%s) - T0 )) s"
