#!/bin/bash
# GPU visit r1t: cohort lanes 4 / 6 / 8 (short runs), then the full default bench line with the best lane count.
TAG=${1:-r1t}
O=gpurun_out
mkdir -p $O
T0=$(date +%s)
BEST=4; BESTV=0
for L in 4 6 8; do
  timeout 120 python bench.py --lanes $L --no-extras --steps 24 > $O/${TAG}_bench_lanes$L.json 2> $O/${TAG}_bench_lanes$L.err
  V=$(python -c "import json;print(int(json.load(open('$O/${TAG}_bench_lanes$L.json'))['value']))" 2>/dev/null || echo 0)
  echo "lanes $L: $V records/s at $(( $(date +%s) - T0 )) s"
  if [ "$V" -gt "$BESTV" ]; then BESTV=$V; BEST=$L; fi
done
echo "best lanes: $BEST"
timeout 420 python bench.py --lanes $BEST > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$? at $(( $(date +%s) - T0 )) s"; tail -3 $O/${TAG}_bench.err
python -c "import json;d=json.load(open('$O/${TAG}_bench.json'));print(d['value'],d['ms_per_step'],d['lanes'],d['roofline']['frac'],d['roofline']['ms_by_kernel_form'],d['kernel_ms_per_step'],d['e2e']['value'])" 2>&1
echo "elapsed $(( $(date +%s) - T0 )) s"
