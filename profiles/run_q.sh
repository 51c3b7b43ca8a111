#!/bin/bash
# GPU visit r1q (diagnostic): ring forms with the len(SEQ) register prefetch and the L2 residency hints -- parity with the
# hints on, then every (form, hints) pair timed on the workload.
TAG=${1:-r1q}
O=gpurun_out
mkdir -p $O
T0=$(date +%s)
MMLST_SCORE_L2_HINTS=1 MMLST_TEST_SCORE_VARIANTS=2,3,4,5 timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "score or device_pipeline" > $O/${TAG}_pytest_gpu_ring_hints.log 2>&1; echo "pytest ring forms + hints rc=$? at $(( $(date +%s) - T0 )) s"
tail -2 $O/${TAG}_pytest_gpu_ring_hints.log
timeout 300 python bench.py --score-variant auto --no-extras > $O/${TAG}_bench_auto.json 2> $O/${TAG}_bench_auto.err; echo "bench rc=$? at $(( $(date +%s) - T0 )) s"
tail -2 $O/${TAG}_bench_auto.err
python -c "import json;d=json.load(open('$O/${TAG}_bench_auto.json'));print(d['value'],d['ms_per_step'],d['roofline']['frac'],d['roofline']['kernel_form'],d['roofline']['l2_hints'],d['roofline']['ms_by_kernel_form'],d['kernel_ms_per_step'])" 2>&1
echo "elapsed $(( $(date +%s) - T0 )) s"
