#!/bin/bash
T=${1:-r2y}
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
for f in 1 0; do
  MMLST_FUSED_TAIL=$f timeout 400 python bench.py --no-extras --ingest-reads 0 > gpurun_out/${T}_bench_fused${f}.json 2> gpurun_out/${T}_bench_fused${f}.err; echo "fused=$f rc=$?"; tail -2 gpurun_out/${T}_bench_fused${f}.err | cut -c1-200
  python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_fused${f}.json')); print('fused=$f value %.4e ms %.4f serial %.4f lat %.4f launches %d parity %s' % (d['value'], d['ms_per_step'], d['serial_ms_per_step'], d['latency_ms_per_step_with_host_sync'], d['gpu_launches'], d['parity_full_workload']['ok']))"
done
