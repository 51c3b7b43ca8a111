#!/bin/bash
# PDL on/off, wide Hamming, whole GPU suite
T=${1:-r2l}
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/${T}_pytest_gpu.log; cat gpurun_out/${T}_pytest_gpu.log
for pdl in 1 0; do
  MMLST_PDL=$pdl timeout 400 python bench.py --no-extras --no-parity-check --ingest-reads 0 > gpurun_out/${T}_bench_pdl${pdl}.json 2> gpurun_out/${T}_bench_pdl${pdl}.err; echo "pdl=$pdl rc=$?"; tail -2 gpurun_out/${T}_bench_pdl${pdl}.err
  python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_pdl${pdl}.json')); print('pdl=$pdl value %.4e ms %.4f serial %.4f lat %.4f graph %s e2e %.3e' % (d['value'], d['ms_per_step'], d['serial_ms_per_step'], d['latency_ms_per_step_with_host_sync'], d['cuda_graph'], d['e2e']['value']), d['kernel_ms_per_step'])"
done
