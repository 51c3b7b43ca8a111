#!/bin/bash
T=${1:-r2n}
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "deflated or score_bit_exact" 2>&1 | tail -5
timeout 600 python -m pytest tests/test_ingest_gpu.py tests/test_hamming_tc.py tests/test_dropin.py -x -q 2>&1 | tail -4
timeout 500 python bench.py --no-extras --ingest-reads 0 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo rc=$?; tail -3 gpurun_out/${T}_bench.err
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench.json')); e=d['e2e']; print('value %.4e e2e %.4e (%.3f ms, h2d %d) plain %.4e (%.3f ms) ' % (d['value'], e['value'], e['ms_per_step'], e['h2d_bytes_per_step'], e['uncompressed']['value'], e['uncompressed']['ms_per_step'])); print(e['stream_form'])"
