#!/bin/bash
# multi-GPU bench, the way the driver launches it
T=${1:-r2p}; N=${2:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --steps 48 --warmup 3 > gpurun_out/${T}_bench_n${N}.json 2> gpurun_out/${T}_bench_n${N}.err; echo rc=$?; tail -4 gpurun_out/${T}_bench_n${N}.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_n${N}.json')); e=d['e2e']; print('N=$N value %.4e ms %.4f e2e %.4e (%.3f ms) plain %.4e' % (d['value'], d['ms_per_step'], e['value'], e['ms_per_step'], e['uncompressed']['value'])); print(d.get('parity_full_workload')); print(d.get('hamming_scaling')); print(e.get('host_numa_binding'))"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29656 bench.py --impl reference --gpus $N --steps 5 --warmup 1 > gpurun_out/${T}_ref_n${N}.json 2> gpurun_out/${T}_ref_n${N}.err; echo ref rc=$?; cut -c1-200 gpurun_out/${T}_ref_n${N}.json
