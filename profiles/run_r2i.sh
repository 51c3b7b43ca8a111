#!/bin/bash
# configs[3] (64-sample cohort, 8 GPUs) and configs[2] (100 M-read metagenome over 8 GPUs, all-reduce form)
T=${1:-r2i}
nvidia-smi topo -m > gpurun_out/${T}_topo.txt 2>&1; free -g >> gpurun_out/${T}_topo.txt; nproc >> gpurun_out/${T}_topo.txt
timeout 900 python bench_configs.py c4 --samples 64 --reads 250000 --devices 8 > gpurun_out/${T}_c4.json 2> gpurun_out/${T}_c4.err; echo c4 rc=$?; tail -3 gpurun_out/${T}_c4.err
python -c "
import json; d=json.load(open('gpurun_out/${T}_c4.json')); print('c4', d['value'], d['samples_per_s'], d['per_sample_latency_ms'], d['per_sample_phase_ms_median'], d['parity']['ok'])"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 bench_configs.py c3 --reads ${2:-100000000} > gpurun_out/${T}_c3.json 2> gpurun_out/${T}_c3.err; echo c3 rc=$?; tail -4 gpurun_out/${T}_c3.err
python -c "
import json; d=json.load(open('gpurun_out/${T}_c3.json')); print('c3', d['value'], d['ms_per_step'], d['records_per_gpu'], d['loci_typed'], d['parity_locus_sample'], d['setup_seconds'])"
