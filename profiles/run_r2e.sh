#!/bin/bash
# small-size shakedown of bench_configs.py on 2 GPUs
T=${1:-r2e}
timeout 600 python bench_configs.py c4 --samples 8 --reads 50000 --devices 2 > gpurun_out/${T}_c4_small.json 2> gpurun_out/${T}_c4_small.err; echo c4 rc=$?; tail -5 gpurun_out/${T}_c4_small.err; cat gpurun_out/${T}_c4_small.json | cut -c1-1500
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench_configs.py c3 --reads 2000000 --organisms 20 --present 6 --max-alleles 400 > gpurun_out/${T}_c3_small.json 2> gpurun_out/${T}_c3_small.err; echo c3 rc=$?; tail -8 gpurun_out/${T}_c3_small.err; cat gpurun_out/${T}_c3_small.json | cut -c1-1800
timeout 600 python -m pytest tests/test_streams.py -m gpu -x -q 2>&1 | tail -5
