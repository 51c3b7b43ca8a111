#!/bin/bash
# One GPU-box visit: GPU parity tests, smoke, the default bench line, and the ncu launch list of the timed steps.
TAG=${1:-r1k}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" 
tail -5 gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
MMLST_CUDA_PROFILER=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/${TAG}_launches_timed_steps.csv python bench.py --steps 3 --warmup 3 --no-extras --no-graph > gpurun_out/${TAG}_launches.log 2>&1
python profiles/summarize_launches.py gpurun_out/${TAG}_launches_timed_steps.csv > gpurun_out/${TAG}_launches_summary.txt 2>&1
cat gpurun_out/${TAG}_launches_summary.txt
head -c 1500 gpurun_out/${TAG}_bench.json
