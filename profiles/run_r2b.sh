#!/bin/bash
# round-2 GPU visit b: full GPU suite, bench (new CPU baseline + full-workload parity), reference arm, DE alignment probes
T=${1:-r2b}
for a in "8 65280" "1 65280" "1 65279" "1 64512"; do set -- $a; ./profiles/tools/de_probe 2048 $1 $2 >> gpurun_out/${T}_de_probe_align.jsonl 2>&1; done
cat gpurun_out/${T}_de_probe_align.jsonl
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${T}_pytest_gpu.log; cat gpurun_out/${T}_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo bench rc=$?; tail -3 gpurun_out/${T}_bench.err
timeout 400 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; echo ref rc=$?; tail -2 gpurun_out/${T}_bench_reference.err
python - <<P
import json
d=json.load(open("gpurun_out/${T}_bench.json"))
print("value %.3e ms %.4f e2e %.3e" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
print("roofline", d["roofline"]["kernel"], d["roofline"]["frac"], d["kernel_ms_per_step"])
print("parity", d.get("parity_full_workload"))
print("cpu", d.get("cpu_baseline"))
r=json.load(open("gpurun_out/${T}_bench_reference.json"))
print("ref %.3e" % r["value"], r["seconds_by_phase"], r["cpu_baseline"]["sample"])
P
