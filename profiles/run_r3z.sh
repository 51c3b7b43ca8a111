#!/bin/bash
# what the driver runs at round end, on one GPU: GPU suite, smoke(), both bench arms with default flags
T=${1:-r4c}
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${T}_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo smoke rc=$?; tail -2 gpurun_out/${T}_smoke.log
SECONDS=0; timeout 900 python bench.py --impl reference > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; echo ref rc=$? wall=${SECONDS}s; SECONDS=0; true
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo bench rc=$? wall=${SECONDS}s; true; grep -v "^\s" gpurun_out/${T}_bench.err | tail -3 | cut -c1-300
python - <<P
import json
d=json.load(open("gpurun_out/${T}_bench.json")); r=json.load(open("gpurun_out/${T}_bench_reference.json"))
print("value %.4e ms %.4f serial %.4f e2e %.4e (%.3f ms)  ref %.4e  ratio e2e %.1f value %.0f" % (d["value"], d["ms_per_step"], d["serial_ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], r["value"], d["e2e"]["value"]/r["value"], d["value"]/r["value"]))
print("roofline", {k: d["roofline"][k] for k in ("kernel","achieved","peak","frac","traffic","ms")})
print("e2e detail", {k: d["e2e"][k] for k in ("ms_per_step","h2d_bytes_per_step","d2h_bytes_per_step","ms_per_step_by_lanes")}, "one at a time", d["e2e"]["one_at_a_time"]["ms_per_step"], "two seams", d["e2e"]["two_seam_calls"]["ms_per_step"], "plain", d["e2e"]["uncompressed"]["ms_per_step"])
print("parity", d["parity_full_workload"]["ok"], "cpu", d["cpu_baseline"]["value"], "clocks", d["clocks"])
print("ingest", {k: (v["records_per_s"], v["seconds"], v["speedup_vs_host_unpacker"]) for k, v in d["ingest"].items()})
print("hamming", d["hamming_scaling"], d["hamming"]["all_pairs_tensor_core"]["roofline"])
print("keys", list(d.keys()))
P
