#!/bin/bash
# quick GPU iteration: score / pipeline parity tests + the main bench pass without extras
TAG=${1:-q}
python -m pytest tests -m gpu -x -q -k "${2:-score or pipeline or coverage}" 2>&1 | tail -3
python bench.py --no-extras > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
python - <<P
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print("value %.3e ms/step %.4f kernels %s" % (d["value"], d["ms_per_step"], {k: round(v*1e3,1) for k,v in d["kernel_ms_per_step"].items()}))
print("roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"],3), "e2e %.3e %.2f ms" % (d["e2e"]["value"], d["e2e"]["ms_per_step"]))
P
python - <<P
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print("serial_ms", d.get("serial_ms_per_step"), "lanes", d.get("lanes"), "launches", d.get("gpu_launches"))
P
