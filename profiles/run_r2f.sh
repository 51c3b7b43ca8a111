#!/bin/bash
T=${1:-r2f}
timeout 600 python -m pytest tests/test_streams.py -m gpu -x -q -k two_ranks > gpurun_out/${T}_sharded.log 2>&1; tail -60 gpurun_out/${T}_sharded.log | cut -c1-400
timeout 600 python -m pytest tests/test_ingest_gpu.py -x -q 2>&1 | tail -5
CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py --only-ingest > gpurun_out/${T}_ingest.json 2> gpurun_out/${T}_ingest.err; echo rc=$?; tail -5 gpurun_out/${T}_ingest.err
python - <<P
import json
d=json.load(open("gpurun_out/${T}_ingest.json"))["ingest"]
for k,v in d.items():
    print(k, "records %d  %.3e rec/s  %.1f ms  bam %.2f GB/s  host %.3e rec/s  speedup %.1f" % (v["records"], v["records_per_s"], v["seconds"]*1e3, v["bam_GBps"], v["host_unpacker"]["records_per_s"], v["speedup_vs_host_unpacker"]))
    print("   phases(ms)", {a: round(b*1e3,2) for a,b in v["device_seconds_by_phase"].items()}, "sum %.1f" % (v["device_seconds_total"]*1e3), "bam->result %.1f ms" % (v["bam_to_result_seconds"]*1e3), "repairs", v["boundary_repairs"])
P
