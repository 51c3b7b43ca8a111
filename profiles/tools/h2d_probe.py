#!/usr/bin/env python
"""Host -> device copy bandwidth with k of the N ranks copying at once (one process per GPU, torchrun): what the box gives the host-buffer path when
every GPU pulls its sample over PCIe at the same time.  256 MB from page-locked memory per copy, 20 copies per measurement, wall clock between barriers.
Usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/tools/h2d_probe.py [out.json]"""
import json
import os
import sys
import time

import torch
import torch.distributed as td

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
td.init_process_group("nccl", device_id=torch.device("cuda:%d" % local))
n = 256 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h.fill_(rank + 1)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
tok = torch.zeros(1, device="cuda")
out = {}
ks = [k for k in (1, 2, 4, 8) if k <= world]
for k in ks:
    for who in ("first", "strided"):
        active = set(range(k)) if who == "first" else set(range(0, world, world // k))
        for _ in range(3):
            if rank in active:
                d.copy_(h, non_blocking=True)
        torch.cuda.synchronize(); td.all_reduce(tok); torch.cuda.synchronize()
        t0 = time.perf_counter()
        if rank in active:
            for _ in range(20):
                d.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0 if rank in active else 0.0], device="cuda", dtype=torch.float64)
        td.all_reduce(dt, op=td.ReduceOp.MAX)
        td.all_reduce(tok); torch.cuda.synchronize()
        gbs = 20 * n * k / float(dt.item()) / 1e9
        out["%d ranks (%s)" % (k, ",".join(map(str, sorted(active))))] = {"aggregate_GBps": gbs, "per_gpu_GBps": gbs / k}
if rank == 0:
    for key, v in out.items():
        print("%-28s aggregate %7.1f GB/s   per GPU %6.1f GB/s" % (key, v["aggregate_GBps"], v["per_gpu_GBps"]))
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)
td.destroy_process_group()
