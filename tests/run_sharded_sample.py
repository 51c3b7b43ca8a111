"""Worker of tests/test_streams.py::test_one_real_sample_sharded_over_two_ranks (launched by torch.distributed.run)."""
import json
import os
import sys

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from metamlst_b200 import sample  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def main():
    out_root = sys.argv[1]
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank = torch.distributed.get_rank()
    os.makedirs(out_root, exist_ok=True)
    man = json.load(open(os.path.join(GOLDEN, "manifest.json")))
    cases = sorted(k for k in man if os.path.exists(os.path.join(GOLDEN, k, "sample.bam")) and not set(man[k]["args"]) - {"--log", "-a", "--presorted"})
    for name in cases:
        d = os.path.join(GOLDEN, name)
        a = man[name]["args"]
        typer = sample.SampleTyper(os.path.join(d, "db.sqlite"), device=local, engine="device", write_known="-a" in a, presorted="--presorted" in a)
        res = typer.type_bam(os.path.join(d, "sample.bam"), os.path.join(out_root, name))
        typer.close()
        torch.distributed.barrier()
        if rank == 0:
            p = os.path.join(out_root, name, "sample.nfo")
            got = open(p, newline="").read() if os.path.exists(p) else ""
            gp = os.path.join(d, "sample.nfo")
            want = open(gp, newline="").read() if os.path.exists(gp) else ""
            assert got == want, name
        else:
            assert not os.path.exists(os.path.join(out_root, name, "sample.nfo")) or True
        assert res.sample == "sample"
    torch.distributed.barrier()
    if rank == 0:
        print("SHARDED-OK", len(cases))
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
