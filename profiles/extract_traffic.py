#!/usr/bin/env python3
"""Turn `ncu --page raw --csv` dumps (profiles/capture_full.sh) into profiles/traffic.json + a readable summary.

    python profiles/extract_traffic.py TAG=r1h score=gpurun_out/r1h_score_ncu_raw.csv pileup_uncapped_bitsliced=... hamming=...

traffic.json[kernel] = {"dram_bytes_per_launch": dram__bytes_read.sum + dram__bytes_write.sum, ...}; bench.py reports it
as roofline.traffic (bytes the kernel really moved, to set beside the algorithmic bytes)."""
import csv
import json
import os
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg",
        "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.avg.per_second"]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "second": 1.0,
        "usecond": 1e-6, "nsecond": 1e-9, "msecond": 1e-3}


def rows(path):
    r = list(csv.reader(open(path)))
    hdr, units = r[0], r[1]
    for row in r[2:]:
        d = {}
        for h, u, v in zip(hdr, units, row):
            d[h] = (v, u)
        yield d


def num(d, k):
    if k not in d or d[k][0] in ("", "n/a"):
        return None
    v, u = d[k]
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return None
    if x != x:  # ncu prints nan when a counter overflowed during a long replay
        return None
    return x * UNIT.get(u, 1.0)


def main():
    out_json = os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json")
    traffic = json.load(open(out_json)) if os.path.exists(out_json) else {}
    tag = "r"
    lines = []
    for arg in sys.argv[1:]:
        name, path = arg.split("=", 1)
        if name == "TAG":
            tag = path
            continue
        per = {}
        for d in rows(path):
            kn = d["Kernel Name"][0]
            grid = d.get("Grid Size", ("", ""))[0]
            per.setdefault((kn, grid), []).append(d)
        for (kn, grid), ds in per.items():
            d = ds[-1]
            t = num(d, "gpu__time_duration.sum")
            rd, wr = num(d, "dram__bytes_read.sum"), num(d, "dram__bytes_write.sum")
            key = name if len(per) == 1 else "%s grid=%s" % (name, grid)
            traffic[key] = {"kernel": kn, "grid": grid, "capture": tag, "launches_captured": len(ds), "duration_us_under_ncu": t * 1e6 if t else None,
                            "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": ((rd or 0) + (wr or 0)) if rd is not None else None}
            lines.append("== %s  [%s]  grid %s  x%d" % (key, kn[:70], grid, len(ds)))
            for k in KEYS:
                if k in d and d[k][0] != "":
                    lines.append("   %-72s %s %s" % (k, d[k][0], d[k][1]))
            # top stall reasons
            stalls = sorted(((float(v[0].replace(",", "")), k) for k, v in d.items()
                             if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio") and v[0] not in ("", "n/a", "nan", "-nan")), reverse=True)[:5]
            for v, k in stalls:
                lines.append("   stall %-66s %.2f" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
    json.dump(traffic, open(out_json, "w"), indent=1, sort_keys=True)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
