#!/usr/bin/env python3
"""Summarise ncu exports for profiles/: (1) a launch list CSV (gpu__time_duration.sum per launch) -> per-kernel
totals and shares; (2) the raw page of a --set full capture -> the handful of metrics DESIGN.md/bench.py quote."""
import csv
import subprocess
import sys
from collections import defaultdict


def launch_summary(path):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr = rows[0]
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        name = r[ik].split("(")[0].replace("void zkb::", "").replace("zkb::", "")
        tot[name] += float(r[iv].replace(",", ""))
        cnt[name] += 1
    unit = rows[1][hdr.index("Metric Unit")] if len(rows) > 1 else "ns"
    scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "msecond": 1.0, "ms": 1.0}.get(unit, 1e-6)
    total = sum(tot.values())
    print(f"# launch list {path}: {sum(cnt.values())} launches, {total * scale:.2f} ms total (cold-cache, serialised: compare shares)")
    print(f"{'kernel':60s} {'launches':>8s} {'ms':>10s} {'share':>7s}")
    for k in sorted(tot, key=tot.get, reverse=True):
        print(f"{k[:60]:60s} {cnt[k]:8d} {tot[k] * scale:10.3f} {tot[k] / total:7.1%}")


KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]


def full_summary(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full capture {rep}")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("kernel:", d.get("Kernel Name", "?")[:110])
        for k in KEYS:
            if k in d:
                print(f"  {k:85s} {d[k]:>16s} {units[hdr.index(k)]}")


def raw_summary(path, traffic_json=None):
    """`ncu -i rep --page raw --csv` export (made on the GPU box when the .ncu-rep is too large to bring back):
    the same metric digest, plus - with traffic_json - the DRAM traffic per launch of the dominant kernel
    (the largest k_accumulate<Fq> launch) for bench.py's roofline.traffic."""
    import json
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full capture, raw page export {path}")
    best = None
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("kernel:", d.get("Kernel Name", "?")[:110], "grid", d.get("launch__grid_size"))
        for k in KEYS:
            if k in d:
                print(f"  {k:85s} {d[k]:>16s} {units[hdr.index(k)]}")
        if "k_accumulate<Fp<FqParams>" in d.get("Kernel Name", "") or "k_accumulate<zkb::Fp<zkb::FqParams>" in d.get("Kernel Name", ""):
            g = int(float(d["launch__grid_size"]))
            if best is None or g > best[0]:
                best = (g, d)
    if traffic_json and best:
        d = best[1]

        def to_bytes(key):
            v, u = float(d[key]), units[hdr.index(key)].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}[u]
        out = {"kernel": "k_accumulate<Fq> (H MSM launch of one 128-proof chunk)", "grid_blocks": best[0],
               "dram_bytes_read": to_bytes("dram__bytes_read.sum"), "dram_bytes_write": to_bytes("dram__bytes_write.sum"),
               "duration_ms_under_ncu": float(d["gpu__time_duration.sum"]), "source": path}
        out["dram_bytes"] = out["dram_bytes_read"] + out["dram_bytes_write"]
        json.dump(out, open(traffic_json, "w"), indent=1)
        print("# traffic ->", traffic_json, out)


if __name__ == "__main__":
    args = sys.argv[1:]
    if args and args[0] == "--raw":
        raw_summary(args[1], args[2] if len(args) > 2 else None)
    else:
        for a in args:
            (full_summary if a.endswith(".ncu-rep") else launch_summary)(a)
