#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.

    python scripts/ncu_summary.py launches gpurun_out/launches_X.csv profiles/r1_launches_X.txt
    python scripts/ncu_summary.py full gpurun_out/prof.ncu-rep profiles/r1_full_X.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__cycles_elapsed.max",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active"]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        v = v / 1e3 if r["Metric Unit"] in ("ns", "nsecond") else v
        k = r["Kernel Name"].split("(")[0][:70]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none  ({src})\n")
        f.write(f"# per-launch times are cold-cache and serialised: compare SHARES\n# total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k:70s} n={v[0]:6d} us={v[1]:12.1f} mean_us={v[1] / v[0]:10.2f} share={v[1] / tot:.4f}\n")
    print(open(dst).read()[:3000])


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on  ({src})\n")
        for i, h in enumerate(hdr):
            if h in ("Kernel Name",) or any(h == k or h.endswith(k) for k in KEYS):
                f.write(f"{h} [{units[i]}]: " + " | ".join(r[i][:60] for r in data) + "\n")
    print(open(dst).read())
    # per-kernel DRAM traffic per launch (mean over the captured launches) for bench.py's roofline.traffic
    import json, os, re
    ci = {h: i for i, h in enumerate(hdr)}
    def col(name):
        return next(i for h, i in ci.items() if h == name or h.endswith(name))
    def to_bytes(v, unit):
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]
        return float(v.replace(",", "")) * scale
    kn, rd, wr, du = col("Kernel Name"), col("dram__bytes_read.sum"), col("dram__bytes_write.sum"), col("gpu__time_duration.sum")
    agg = {}
    for r in data:
        name = re.sub(r"^void ", "", r[kn]).split("(")[0].split("<")[0]
        a = agg.setdefault(name, {"launches": 0, "dram_bytes": 0.0})
        a["launches"] += 1
        a["dram_bytes"] += to_bytes(r[rd], units[rd]) + to_bytes(r[wr], units[wr])
    path = os.path.join(os.path.dirname(dst), "ncu_traffic.json")
    old = json.load(open(path)) if os.path.exists(path) else {}
    for k, a in agg.items():
        old[k] = {"dram_bytes_per_launch": a["dram_bytes"] / a["launches"], "launches_captured": a["launches"], "source": os.path.basename(dst)}
    json.dump(old, open(path, "w"), indent=1, sort_keys=True)
    print("updated", path)


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
