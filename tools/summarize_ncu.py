"""Summarise ncu CSV exports into the small tracked files under profiles/ (run on the CPU box; no GPU needed).

    python tools/summarize_ncu.py gpurun_out/prof_r1_block0_bf16_raw.csv profiles/r1_ncu_block0_bf16_summary.json
    python tools/summarize_ncu.py --launches gpurun_out/launches_warm_bf16.csv profiles/r1_launches_warm_bf16_summary.json
"""
import collections
import csv
import json
import sys


def launches(path, out):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]; ki = h.index("Kernel Name"); vi = h.index("Metric Value"); gi = h.index("Grid Size")
    agg = collections.OrderedDict(); total = 0.0; seq = []
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("void ", "").strip()
        us = float(r[vi].replace(",", "")) / 1e3
        a = agg.setdefault(name, {"launches": 0, "total_us": 0.0}); a["launches"] += 1; a["total_us"] += us; total += us
        seq.append([name, r[gi], round(us, 2)])
    for a in agg.values():
        a["avg_us"] = round(a["total_us"] / a["launches"], 2); a["share"] = round(a["total_us"] / total, 4); a["total_us"] = round(a["total_us"], 1)
    json.dump({"source": path, "total_us": round(total, 1), "kernels": agg, "first_block_sequence": seq[:32]}, open(out, "w"), indent=1)
    print(json.dumps({"total_us": round(total, 1), "kernels": agg}, indent=1))


def full(path, out):
    rows = list(csv.reader(open(path)))
    h, units = rows[0], rows[1]
    idx = {n: i for i, n in enumerate(h)}
    want = {"gpu__time_duration.sum": "duration", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
            "launch__registers_per_thread": "registers", "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
            "launch__grid_size": "grid", "launch__block_size": "block", "sm__inst_executed.sum": "inst_executed"}
    out_rows = []
    for r in rows[2:]:
        e = {"kernel": r[idx["Kernel Name"]].split("(")[0].replace("void ", "").strip(), "grid": r[idx["Grid Size"]], "block": r[idx["Block Size"]]}
        for k, short in want.items():
            if k in idx:
                e[short] = f"{r[idx[k]]} {units[idx[k]]}".strip()
        out_rows.append(e)
    json.dump({"source": path, "note": "ncu --set full --clock-control none (default cache control: caches flushed before every launch)",
               "kernels": out_rows}, open(out, "w"), indent=1)
    for e in out_rows:
        print(e)


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[1], sys.argv[2])
