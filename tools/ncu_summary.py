"""Condense `ncu -i X.ncu-rep --page raw --csv` exports into the figures DESIGN.md / bench.py quote:
    python tools/ncu_summary.py gpurun_out/full_raw.csv [gpurun_out/replay_raw.csv] > profiles/<name>.csv
writes a per-launch table (duration, DRAM bytes, tensor-pipe activity, registers, grid) to stdout and refreshes
profiles/ncu_summary.json (what bench.py attaches to its roofline object: only a profiler sees DRAM bytes)."""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_pct"), ("launch__registers_per_thread", "regs"),
        ("lts__t_sector_hit_rate.pct", "l2_hit_pct")]


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in data:
        rec = {"kernel": re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", ""), "grid": r[idx["Grid Size"]].replace(" ", "")}
        for col, short in COLS:
            if col not in idx:
                continue
            try:
                v = float(r[idx[col]].replace(",", ""))
            except ValueError:
                continue
            u = units[idx[col]]
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
            rec[short] = v * scale
        out.append(rec)
    return out


def table(recs):
    names = ["kernel", "grid"] + [s for _, s in COLS]
    print(",".join(names))
    for r in recs:
        print(",".join(str(round(r[k], 3)) if isinstance(r.get(k), float) else str(r.get(k, "")) for k in names))


upd = load(sys.argv[1])
table(upd)
summary = {"source": "ncu --set full --clock-control none over tools/profile_step.py (one step, CUDA graphs off) and tools/profile_replay.py: "
                     "profiles/r2_ncu_full_kernels.csv"}
fwd = [r for r in upd if r["kernel"].startswith("mlp_fwd")]
if fwd:
    summary["fwd_dram_bytes_per_launch"] = sum(r["dram_rd"] + r["dram_wr"] for r in fwd) / len(fwd)
tp = {}
for r in upd:
    if r.get("tensor_pct", 0) > 0:
        tp.setdefault(f"{r['kernel']} {r['grid']}", []).append(round(r["tensor_pct"], 1))
summary["tensor_pipe_active_pct"] = {k: (v[0] if len(set(v)) == 1 else [min(v), max(v)]) for k, v in tp.items()}
summary["tensor_pipe_active_pct"]["metric"] = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active, per launch (min, max over the step's launches)"
if len(sys.argv) > 2:
    rep = load(sys.argv[2])
    print()
    table(rep)
    traffic = {}
    for k, r in enumerate(rep):
        traffic[f"launch {k}: {r['kernel']} {r['grid']}"] = {"dram_read": r["dram_rd"], "dram_write": r["dram_wr"], "us": r["us"],
                                                            "dram_gbs": round((r["dram_rd"] + r["dram_wr"]) / r["us"] / 1e3, 1)}
    summary["replay_traffic_note"] = ("tools/profile_replay.py: inserts of 4096 / 122880 / 491520 rows walking the ring, then gathers of 8192 / 65536 / "
                                      "262144 fresh indices; DRAM bytes and durations under ncu (cold L2, serialised)")
    summary["replay_traffic"] = traffic
json.dump(summary, open(os.path.join(ROOT, "profiles", "ncu_summary.json"), "w"), indent=1)
