#!/usr/bin/env python
"""Summarise an `ncu --set full` report (read here, no GPU needed) into profiles/:
    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep r01c
writes profiles/ncu_summary.json (what bench.py reads for `roofline.traffic`), a per-kernel metric
table profiles/<tag>_ncu_metrics.csv and the top stall lines per kernel profiles/<tag>_ncu_stalls.txt."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_bytes_read",
    "dram__bytes_write.sum": "dram_bytes_write",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__cycles_elapsed.max": "sm_cycles",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
}
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6}


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    out, table = {}, []
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void <unnamed>::", "")
        d = {}
        for k, short in KEYS.items():
            if k in hdr:
                i = hdr.index(k)
                try:
                    d[short] = float(r[i].replace(",", "")) * UNIT.get(units[i], 1)
                except ValueError:
                    pass
        stalls = {h.split("issue_stalled_")[1].split("_per_issue")[0]: float(r[i]) for i, h in enumerate(hdr)
                  if "average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h and r[i]}
        d["top_stalls"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:5])
        if name not in out:  # first captured launch of each kernel
            out[name] = d
            table.append((name, d))
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    json.dump({"source": os.path.basename(rep), "tag": tag, "kernels": out},
              open(os.path.join(ROOT, "profiles", "ncu_summary.json"), "w"), indent=1)
    with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_metrics.csv"), "w") as f:
        w = csv.writer(f)
        cols = list(KEYS.values())
        w.writerow(["kernel"] + cols + ["top_stalls"])
        for name, d in table:
            w.writerow([name] + [d.get(c, "") for c in cols] + [json.dumps(d["top_stalls"])])
    # per-kernel top stall source lines
    with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_stalls.txt"), "w") as f:
        for name, _ in table:
            key = name.split("<")[0]
            txt = ncu(["-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", f"regex:{key}",
                       "--launch-count", "1"]) if False else ncu(["-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", f"regex:{key}"])
            cur, acc = None, {}
            for r in csv.reader(io.StringIO(txt)):
                if len(r) >= 2 and r[0] == "File Path":
                    cur = os.path.basename(r[1])
                elif r and r[0].isdigit() and len(r) > 5 and r[4].isdigit():
                    k2 = (cur, int(r[0]), r[1].strip()[:110])
                    acc[k2] = acc.get(k2, 0) + int(r[4])
            tot = sum(acc.values()) or 1
            f.write(f"== {name}: warp stall samples by source line (total {tot})\n")
            for (fn, ln, src), n in sorted(acc.items(), key=lambda kv: -kv[1])[:25]:
                f.write(f"{n:7d} {100 * n / tot:5.1f}%  {fn}:{ln}  {src}\n")
            f.write("\n")
    print(json.dumps({k: {kk: vv for kk, vv in v.items() if kk != "top_stalls"} for k, v in out.items()}, indent=1))


if __name__ == "__main__":
    main()
