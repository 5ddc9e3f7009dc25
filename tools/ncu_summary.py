"""Condenses gpurun_out/*.ncu-rep into profiles/<round>_ncu_summary.csv (run in the build container)."""
import csv
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "sm__cycles_elapsed.max"]


def main(tag):
    out_rows = []
    reps = sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "*.ncu-rep")) + glob.glob(os.path.join(ROOT, "gpurun_out", "ncu_raw", "*.csv")))
    for rep in reps:
        if rep.endswith(".csv"):          # raw page exported on the GPU box (tools/ncu_round.sh)
            text = open(rep).read()
        else:
            text = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(text.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            d = dict(zip(hdr, vals))
            u = dict(zip(hdr, units))
            rec = {"report": os.path.basename(rep).replace(".csv", ".ncu-rep"), "kernel": d.get("Kernel Name", "")[:80]}
            for w in WANT:
                if w in d:
                    rec[w + (" [" + u[w] + "]" if u.get(w) else "")] = d[w]
            out_rows.append(rec)
    keys = []
    for r in out_rows:
        for k in r:
            if k not in keys:
                keys.append(k)
    path = os.path.join(ROOT, "profiles", f"{tag}_ncu_summary.csv")
    with open(path, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=keys)
        w.writeheader()
        w.writerows(out_rows)
    print(path, len(out_rows), "kernels")
    # DRAM traffic per launch (read + write) for bench.py's roofline.traffic
    import json
    traffic = {}
    for r in out_rows:
        rd = next((float(v) for k, v in r.items() if k.startswith("dram__bytes_read.sum [")), None)
        wr = next((float(v) for k, v in r.items() if k.startswith("dram__bytes_write.sum [")), None)
        ru = next((k for k in r if k.startswith("dram__bytes_read.sum [")), "")
        mult = {"[Gbyte]": 1e9, "[Mbyte]": 1e6, "[Kbyte]": 1e3, "[byte]": 1.0}
        def scale(key):
            for u, m in mult.items():
                if key.endswith(u):
                    return m
            return 1.0
        wu = next((k for k in r if k.startswith("dram__bytes_write.sum [")), "")
        if rd is not None and wr is not None:
            traffic[r["report"].replace(".ncu-rep", "")] = {"kernel": r["kernel"], "dram_bytes": rd * scale(ru) + wr * scale(wu)}
    json.dump(traffic, open(os.path.join(ROOT, "profiles", f"{tag}_traffic.json"), "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r01")
