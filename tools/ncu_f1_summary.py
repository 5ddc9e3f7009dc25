"""Sums an ncu --csv launch list (gpu__time_duration, dram bytes read/written) per kernel and in total.
    python tools/ncu_f1_summary.py fused.csv unfused.csv   -> markdown on stdout"""
import csv
import sys
from collections import OrderedDict


def load(path):
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    for d in csv.DictReader(lines):
        rows.append(d)
    per = OrderedDict()
    for d in rows:
        key = (d["ID"], d["Kernel Name"][:70])
        rec = per.setdefault(key, {"us": 0.0, "rd": 0.0, "wr": 0.0})
        val = float(d["Metric Value"].replace(",", ""))
        unit = d["Metric Unit"]
        name = d["Metric Name"]
        if name == "gpu__time_duration.sum":
            rec["us"] = val / 1e3 if unit in ("ns", "nsecond") else val * ({"us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3}.get(unit, 1))
        else:
            mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
            rec["rd" if "read" in name else "wr"] = val * mul
    return per


if __name__ == "__main__":
    for path in sys.argv[1:]:
        per = load(path)
        # skip the input-generation kernels of the driver script (before the first synchronize): torch RNG / fill kernels
        items = [(k, v) for k, v in per.items()
                 if not any(t in k[1] for t in ("distribution_", "philox", "FillFunctor", "normal_kernel"))]
        print("\n### %s" % path)
        print("| # | kernel | us | DRAM read MB | DRAM write MB |")
        print("|---|---|---|---|---|")
        tot = {"us": 0.0, "rd": 0.0, "wr": 0.0}
        for (i, name), v in items:
            print("| %s | %s | %.1f | %.1f | %.1f |" % (i, name, v["us"], v["rd"] / 1e6, v["wr"] / 1e6))
            for k in tot:
                tot[k] += v[k]
        print("| | **total (%d launches)** | **%.1f** | **%.1f** | **%.1f** |" % (len(items), tot["us"], tot["rd"] / 1e6, tot["wr"] / 1e6))
