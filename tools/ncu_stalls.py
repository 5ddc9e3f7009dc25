"""Top warp-stall sites per captured kernel (ncu --page source) -> profiles/<tag>_ncu_stalls.md.
Run in the build container on gpurun_out/*.ncu-rep (captured with --import-source on)."""
import csv
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(tag, top=8):
    out = [f"# Warp-stall hot spots per kernel ({tag}; `ncu --set full --import-source on`, `--page source`)\n",
           "Share of all warp-state samples of the kernel attributed to one SASS instruction; `ex` = times the\n"
           "instruction was executed (warp level).  Spin loops on mbarriers show up as `SYNCS.PHASECHK...TRYWAIT` /\n"
           "`BRA` pairs: a high share on the epilogue's accumulator wait means the epilogue is starved, not slow.\n"]
    for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "*.ncu-rep"))):
        r = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True)
        rows = list(csv.reader(r.stdout.splitlines()))
        if len(rows) < 3:
            continue
        kernel = rows[0][1] if len(rows[0]) > 1 else ""
        hdr = rows[1]
        try:
            isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
        except ValueError:
            continue
        data = [x for x in rows[2:] if len(x) > isamp and x[isamp].strip().isdigit()]
        tot = sum(int(x[isamp]) for x in data) or 1
        out.append(f"\n## {os.path.basename(rep)} — `{kernel[:110]}`\n\n| share | ex | SASS |\n|---|---|---|\n")
        for x in sorted(data, key=lambda x: -int(x[isamp]))[:top]:
            sass = " ".join(x[isrc].split())[:90].replace("|", "\\|")
            out.append(f"| {100 * int(x[isamp]) / tot:.1f} % | {x[iex]} | `{sass}` |\n")
    path = os.path.join(ROOT, "profiles", f"{tag}_ncu_stalls.md")
    open(path, "w").write("".join(out))
    print(path)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r01")
