"""CPU: the JSON line of `bench.py --impl reference` (the only arm that runs without a GPU) carries the keys
the driver reads, and the committed round profiles of our arm do too."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def test_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1", "--config", "cfg0"], capture_output=True, text=True, timeout=600,
                       env={k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")})
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1                                    # exactly one line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and BASE_KEYS <= set(d)
    assert d["metric"] == "lafs_ssl_hot_path_faces_per_sec" and d["unit"] == "faces/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and "workload" in d["config"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["steps"] == 1 and d["warmup"] == 1               # the arm honours --steps / --warmup
    if os.path.isfile("/root/reference/lafs_train.py") or os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "lafs_train.py")):
        assert cb["kind"] == "reference"                      # the unmodified reference modules, not the port


def test_committed_profiles_of_our_arm_follow_the_contract():
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_bench_n*.json")))
    assert files
    for f in files:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        assert (BASE_KEYS - {"cpu_baseline"}) <= set(d), (f, BASE_KEYS - set(d))
        assert d["metric"] == "lafs_ssl_hot_path_faces_per_sec" and d["scaling"] == "weak" and d["data"] == "synthetic"
        assert d["n_gpus"] >= 1 and d["warmup"] >= 3 and d["gpu_launches"] > 0
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
        assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["value"] != d["value"]
        rf = d["roofline"]
        assert rf["bound"] in ("hbm", "tensor") and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-3
        assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
        assert not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(d["clocks"]["reasons"]))
        if d["n_gpus"] == 1:
            assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])


def test_round2_profiles_carry_the_parity_and_fused_head_evidence():
    """what DESIGN.md quotes from profiles/: sharded-vs-unsharded head parity at every N > 1, and the fused
    DINOHead-tail + loss leg next to its unfused form at N = 1."""
    for n in (2, 4, 8):
        d = json.loads(open(os.path.join(ROOT, "profiles", f"r02_bench_n{n}.json")).read().strip().splitlines()[-1])
        assert d["n_gpus"] == n
        for name, h in d["head"].items():
            assert h["shards"] == n and 0 <= h["parity_max_rel"] <= 2e-3, (n, name, h.get("parity_max_rel"))
    d = json.loads(open(os.path.join(ROOT, "profiles", "r02_bench_n1.json")).read().strip().splitlines()[-1])
    leg = d["extras"]["dino_head_fused(last_layer+loss)"]
    assert leg["out_dim"] == 65536 and leg["B"] == 256 and leg["fused"]["logit_bytes_in_hbm"] == 0
    assert leg["fused"]["ms_fwd_bwd"] < leg["unfused_same_gpu"]["ms_fwd_bwd"]
    assert leg["loss_rel_diff_fused_vs_unfused"] <= 1e-3


def test_fused_head_cpu_reference_leg_runs_the_reference():
    """bench.py's host-core baseline for the fused DINO-head leg: the reference's DINOHead tail + DINOLoss (small size
    here); it must report instead of raising."""
    sys.path.insert(0, ROOT)
    import bench
    out = bench.dino_head_cpu_reference(4, 4, 1024, 64)
    if os.path.isfile("/root/reference/lafs_train.py") or os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "lafs_train.py")):
        cb = out["cpu_reference"]
        assert cb["kind"] == "reference" and cb["ms_fwd_bwd"] > 0 and cb["cores"] >= 1 and 0 < cb["loss"] < 20
    else:
        assert out == {}
