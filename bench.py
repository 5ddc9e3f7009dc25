#!/usr/bin/env python
"""LAFS hot-path benchmark (BASELINE.json metric: hot-path faces/sec, % of HBM/TC roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...     (N > 1)

Workload at every N: BASELINE.json configs[1] -- "LAFS SSL pretrain step, Part-fViT ViT-B
(196 landmark patches) + DINO head out_dim 65536, batch 256 synthetic faces" PER GPU (weak
scaling; the only data-path collective is the reference's all-reduce of the DINO centre).
One step = one pass of the hot path over one batch (lafs_cvpr2024_b200/ssl_step.py):
landmark tail + fused patch gather -> patch_to_embedding (student+teacher on the 2 global views,
student on the 4 local views), DINO loss forward+backward+centre
update on [6*256, 65536] / [2*256, 65536] bf16 logits, and the teacher EMA over the 147
ViT-B + DINO-head parameter tensors.  The transformer blocks, DINO head MLP and landmark-CNN
trunk are outside the path and are not run.

`value`  : faces/s with every input resident in HBM when the timed region starts.
`e2e`    : the same step through the public Python API with the per-step HOST inputs (the
           data loader's images, CPU-generator noise and indices) copied from pinned memory
           inside the timed region and the loss read back.  The logits and the parameters are
           device-resident in the real system (they are produced/owned by the device-side
           model), so they are not copied.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_PER_GPU = 256
N_LOCAL = 4
OUT_DIM = 65536
N_LAND = 196
KEEP_LOCAL = 36
DIM = 768
METRIC = "lafs_ssl_hot_path_faces_per_sec"

# --config: the default is BASELINE.json configs[1] (the configuration the metric is quoted on); the others are
# the remaining SSL configurations of BASELINE.json / SURVEY 8(d), run by hand and committed under profiles/
WORKLOADS = {
    "cfg1": {"B": 256, "L": 4, "vit": "B", "name": "BASELINE configs[1]: LAFS SSL pretrain hot path, ViT-B, 196 landmark "
             "patches, out_dim 65536, 2 global + 4 local crops, batch 256 per GPU"},
    "cfg1_L8": {"B": 256, "L": 8, "vit": "B", "name": "configs[1] with the reference's default of 8 local crops "
                "(lafs_train.py:101), batch 256 per GPU"},
    "cfg4": {"B": 128, "L": 4, "vit": "B", "name": "BASELINE configs[4]: SSL hot path with teacher EMA + landmark "
             "augmentations, batch 128 per GPU (1024 over 8 GPUs)"},
    "cfg0": {"B": 8, "L": 4, "vit": "S", "name": "BASELINE configs[0]: Part-fViT ViT-S (dim 384), batch 8, landmark patch "
             "sampling + DINOLoss (2 global + 4 local crops)"},
}


def vit_param_shapes(vit="B", out_dim=OUT_DIM):
    """The parameter tensors EMA'd every step: Part-fViT backbone (incl. the unused CosFace weight
    [30000,dim]) + DINOHead.  ViT-B (dim 768, 11 heads x 64, mlp 2048: ViT_face.py / lafs_train.py:
    302-333): 147 tensors, 110,261,248 parameters.  ViT-S (dim 384, 6 heads x 64, mlp 1536)."""
    dim, inner, mlp = (768, 704, 2048) if vit == "B" else (384, 384, 1536)
    shapes = [(1, 197, dim), (dim, 192), (dim,), (1, 1, dim)]
    for _ in range(12):
        shapes += [(dim,), (dim,), (3 * inner, dim), (dim, inner), (dim,), (dim,), (dim,), (mlp, dim), (mlp,),
                   (dim, mlp), (dim,)]
    shapes += [(dim,), (dim,), (30000, dim)]
    shapes += [(2048, dim), (2048,), (2048, 2048), (2048,), (256, 2048), (256,), (out_dim, 1), (out_dim, 256)]
    return shapes


def vit_b_param_shapes(out_dim=OUT_DIM):
    return vit_param_shapes("B", out_dim)


def set_workload(name):
    global B_PER_GPU, N_LOCAL, DIM, VIT, WL_NAME, WL_KEY
    WL_KEY = name
    w = WORKLOADS[name]
    B_PER_GPU, N_LOCAL, VIT, WL_NAME = w["B"], w["L"], w["vit"], w["name"]
    DIM = 768 if VIT == "B" else 384


VIT, WL_NAME, WL_KEY = "B", WORKLOADS["cfg1"]["name"], "cfg1"
_ORIG_AFFINITY = os.sched_getaffinity(0)


def ncu_traffic(report):
    """DRAM bytes per launch of a kernel from the committed ncu capture (profiles/*_traffic.json)."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")))
    if not files:
        return None
    try:
        return float(json.load(open(files[-1]))[report]["dram_bytes"])
    except Exception:
        return None


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return {"hbm": float(p["hbm_gbs"]), "tc": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                "src": "measured (MEASURED_PEAKS.json)"}
    except Exception:
        return {"hbm": 6650.0, "tc": 1590.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.th = index, [], False, None

    def _run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.05)

    def start(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def stop(self):
        self.stop_flag = True
        if self.th:
            self.th.join(timeout=10)
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(local):
    """Pins this process (and hence the first-touch placement of the pinned host buffers allocated next) to
    the CPUs of the NUMA node the GPU hangs off: with several ranks per box the H2D copies then read local
    memory instead of crossing the socket interconnect.  Returns a short description for the JSON line."""
    try:
        pr = torch.cuda.get_device_properties(local)
        addr = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{addr}/numa_node").read().strip())
        if node < 0:
            return f"gpu {addr}: no NUMA node reported"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return f"gpu {addr}: node {node} has no CPU this process may use"
        os.sched_setaffinity(0, cpus)
        return f"gpu {addr} -> NUMA node {node} ({len(cpus)} cpus)"
    except Exception as e:
        return "unbound (%s)" % str(e).splitlines()[0][:60]


def make_host_inputs(B, seed, uint8):
    """Per-step host-side inputs: the data loader's augmented views and the CPU-generator noise /
    indices (ViT_face.py:1361,1366).  Pinned.
    uint8=True : decoded pixels; ToTensor + Normalize((0.5,)*3, (0.5,)*3) (lafs_train.py:800-803)
                 run inside the gather kernel -- 1 byte/pixel over PCIe and from HBM.
    uint8=False: the fp32 tensors normalised to [-1,1] that the reference's loader emits."""
    g = torch.Generator().manual_seed(seed)
    L = N_LOCAL
    u8g = torch.randint(0, 256, (2 * B, 3, 112, 112), generator=g, dtype=torch.uint8)
    u8l = torch.randint(0, 256, (L * B, 3, 112, 112), generator=g, dtype=torch.uint8)
    h = {
        "img_g": u8g if uint8 else (u8g.float() / 255 - 0.5) / 0.5,
        "img_l": u8l if uint8 else (u8l.float() / 255 - 0.5) / 0.5,
        "noise_g": torch.randn(2 * B, N_LAND, 2, generator=g) * 5,
        "noise_l": torch.randn(L * B, N_LAND, 2, generator=g) * 5,
        "idx_l": torch.randint(0, N_LAND, (L * B, KEEP_LOCAL), generator=g),
    }
    return {k: v.pin_memory() for k, v in h.items()}


def make_device_state(B, seed, dev):
    g = torch.Generator(device=dev).manual_seed(seed)
    L = N_LOCAL
    st = {
        # outputs of the (out-of-path) landmark CNN / student / teacher networks
        "raw_g": torch.randn(2 * B, 2 * N_LAND, device=dev, generator=g),
        "raw_l": torch.randn(L * B, 2 * N_LAND, device=dev, generator=g),
        "student_out": torch.randn((L + 2) * B, OUT_DIM, device=dev, generator=g).bfloat16(),
        "teacher_out": torch.randn(2 * B, OUT_DIM, device=dev, generator=g).bfloat16(),
        # gradients of the student's embedded tokens, as the (out-of-path) transformer backward would deliver them
        "grad_s_g": (torch.randn(2 * B, N_LAND, DIM, device=dev, generator=g) * 0.01).bfloat16(),
        "grad_s_l": (torch.randn(L * B, KEEP_LOCAL, DIM, device=dev, generator=g) * 0.01).bfloat16(),
    }
    shapes = vit_param_shapes(VIT)
    st["student_params"] = [torch.randn(*s, device=dev, generator=g) * 0.02 for s in shapes]
    st["teacher_params"] = [p.clone() for p in st["student_params"]]
    return st


def run_ours(args):
    import torch.distributed as dist
    import lafs_cvpr2024_b200 as P
    from lafs_cvpr2024_b200 import _lib
    from lafs_cvpr2024_b200.ssl_step import GraphedSSLStep, SSLHotPath
    torch.backends.cuda.matmul.allow_tf32 = False

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch N>1 with torch.distributed.run")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # rank 0 prints exactly ONE line on stdout: keep NCCL's "NCCL version ..." banner off it
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    if _lib.lib().lafs_device_ok() != 1:
        raise SystemExit("bench.py needs a compute-capability 10.x device (B200)")

    numa = bind_to_gpu_numa_node(local)
    B, L = B_PER_GPU, N_LOCAL
    host = make_host_inputs(B, 1000 + rank, uint8=True)
    host_f32 = make_host_inputs(B, 1000 + rank, uint8=False)
    st = make_device_state(B, 2000 + rank, dev)
    dev_in = {k: v.to(dev) for k, v in host.items()}
    dev_in_f32 = {k: v.to(dev) for k, v in host_f32.items()}
    # parameter list order: [pos_embedding, patch_to_embedding.weight, patch_to_embedding.bias, ...]
    s_embed = (st["student_params"][1], st["student_params"][2])
    t_embed = (st["teacher_params"][1], st["teacher_params"][2])
    path = SSLHotPath(OUT_DIM, L, st["teacher_params"], st["student_params"], student_embed=s_embed, teacher_embed=t_embed)
    path.loss.center = torch.randn(1, OUT_DIM, device=dev) * 0.1
    centre_exchange = "none (single process)"
    if world > 1:
        # the path's one collective (teacher column sums, lafs_train.py:675) through the peer-memory kernel;
        # NCCL all_reduce if symmetric memory cannot be set up on this box
        try:
            path.loss.enable_peer_exchange()
            path.loss._exchange(OUT_DIM, dev)
            centre_exchange = "peer-memory all-reduce kernel (csrc/exchange.cu)"
        except Exception as e:
            path.loss.enable_peer_exchange(enabled=False)
            centre_exchange = "nccl all_reduce (peer exchange unavailable: %s)" % str(e).splitlines()[0][:80]
    sched = 0.996 + 0.5 * (1 - 0.996) * (1 - np.cos(np.pi * np.arange(100000) / 100000))  # utils.py:187-198
    names = ["landmark+gather_embed", "patch_embed_bwd(student)", "dino_fwd+center", "dino_bwd", "ema"]

    def step(inp, it, evs=None):
        def mark(i):
            if evs is not None:
                evs[i].record()
        mark(0)
        path.landmarks_and_embeddings(st["raw_g"], inp["noise_g"], inp["img_g"], st["raw_l"], inp["noise_l"],
                                      inp["idx_l"], inp["img_l"], keep_tokens=True)
        mark(1)
        path.student_embed_backward(st["grad_s_g"], st["grad_s_l"])
        mark(2)
        s = st["student_out"].requires_grad_(True)
        s.grad = None
        loss = path.loss(s, st["teacher_out"], it % 41)
        mark(3)
        loss.backward()
        mark(4)
        path.ema_step(sched[it])
        mark(5)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, per_kernel=False):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(6)] for _ in range(steps)] if per_kernel else None
        barrier()
        e0.record()
        for i in range(steps):
            fn(i, evs[i] if per_kernel else None)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        parts = None
        if per_kernel:
            parts = [float(np.mean([evs[i][j].elapsed_time(evs[i][j + 1]) for i in range(steps)])) for j in range(5)]
        return float(t.item()), parts

    # ---- device-resident arm --------------------------------------------------------------
    for i in range(args.warmup):
        step(dev_in, i)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_total, parts = timed(lambda i, ev: step(dev_in, args.warmup + i, ev), args.steps, per_kernel=True)
    # the same 11-kernel step captured into ONE CUDA graph and replayed (static device buffers):
    # this is the device-resident headline `value`; the eager run above gives the per-kernel split
    ginp = dict(dev_in, raw_g=st["raw_g"], raw_l=st["raw_l"], student_out=st["student_out"], teacher_out=st["teacher_out"],
                grad_s_g=st["grad_s_g"], grad_s_l=st["grad_s_l"])
    # teacher EMA and the patch-embed backward GEMMs on two captured side streams next to the DINO kernels (measured
    # on B200, tools/graph_probe.py: 0.6084 -> 0.5863 ms; the kernels on the main stream leave ~half of the HBM
    # bandwidth idle, the EMA fills it)
    graphed = GraphedSSLStep(path, ginp, epoch=3, momentum=float(sched[1000]), overlap_ema="late2", ema_ctas=0)
    for _ in range(3):
        graphed.replay()
    ms_graph, _ = timed(lambda i, ev: graphed.replay(), args.steps)
    path.loss.center = graphed.center.clone()
    del graphed
    # same step fed with the reference's fp32 image tensors (strict drop-in input format)
    for i in range(3):
        step(dev_in_f32, i)
    ms_total_f32, parts_f32 = timed(lambda i, ev: step(dev_in_f32, args.warmup + i, ev), args.steps, per_kernel=True)
    del dev_in_f32

    # ---- end-to-end arm: every step's HOST inputs are copied from pinned memory inside the timed
    # region (copy stream, double-buffered, like the reference's pin_memory + non_blocking loader
    # hand-off, lafs_train.py:521) and the loss is read back to the host every step ------------------
    def run_e2e(hbuf):
        """Per step: the HOST inputs are copied from pinned memory on a copy stream into one of two static device
        input sets (double buffered, like the reference's pin_memory + non_blocking loader hand-off,
        lafs_train.py:521), the step's CUDA graph over that set is replayed (GraphedSSLStep: the public call), and
        the loss is copied back to pinned host memory; the host inspects the loss of step i-2 (the reference's
        isfinite check, lafs_train.py:585, lagging two steps) so that it never drains the pipeline."""
        h2d = sum(v.numel() * v.element_size() for v in hbuf.values())
        shared = dict(raw_g=st["raw_g"], raw_l=st["raw_l"], student_out=st["student_out"], teacher_out=st["teacher_out"],
                      grad_s_g=st["grad_s_g"], grad_s_l=st["grad_s_l"])
        bufs = [dict({k: v.to(dev) for k, v in hbuf.items()}, **shared) for _ in range(2)]
        center = path.loss.center.detach().clone().contiguous()
        graphs = [GraphedSSLStep(path, bufs[b], epoch=3, momentum=float(sched[1000]), center=center, overlap_ema="late2",
                                 ema_ctas=0) for b in range(2)]
        copy_stream = torch.cuda.Stream(device=dev)
        ready = [torch.cuda.Event() for _ in range(2)]
        free = [torch.cuda.Event() for _ in range(2)]
        ring = 4
        loss_host = torch.zeros(ring, dtype=torch.float32).pin_memory()
        loss_ev = [torch.cuda.Event() for _ in range(ring)]

        def upload(i):
            b = i & 1
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(free[b])
                for k, v in hbuf.items():
                    bufs[b][k].copy_(v, non_blocking=True)
                ready[b].record(copy_stream)

        def loop(nsteps):
            main = torch.cuda.current_stream()
            for b in range(2):
                free[b].record()
            upload(0)
            for i in range(nsteps):
                if i + 1 < nsteps:
                    upload(i + 1)
                main.wait_event(ready[i & 1])
                loss = graphs[i & 1].replay()
                free[i & 1].record()
                loss_host[i % ring].copy_(loss, non_blocking=True)      # device -> host read of the step's result
                loss_ev[i % ring].record()
                if i >= 2:
                    loss_ev[(i - 2) % ring].synchronize()
                    if not np.isfinite(float(loss_host[(i - 2) % ring])):
                        raise SystemExit("bench.py: non-finite loss in the end-to-end loop")
            torch.cuda.synchronize()
            return float(loss_host[(nsteps - 1) % ring])

        loop(3)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        loop(args.steps)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        path.loss.center = center.clone()
        del graphs, bufs
        torch.cuda.empty_cache()
        return float(t.item()), h2d

    ms_e2e, h2d = run_e2e(host)
    ms_e2e_f32, h2d_f32 = run_e2e(host_f32)
    clocks = sampler.stop() if sampler else None    # sampled across every timed region above
    xc = getattr(path.loss, "_xchg", None)
    if xc is not None:
        xc.check()                    # raises if a rank missed the centre exchange's spin bound

    # ---- secondary: clip + AdamW + teacher EMA as one multi-tensor update ------------------
    extras = {}
    try:
        upd = P.StudentUpdate(st["student_params"], st["teacher_params"], regularized=[p.dim() > 1 for p in st["student_params"]])
        ugr = [torch.randn_like(p) * 0.01 for p in st["student_params"]]
        for _ in range(3):
            upd.step(ugr, 5e-4, 0.04, 3.0, 0.996)
        ms_upd, _ = timed(lambda i, ev: upd.step(ugr, 5e-4, 0.04, 3.0, 0.996), max(5, min(args.steps, 20)))
        ms_upd /= max(5, min(args.steps, 20))
        npar = sum(p.numel() for p in st["student_params"])
        extras["student_update(clip+adamw+ema)"] = {
            "ms": round(ms_upd, 5), "alg_bytes": 40 * npar, "GBps": round(40 * npar / ms_upd / 1e6, 1),
            "note": "per-tensor clip (read g) + AdamW + EMA fold (read p,g,m,v,k; write p,m,v,k): 40 B/param; "
                    "replaces utils.clip_gradients + AdamW.step + the EMA loop (utils.py:132-141, lafs_train.py:601-613)"}
        del upd, ugr
    except Exception as e:
        extras["student_update_error"] = str(e).splitlines()[0][:160]
    torch.cuda.empty_cache()

    # ---- secondary: class-sharded margin head (BASELINE configs[2], configs[3]) -----------------
    del st, path, dev_in
    torch.cuda.empty_cache()
    head = bench_head(P, world, rank, dev, dist, args)
    # ---- secondary: DINOHead.last_layer fused with the DINO loss (per-GPU work, no collective
    # beyond the centre all-reduce already measured above: reported at N = 1) -------------------------------------
    if world == 1:
        try:
            extras["dino_head_fused(last_layer+loss)"] = bench_dino_head(dev, iters=max(5, min(args.steps, 20)),
                                                                         cpu_ref=not args.no_cpu)
        except Exception as e:
            extras["dino_head_fused_error"] = str(e).splitlines()[0][:160]
            torch.cuda.synchronize()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    ms_step = ms_graph / args.steps
    ms_step_eager = ms_total / args.steps
    faces = B * world
    K, nc = OUT_DIM, L + 2
    nparam = sum(int(np.prod(s)) for s in vit_param_shapes(VIT))
    tok_out = (2 * 2 * B * 196 + L * B * 36) * DIM * 2
    alg = {
        # images + landmarks in, bf16 tokens of both models out (e_img = 1 byte)
        "landmark+gather_embed": {"bytes": (2 + L) * B * (3 * 112 * 112 * 1) + 2 * B * 196 * 8 + L * B * 36 * 8 + tok_out
                                  + (2 * B * 196 + L * B * 36) * 192 * 2,       # + the bf16 tokens kept for the backward
                                  "flops": 2.0 * 192 * DIM * (2 * 2 * B * 196 + L * B * 36), "write_bytes": tok_out},
        # student weight/bias gradient: read dY (bf16) and the kept bf16 tokens once, 2*M*193*dim flops
        "patch_embed_bwd(student)": {"bytes": (2 * B * 196 + L * B * 36) * (DIM * 2 + 192 * 2) + DIM * 193 * 4,
                                     "flops": 2.0 * 193 * DIM * (2 * B * 196 + L * B * 36)},
        "dino_fwd+center": {"bytes": (nc + 2) * B * K * 2 + 8 * K},
        "dino_bwd": {"bytes": (2 * nc + 2) * B * K * 2},
        "ema": {"bytes": 12 * nparam},
    }
    kernels = {}
    for n, ms in zip(names, parts):
        gbps = alg[n]["bytes"] / ms / 1e6
        kernels[n] = {"ms": round(ms, 5), "alg_bytes": alg[n]["bytes"], "GBps": round(gbps, 1),
                      "frac_hbm": round(gbps / pk["hbm"], 4)}
        if "flops" in alg[n]:
            kernels[n]["TFLOPs"] = round(alg[n]["flops"] / ms / 1e9, 1)
            kernels[n]["frac_tc"] = round(alg[n]["flops"] / ms / 1e9 / pk["tc"], 4)
    # the fp32-image variant of the first stage (e_img = 4 bytes), for the strict drop-in input format
    b32 = alg["landmark+gather_embed"]["bytes"] + (2 + L) * B * 3 * 112 * 112 * 3
    kernels["landmark+gather_embed(fp32 images)"] = {
        "ms": round(parts_f32[0], 5), "alg_bytes": b32, "GBps": round(b32 / parts_f32[0] / 1e6, 1),
        "frac_hbm": round(b32 / parts_f32[0] / 1e6 / pk["hbm"], 4)}
    kernels["landmark+gather_embed"]["note"] = ("bound = max(HBM, TC): %.0f MB of bf16 tokens out; pure-write streams on this part "
                                                "measure 6.2-6.9 TB/s (tools/wbw_probe.cu, profiles/r02_wbw_probe.txt)" % (tok_out / 1e6))
    # the DINO loss as a whole (north star: ">= 70 % HBM roofline on the DINO loss"): forward + centre + backward
    dms = kernels["dino_fwd+center"]["ms"] + kernels["dino_bwd"]["ms"]
    dby = alg["dino_fwd+center"]["bytes"] + alg["dino_bwd"]["bytes"]
    kernels["dino_loss(fwd+bwd)"] = {"ms": round(dms, 5), "alg_bytes": dby, "GBps": round(dby / dms / 1e6, 1),
                                     "frac_hbm": round(dby / dms / 1e6 / pk["hbm"], 4)}
    dom = max(names, key=lambda n: kernels[n]["ms"])
    line = {
        "metric": METRIC, "value": round(faces / (ms_step / 1e3), 1), "unit": "faces/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 5),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16 logits+tokens / uint8 images / fp32 params",
        "data": "synthetic",
        "config": {"workload": WL_NAME,
                   "batch_per_gpu": B, "out_dim": K, "ncrops": nc, "ema_params": nparam, "ema_tensors": len(vit_param_shapes(VIT)),
                   "images": "uint8 decoded pixels, normalised in-kernel (value_fp32_images / e2e_fp32_images: fp32 tensors)",
                   "l2": "inputs larger than L2 (logits 335 MB, tokens out 362 MB, parameters 882 MB per step)",
                   "parallelism": f"dp{world}", "centre_exchange": centre_exchange},
        "e2e": {"value": round(faces / (ms_e2e / args.steps / 1e3), 1), "unit": "faces/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": round(ms_e2e / args.steps, 5),
                "transport": "uint8 pixels + fp32 noise + int64 indices from pinned memory on a copy stream into two static "
                             "input sets; one CUDA-graph replay per step (GraphedSSLStep); loss copied to pinned host memory "
                             "every step and checked two steps later; ToTensor/Normalize fused into the gather kernel",
                "host_numa": numa},
        "e2e_fp32_images": {"value": round(faces / (ms_e2e_f32 / args.steps / 1e3), 1), "unit": "faces/s",
                            "h2d_bytes_per_step": h2d_f32, "d2h_bytes_per_step": 4, "ms_per_step": round(ms_e2e_f32 / args.steps, 5),
                            "transport": "the reference's fp32 normalised image tensors (PCIe-bound)"},
        "value_fp32_images": round(faces / (ms_total_f32 / args.steps / 1e3), 1),
        "value_eager_launches": round(faces / (ms_step_eager / 1e3), 1),
        "launch": "value: the step's 15 kernels replayed as one CUDA graph with three captured streams (DINO kernels | teacher EMA | patch-embed backward) (lafs_cvpr2024_b200.ssl_step.GraphedSSLStep); "
                  "value_eager_launches / kernels / e2e: the same kernels launched one by one from Python",
        # 2 landmark, 3 weight prep, 2 gather-embed, 2 x (dW GEMM + reduce/un-permute), 2 dino fwd (+centre), 1 dino bwd, 1 ema
        "gpu_launches": 15,
        "clocks": clocks,
        "roofline": {"kernel": dom, "bound": "hbm", "achieved": kernels[dom]["GBps"], "peak": pk["hbm"], "unit": "GB/s",
                     "frac": kernels[dom]["frac_hbm"], "alg_bytes": kernels[dom]["alg_bytes"],
                     "traffic": ncu_traffic({"ema": "ema_bench", "dino_bwd": "dino_bwd", "dino_fwd+center": "dino_fwd",
                                             "landmark+gather_embed": "pe_global_u8"}.get(dom, dom)),
                     "peak_source": pk["src"]},
        "kernels": kernels,
    }
    for name, h in head.items():
        h["frac_tc"] = round(h["TFLOPs_6BCD"] / pk["tc"], 4)
    line["head"] = head
    for k, v in extras.items():
        if isinstance(v, dict) and "GBps" in v:
            v["frac_hbm"] = round(v["GBps"] / pk["hbm"], 4)
    line["extras"] = extras
    if world == 1 and not args.no_ref_gpu and _reference_available():
        # the unmodified reference modules in eager PyTorch on this same GPU (N = 1 report)
        line["reference_eager_b200"] = reference_eager_gpu(dev)
        r = line["reference_eager_b200"].get("ssl_step")
        if r:
            line["reference_eager_b200"]["speedup_value_over_reference_eager"] = round(r["ms_per_step"] / ms_step, 1)
    if not args.no_cpu and world == 1:          # the CPU baseline is a rank-0, N = 1 report
        try:
            line["cpu_baseline"] = cpu_baseline()
        except Exception as e:                  # never lose the measured line to the baseline leg
            line["cpu_baseline"] = {"value": None, "unit": "faces/s", "cores": os.cpu_count(), "kind": "reference",
                                    "sample": "failed: " + str(e).splitlines()[0][:120]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def bench_head(P, world, rank, dev, dist, args):
    """Finetune margin head, classes sharded over the N ranks with torch.chunk's rule; every rank
    sees the global batch (post all-gather embeddings).  One step = weight/embedding
    normalisation + fused GEMM/softmax-CE forward + statistics exchange + recompute-G backward
    (dE summed over shards, dW local).  Strong scaling in N (fixed global batch).  With N > 1 the two
    exchange steps are timed both through NCCL and through the peer-memory kernels (csrc/exchange.cu)."""
    iters = max(5, min(args.steps, 20))

    def time_it(fn):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def run_variant(cls, B, C, D, peer):
        torch.manual_seed(7)
        h = cls(D, C, None, shard=(rank, world) if world > 1 else None).to(dev)
        if peer:
            h.enable_peer_exchange()
        x = torch.randn(B, D, device=dev, requires_grad=True)
        lab = torch.randint(0, C, (B,), device=dev)

        def step():
            x.grad = None
            h.weight.grad = None
            loss = h.forward_loss(x, lab)
            loss.backward()
            return loss

        for _ in range(3):
            step()
        ms_eager = time_it(step)
        # the same step (kernels + exchanges) as one CUDA graph: at 8 shards the per-rank GEMMs are a few
        # tens of microseconds and launch latency dominates the eager number
        ms, mode = ms_eager, "eager"
        try:
            graph = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            with torch.cuda.graph(graph):
                step()
            for _ in range(3):
                graph.replay()
            ms = time_it(graph.replay)
            mode = "cuda_graph"
            del graph
        except Exception as e:  # capture not possible on this stack: keep the eager number
            mode = "eager (graph capture failed: %s)" % str(e).splitlines()[0][:80]
            torch.cuda.synchronize()
        if peer:                      # a peer that missed the kernels' spin bound would have produced garbage
            for xc in h._xchg.values():
                xc.check()
        # ---- N > 1: the sharded step against the unsharded one on the same inputs (every rank checks its
        # own class slice; the worst relative error over the ranks is reported and gates the bench) ----------
        parity = None
        if world > 1:
            loss_s = step().detach().clone()
            de_s, dw_s = x.grad.detach().clone(), h.weight.grad.detach().clone()
            torch.manual_seed(7)
            hu = cls(D, C, None).to(dev)
            xu = x.detach().clone().requires_grad_(True)
            loss_u = hu.forward_loss(xu, lab)
            loss_u.backward()
            lo, hi = h.class_lo, h.class_hi
            dw_u = hu.weight.grad[lo:hi]
            errs = torch.stack([
                (loss_s - loss_u.detach()).abs() / loss_u.detach().abs(),
                (de_s - xu.grad).abs().max() / xu.grad.abs().max(),
                (dw_s - dw_u).abs().max() / hu.weight.grad.abs().max()]).double()
            dist.all_reduce(errs, op=dist.ReduceOp.MAX)
            parity = [float(v) for v in errs.tolist()]
            del hu, xu, dw_u
        del h, x
        torch.cuda.empty_cache()
        return ms, ms_eager, mode, parity

    import lafs_cvpr2024_b200 as PK
    out = {}
    for name, B, C, D in HEAD_CFGS:
        cls = PK.CosFace if name.startswith("cosface") else PK.ArcFace
        res = {}
        for variant in (["nccl", "peer"] if world > 1 else ["single"]):
            try:
                res[variant] = run_variant(cls, B, C, D, variant == "peer")
            except Exception as e:            # e.g. symmetric memory unavailable on this box: NCCL number stands
                res[variant + "_error"] = str(e).splitlines()[0][:120]
                torch.cuda.synchronize()
        timed = {k: v for k, v in res.items() if isinstance(v, tuple)}
        best = min(timed, key=lambda k: timed[k][0])
        ms, ms_eager, mode, _ = timed[best]
        out[name] = {"B_global": B, "classes": C, "D": D, "shards": world, "ms_fwd_bwd": round(ms, 4), "launch": mode,
                     "ms_fwd_bwd_eager": round(ms_eager, 4), "exchange": best,
                     "faces_per_s": round(B / ms * 1e3, 1), "TFLOPs_6BCD": round(6.0 * B * C * D / ms / 1e9 / world, 1),
                     "note": "TFLOPs per GPU on the 6*B*C*D/R count; the step also recomputes the logits once (8*B*C*D issued)"}
        if world > 1:
            out[name]["ms_by_exchange"] = {k: round(v[0], 4) for k, v in timed.items()}
            out[name].update({k: v for k, v in res.items() if k.endswith("_error")})
            # sharded vs unsharded on the same inputs: [loss rel, dE max-norm rel, dW-slice max-norm rel], worst rank
            pm = {k: v[3] for k, v in timed.items() if v[3] is not None}
            out[name]["parity_vs_unsharded"] = {k: [float("%.3g" % e) for e in v] for k, v in pm.items()}
            out[name]["parity_max_rel"] = float("%.3g" % max(max(v) for v in pm.values()))
            if out[name]["parity_max_rel"] > HEAD_PARITY_TOL:
                raise SystemExit(f"bench.py: sharded head {name} deviates from the unsharded head by "
                                 f"{out[name]['parity_max_rel']} (> {HEAD_PARITY_TOL}) at world={world}")
    return out


def bench_dino_head(dev, iters=20, B=None, ncrops=None, K=None, D=256, fused_only=False, cpu_ref=False):
    """DINOHead.last_layer (weight-normed 256 -> 65536) fused with the DINO loss
    (lafs_cvpr2024_b200/dino_head.py) against the unfused form of the same work on the same GPU: cuBLAS bf16
    last-layer GEMMs that write the [(ncrops+2)B, K] logits + this repo's DINO loss kernels on those logits + autograd
    (cuBLAS dX / dW, weight-norm and normalize backward).  Both from the bottleneck features to
    (loss, d features, d weight_v, new centre); CUDA-graph replays, device events."""
    import lafs_cvpr2024_b200 as P
    B = B or B_PER_GPU
    ncrops = ncrops or (N_LOCAL + 2)
    K = K or OUT_DIM
    torch.manual_seed(11)
    xs = torch.randn(ncrops * B, D, device=dev)
    xt = torch.randn(2 * B, D, device=dev)
    vs = torch.randn(K, D, device=dev) * 0.02
    vt = vs + torch.randn(K, D, device=dev) * 0.002
    one = torch.ones(K, device=dev)
    center = torch.randn(K, device=dev) * 0.05
    gout = torch.ones((), device=dev)
    inv_ts, inv_tt = 10.0, 25.0

    def fused_fwd():
        return P.dino_head_forward(xs, xt, vs, one, vt, one, center, ncrops, inv_ts, inv_tt, keep_for_backward=False)[0]

    def fused_step():
        loss, colsum, saved = P.dino_head_forward(xs, xt, vs, one, vt, one, center, ncrops, inv_ts, inv_tt)
        dx, dv, _ = P.dino_head_backward(saved, gout)
        return loss, dx, dv

    dl = P.DINOLoss(K, ncrops, 0.04, 0.04, 0, 1).to(dev)
    c0 = center.view(1, -1).clone()
    xs_u = xs.clone().requires_grad_(True)
    vs_u = vs.clone().requires_grad_(True)

    def unfused_step():
        xs_u.grad = None
        vs_u.grad = None
        dl.center = c0
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ws = vs_u * (one / vs_u.norm(dim=1)).unsqueeze(1)            # torch._weight_norm
            s_out = torch.nn.functional.linear(torch.nn.functional.normalize(xs_u, dim=-1, p=2), ws)
            with torch.no_grad():
                wt = vt * (one / vt.norm(dim=1)).unsqueeze(1)
                t_out = torch.nn.functional.linear(torch.nn.functional.normalize(xt, dim=-1, p=2), wt)
        loss = dl(s_out, t_out, 0)
        loss.backward()
        return loss

    def graph_time(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        mode = "cuda_graph"
        run = None
        try:
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                fn()
            run = gr.replay
            for _ in range(3):
                run()
        except Exception as e:
            mode = "eager (graph capture failed: %s)" % str(e).splitlines()[0][:80]
            torch.cuda.synchronize()
            run = fn
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            run()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters, mode

    out = {"B": B, "ncrops": ncrops, "out_dim": K, "bottleneck": D}
    ms_f, mode_f = graph_time(fused_fwd)
    ms_fb, mode_fb = graph_time(fused_step)
    loss_f = float(fused_step()[0])
    rows_s, rows_t, Dt = ncrops * B, 2 * B, D + 64
    issued = 2.0 * K * (2 * rows_t * Dt + rows_t * D + rows_s * D) + 2.0 * K * D * (2 * rows_s + rows_s + rows_t)
    useful = 2.0 * K * D * (rows_s + rows_t) + 4.0 * K * D * rows_s
    # HBM traffic of the composition: fp32 prototypes in (2), bf16 operand copies out and back (3 + 2 GEMM passes each),
    # probabilities (bf16) written once and read twice, dW out and its weight-norm pass; the logits themselves: none
    probs = (rows_s + rows_t) * K * 2
    out["fused"] = {"ms_fwd": round(ms_f, 4), "ms_fwd_bwd": round(ms_fb, 4), "launch": mode_fb, "loss": loss_f,
                    "TFLOPs_issued": round(issued / ms_fb / 1e9, 1), "TFLOPs_6RKD": round(useful / ms_fb / 1e9, 1),
                    "logit_bytes_in_hbm": 0, "probability_bytes_written": probs,
                    "note": "no [rows, K] logits or fp32 gradient in HBM; bf16 probabilities written once by the "
                            "recomputing GEMM (teacher rows in the forward, student rows in the backward)"}
    if fused_only:
        return out
    try:
        ms_u, mode_u = graph_time(unfused_step)
        loss_u = float(unfused_step())
        out["unfused_same_gpu"] = {"ms_fwd_bwd": round(ms_u, 4), "launch": mode_u, "loss": loss_u,
                                   "logit_bytes_in_hbm": (rows_s + rows_t) * K * 2 + rows_s * K * 2,
                                   "note": "cuBLAS bf16 last-layer GEMMs (logits written, bf16) + this repo's DINO loss kernels "
                                           "(logits read twice, bf16 gradient written) + autograd dX / dW GEMMs and "
                                           "weight-norm / normalize backward in eager PyTorch"}
        out["loss_rel_diff_fused_vs_unfused"] = float("%.3g" % (abs(loss_f - loss_u) / abs(loss_u)))
        out["speedup_fused_over_unfused"] = round(ms_u / ms_fb, 3)
    except Exception as e:
        out["unfused_error"] = str(e).splitlines()[0][:160]
        torch.cuda.synchronize()
    del dl
    torch.cuda.empty_cache()
    if cpu_ref:
        out.update(dino_head_cpu_reference(B, ncrops, K, D))
    return out


def dino_head_cpu_reference(B, ncrops, K, D):
    """The reference's own modules for the fused-head leg's work on the host cores (a reported baseline, bounded:
    3 steps of ~1.5 s).  Never raises: the measured line must not be lost to a baseline leg."""
    if not _reference_available():
        return {}
    try:
        from oracle import ref_step
        torch.set_num_threads(os.cpu_count() or 1)
        sec, loss_c = ref_step.dino_head_reference_step("cpu", B, ncrops, K, D, iters=2)
        return {"cpu_reference": {"ms_fwd_bwd": round(sec * 1e3, 1), "cores": torch.get_num_threads(), "kind": "reference",
                                  "loss": loss_c, "sample": "vision_transformer.DINOHead.last_layer on F.normalize'd "
                                  "features (student + teacher) + lafs_train.DINOLoss forward/backward, fp32, "
                                  "median of 2 steps after one warm-up, full size"}}
    except Exception as e:
        return {"cpu_reference_error": str(e).splitlines()[0][:160]}


HEAD_PARITY_TOL = 2e-3     # max-norm relative; the two runs share operands and lse up to fp32 summation order


# ---------------------------------------------------------------------------------------------
# Reference arms.  The reference is pure Python/PyTorch: __graft_entry__.build() stages the seven files
# the path needs, unmodified, in the git-ignored baseline/_ref/ (SURVEY.md section 7 step 1), and
# oracle/ref_step.py drives those modules through the same hot path.  kind = "reference".  If the copy
# is missing (build() was not run where /root/reference exists) the oracle port stands in, kind = "port".
def _reference_available():
    from oracle import ref_harness
    return ref_harness.available()


def cpu_reference_run(steps, warmup, budget_s=200.0):
    """The unmodified reference modules on the host cores (fp32, all threads).  Every step is a
    bounded sample: the per-face regions (extract, embed, dino) run on `faces` <= B faces and are
    scaled to B, the EMA runs on the full parameter list.  `faces` is chosen after the first warm-up
    step so that steps + warmup end within `budget_s`.  Returns (seconds per B-face step, info)."""
    from oracle import ref_step
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B, L = B_PER_GPU, N_LOCAL
    shapes = vit_param_shapes(VIT)
    faces = min(B, 64)
    st = ref_step.SSLReferenceStep("cpu", faces, L, OUT_DIM, DIM, shapes, seed=0)

    def one(it):
        t = st.step(it)
        per_face = t["extract"] + t["embed"] + t["dino"]
        return per_face * (B / faces) + t["ema"], t

    sec, _ = one(0)                                     # first warm-up step doubles as the probe
    total = max(1, steps + warmup)
    want = B
    while want > 8 and sec * (want / faces) * total > budget_s:
        want //= 2
    if want != faces:
        faces = want
        del st
        st = ref_step.SSLReferenceStep("cpu", faces, L, OUT_DIM, DIM, shapes, seed=0)
    for i in range(max(0, warmup - 1)):
        one(i)
    secs, parts = [], []
    for i in range(max(1, steps)):
        sec, t = one(warmup + i)
        secs.append(sec)
        parts.append(t)
    sec = float(np.median(secs))
    regions = {k: round(float(np.median([p[k] for p in parts])) * (1.0 if k == "ema" else B / faces) * 1e3, 2)
               for k in ("extract", "embed", "dino", "ema")}
    info = {"cores": cores, "kind": "reference", "faces": faces, "ms_by_region": regions,
            "sample": f"unmodified reference modules (baseline/_ref via oracle/ref_step.py), fp32, {cores} threads: "
                      f"extract/embed/DINO regions on {faces} of {B} faces scaled x{B / faces:g}, full {len(shapes)}-tensor EMA"}
    return sec, info


def cpu_port_run(sample_faces, threads, reps=1):
    """Fallback when the reference files are not staged: the oracle port (kind = "port")."""
    from oracle import lafs_oracle as O
    torch.set_num_threads(threads)
    Bs, L = sample_faces, N_LOCAL
    g = torch.Generator().manual_seed(0)
    img_g = torch.rand(2 * Bs, 3, 112, 112, generator=g) * 2 - 1
    img_l = torch.rand(L * Bs, 3, 112, 112, generator=g) * 2 - 1
    raw_g = torch.randn(2 * Bs, 392, generator=g)
    raw_l = torch.randn(L * Bs, 392, generator=g)
    s = torch.randn((L + 2) * Bs, OUT_DIM, generator=g)
    t = torch.randn(2 * Bs, OUT_DIM, generator=g)
    center = torch.zeros(1, OUT_DIM)
    q = [torch.randn(*sh, generator=g) for sh in vit_param_shapes(VIT)]
    k = [p.clone() for p in q]
    ws, bs, wt, bt = q[1], q[2], k[1], k[2]     # patch_to_embedding of student / teacher
    best_face, best_ema = float("inf"), float("inf")
    for _ in range(reps + 1):
        t0 = time.perf_counter()
        th_g = O.landmark_post(raw_g, torch.randn(2 * Bs, 196, 2) * 5)
        tok_g = O.extract_tokens(img_g, th_g)
        O.patch_embed(tok_g, ws, bs); O.patch_embed(tok_g, wt, bt)
        th_l = O.landmark_post(raw_l, torch.randn(L * Bs, 196, 2) * 5, torch.randint(0, 196, (L * Bs, 36, 1)))
        O.patch_embed(O.extract_tokens(img_l, th_l), ws, bs)
        O.dino_loss_and_grad(s, t, center, L + 2, 0.04)
        O.dino_center_update(center, t)
        t1 = time.perf_counter()
        O.ema_update_(k, q, 0.996)
        t2 = time.perf_counter()
        best_face, best_ema = min(best_face, t1 - t0), min(best_ema, t2 - t1)
    return best_face * (B_PER_GPU / Bs) + best_ema


def cpu_baseline():
    """cpu_baseline leg of our own line (rank 0, N = 1): a short run of the reference arm, in a fresh process
    with the original CPU affinity (this process is pinned to the GPU's NUMA node for the end-to-end loop)."""
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "2", "--warmup", "1",
                            "--config", WL_KEY], capture_output=True, text=True, timeout=900,
                           preexec_fn=lambda: os.sched_setaffinity(0, _ORIG_AFFINITY),
                           env={k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")})
        return json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])["cpu_baseline"]
    except Exception:
        pass                                  # fall through: run it here
    if _reference_available():
        sec, info = cpu_reference_run(steps=2, warmup=1, budget_s=25.0)
    else:
        cores = os.cpu_count() or 1
        sec = cpu_port_run(min(B_PER_GPU, 64), cores)
        info = {"cores": cores, "kind": "port", "sample": "oracle port (baseline/_ref not staged), 64 faces scaled"}
    out = {"value": round(B_PER_GPU / sec, 2), "unit": "faces/s"}
    out.update(info)
    return out


def reference_eager_gpu(dev, steps=3):
    """The same unmodified reference modules in eager PyTorch on this GPU (fp16 autocast as in
    lafs_train.py:577): per region and per step -- the practical kernel to beat."""
    from oracle import ref_step
    out = {}
    try:
        st = ref_step.SSLReferenceStep(dev, B_PER_GPU, N_LOCAL, OUT_DIM, DIM, vit_param_shapes(VIT), seed=0)
        st.step(0)
        ts = [st.step(1 + i) for i in range(steps)]
        reg = {k: float(np.median([t[k] for t in ts])) * 1e3 for k in ("extract", "embed", "dino", "ema")}
        ms = sum(reg.values())
        out["ssl_step"] = {"ms_per_step": round(ms, 3), "faces_per_s": round(B_PER_GPU / ms * 1e3, 1),
                           "ms_by_region": {k: round(v, 3) for k, v in reg.items()},
                           "note": "reference modules, eager CUDA, fp16 autocast; device-resident inputs; regions timed with "
                                   "CUDA events, launch overhead included (that is the reference's cost)"}
        del st
        torch.cuda.empty_cache()
    except Exception as e:
        out["ssl_step_error"] = str(e).splitlines()[0][:160]
    try:
        sec = ref_step.student_update_reference_step(dev, vit_param_shapes(VIT), iters=3)
        out["student_update"] = {"ms": round(sec * 1e3, 3), "note": "utils.clip_gradients + torch.optim.AdamW.step + EMA loop, eager CUDA"}
    except Exception as e:
        out["student_update_error"] = str(e).splitlines()[0][:160]
    torch.cuda.empty_cache()
    for name, B, C, D in HEAD_CFGS:
        if D != 512:
            continue
        try:
            sec, _ = ref_step.head_reference_step(dev, B, C, D, iters=3)
            out["head_" + name] = {"ms_fwd_bwd": round(sec * 1e3, 3), "faces_per_s": round(B / sec, 1),
                                   "note": "VF.CosFace + CrossEntropyLoss fwd+bwd, eager CUDA fp32 (CPU one-hot + H2D as in "
                                           "ViT_face.py:67-82)"}
        except Exception as e:
            out["head_" + name + "_error"] = str(e).splitlines()[0][:160]
        torch.cuda.empty_cache()
    return out


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path on the box's host cores,
    honouring --steps / --warmup.  Rank 0 alone runs; the other ranks exit 0 without work."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if _reference_available():
        sec, info = cpu_reference_run(args.steps, args.warmup)
        steps, warm = max(1, args.steps), args.warmup
    else:
        cores = os.cpu_count() or 1
        secs = [cpu_port_run(min(B_PER_GPU, 64), cores, reps=0) for _ in range(max(1, min(args.steps, 5)))]
        sec, steps, warm = float(np.median(secs)), len(secs), 0
        info = {"cores": cores, "kind": "port", "sample": "oracle port (baseline/_ref not staged), 64 faces scaled"}
    val = round(B_PER_GPU / sec, 2)
    cb = {"value": val, "unit": "faces/s"}
    cb.update(info)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "faces/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": round(sec * 1e3, 2), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WL_NAME + " -- on host cores (bounded sample)", "batch_per_gpu": B_PER_GPU,
                   "out_dim": OUT_DIM, "ncrops": N_LOCAL + 2},
        "cpu_baseline": cb,
        "e2e": {"value": val, "unit": "faces/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


HEAD_CFGS = [("cosface_ms1mv3", 512, 93431, 512), ("arcface_webface4m", 1024, 205990, 512),
             ("cosface_ms1mv3_d768", 512, 93431, 768), ("arcface_webface4m_d768", 1024, 205990, 768)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg1", choices=sorted(WORKLOADS), help="SSL workload (default: BASELINE configs[1])")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip the reference_eager_b200 report")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    args = ap.parse_args()
    set_workload(args.config)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
