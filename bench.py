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
METRIC = "lafs_ssl_hot_path_faces_per_sec"


def vit_b_param_shapes(out_dim=OUT_DIM):
    """The 147 parameter tensors EMA'd every step: Part-fViT ViT-B backbone (incl. the unused
    CosFace weight [30000,768]) + DINOHead.  110,261,248 parameters."""
    shapes = [(1, 197, 768), (768, 192), (768,), (1, 1, 768)]
    for _ in range(12):
        shapes += [(768,), (768,), (2112, 768), (768, 704), (768,), (768,), (768,), (2048, 768), (2048,),
                   (768, 2048), (768,)]
    shapes += [(768,), (768,), (30000, 768)]
    shapes += [(2048, 768), (2048,), (2048, 2048), (2048,), (256, 2048), (256,), (out_dim, 1), (out_dim, 256)]
    return shapes


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
    }
    shapes = vit_b_param_shapes()
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
    names = ["landmark+gather_embed", "dino_fwd+center", "dino_bwd", "ema"]

    def step(inp, it, evs=None):
        def mark(i):
            if evs is not None:
                evs[i].record()
        mark(0)
        path.landmarks_and_embeddings(st["raw_g"], inp["noise_g"], inp["img_g"], st["raw_l"], inp["noise_l"],
                                      inp["idx_l"], inp["img_l"])
        mark(1)
        s = st["student_out"].requires_grad_(True)
        s.grad = None
        loss = path.loss(s, st["teacher_out"], it % 41)
        mark(2)
        loss.backward()
        mark(3)
        path.ema_step(sched[it])
        mark(4)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, per_kernel=False):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(steps)] if per_kernel else None
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
            parts = [float(np.mean([evs[i][j].elapsed_time(evs[i][j + 1]) for i in range(steps)])) for j in range(4)]
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
    ginp = dict(dev_in, raw_g=st["raw_g"], raw_l=st["raw_l"], student_out=st["student_out"], teacher_out=st["teacher_out"])
    graphed = GraphedSSLStep(path, ginp, epoch=3, momentum=float(sched[1000]))
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
        h2d = sum(v.numel() * v.element_size() for v in hbuf.values())
        bufs = [{k: torch.empty_like(v, device=dev) for k, v in hbuf.items()} for _ in range(2)]
        copy_stream = torch.cuda.Stream(device=dev)
        ready = [torch.cuda.Event() for _ in range(2)]
        free = [torch.cuda.Event() for _ in range(2)]

        def upload(i):
            b = i & 1
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(free[b])
                for k, v in hbuf.items():
                    bufs[b][k].copy_(v, non_blocking=True)
                ready[b].record(copy_stream)

        def loop(nsteps):
            for b in range(2):
                free[b].record()
            upload(0)
            last = 0.0
            for i in range(nsteps):
                if i + 1 < nsteps:
                    upload(i + 1)
                torch.cuda.current_stream().wait_event(ready[i & 1])
                loss = step(bufs[i & 1], args.warmup + i)
                free[i & 1].record()
                last = float(loss.item())          # device -> host read of the step's result
            return last

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
        return float(t.item()), h2d

    ms_e2e, h2d = run_e2e(host)
    ms_e2e_f32, h2d_f32 = run_e2e(host_f32)
    clocks = sampler.stop() if sampler else None    # sampled across every timed region above
    xc = getattr(path.loss, "_xchg", None)
    if xc is not None:
        xc.check()                    # raises if a rank missed the centre exchange's spin bound

    # ---- secondary: class-sharded margin head (BASELINE configs[2], configs[3]) -----------------
    del st, path, dev_in
    torch.cuda.empty_cache()
    head = bench_head(P, world, rank, dev, dist, args)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    ms_step = ms_graph / args.steps
    ms_step_eager = ms_total / args.steps
    faces = B * world
    K, nc = OUT_DIM, L + 2
    nparam = sum(int(np.prod(s)) for s in vit_b_param_shapes())
    tok_out = (2 * 2 * B * 196 + L * B * 36) * 768 * 2
    alg = {
        # images + landmarks in, bf16 tokens of both models out (e_img = 1 byte)
        "landmark+gather_embed": {"bytes": (2 + L) * B * (3 * 112 * 112 * 1) + 2 * B * 196 * 8 + L * B * 36 * 8 + tok_out,
                                  "flops": 2.0 * 192 * 768 * (2 * 2 * B * 196 + L * B * 36), "write_bytes": tok_out},
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
    kernels["landmark+gather_embed"]["note"] = ("write-dominated: %.0f MB of bf16 tokens out; a pure-write stream on this "
                                                "part measures 3.9 TB/s (tools/bw_probe.py), i.e. >= %.3f ms"
                                                % (tok_out / 1e6, tok_out / 3.9e9))
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
        "config": {"workload": "BASELINE configs[1]: LAFS SSL pretrain hot path, ViT-B, 196 landmark patches, "
                               "out_dim 65536, 2 global + 4 local crops, batch 256 per GPU",
                   "batch_per_gpu": B, "out_dim": K, "ncrops": nc, "ema_params": nparam, "ema_tensors": len(vit_b_param_shapes()),
                   "images": "uint8 decoded pixels, normalised in-kernel (value_fp32_images / e2e_fp32_images: fp32 tensors)",
                   "l2": "inputs larger than L2 (logits 335 MB, tokens out 362 MB, parameters 882 MB per step)",
                   "parallelism": f"dp{world}", "centre_exchange": centre_exchange},
        "e2e": {"value": round(faces / (ms_e2e / args.steps / 1e3), 1), "unit": "faces/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": round(ms_e2e / args.steps, 5),
                "transport": "uint8 pixels + fp32 noise + int64 indices from pinned memory on a copy stream (double "
                             "buffered); ToTensor/Normalize fused into the gather kernel"},
        "e2e_fp32_images": {"value": round(faces / (ms_e2e_f32 / args.steps / 1e3), 1), "unit": "faces/s",
                            "h2d_bytes_per_step": h2d_f32, "ms_per_step": round(ms_e2e_f32 / args.steps, 5),
                            "transport": "the reference's fp32 normalised image tensors (PCIe-bound)"},
        "value_fp32_images": round(faces / (ms_total_f32 / args.steps / 1e3), 1),
        "value_eager_launches": round(faces / (ms_step_eager / 1e3), 1),
        "launch": "value: the step's 11 kernels replayed as one CUDA graph (lafs_cvpr2024_b200.ssl_step.GraphedSSLStep); "
                  "value_eager_launches / kernels / e2e: the same kernels launched one by one from Python",
        "gpu_launches": 11,   # 2 landmark, 3 weight prep, 2 gather-embed, 2 dino fwd (+centre), 1 dino bwd, 1 ema
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
    if not args.no_cpu and world == 1:          # the CPU baseline is a rank-0, N = 1 report
        line["cpu_baseline"] = cpu_baseline(sample_faces=args.cpu_faces)
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
        del h, x
        torch.cuda.empty_cache()
        return ms, ms_eager, mode

    out = {}
    cfgs = [("cosface_ms1mv3", P.CosFace, 512, 93431, 512), ("arcface_webface4m", P.ArcFace, 1024, 205990, 512)]
    for name, cls, B, C, D in cfgs:
        res = {}
        for variant in (["nccl", "peer"] if world > 1 else ["single"]):
            try:
                res[variant] = run_variant(cls, B, C, D, variant == "peer")
            except Exception as e:            # e.g. symmetric memory unavailable on this box: NCCL number stands
                res[variant + "_error"] = str(e).splitlines()[0][:120]
                torch.cuda.synchronize()
        timed = {k: v for k, v in res.items() if isinstance(v, tuple)}
        best = min(timed, key=lambda k: timed[k][0])
        ms, ms_eager, mode = timed[best]
        out[name] = {"B_global": B, "classes": C, "D": D, "shards": world, "ms_fwd_bwd": round(ms, 4), "launch": mode,
                     "ms_fwd_bwd_eager": round(ms_eager, 4), "exchange": best,
                     "faces_per_s": round(B / ms * 1e3, 1), "TFLOPs_6BCD": round(6.0 * B * C * D / ms / 1e9 / world, 1),
                     "note": "TFLOPs per GPU on the 6*B*C*D/R count; the step also recomputes the logits once (8*B*C*D issued)"}
        if world > 1:
            out[name]["ms_by_exchange"] = {k: round(v[0], 4) for k, v in timed.items()}
            out[name].update({k: v for k, v in res.items() if k.endswith("_error")})
    return out


# ---------------------------------------------------------------------------------------------
def cpu_step_time(sample_faces, threads, reps=1):
    """The oracle port of the same step on the host cores: per-face parts on `sample_faces`
    faces, the (batch-independent) EMA on the full parameter list.  Returns seconds per
    256-face step, extrapolated linearly in the per-face parts."""
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
    shapes = vit_b_param_shapes()
    q = [torch.randn(*sh, generator=g) for sh in shapes]
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
    return best_face * (B_PER_GPU / Bs) + best_ema, best_face, best_ema


def cpu_baseline(sample_faces=256):
    cores = os.cpu_count() or 1
    sec, t_face, t_ema = cpu_step_time(sample_faces, cores, reps=3)
    return {"value": round(B_PER_GPU / sec, 2), "unit": "faces/s", "cores": cores, "kind": "port",
            "sample": f"oracle port (torch-CPU fp32, {cores} threads): per-face parts timed on {sample_faces} of 256 faces "
                      f"({t_face:.2f} s) and scaled x{B_PER_GPU // sample_faces}; full 147-tensor EMA timed once ({t_ema:.2f} s)"}


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path.  The reference is
    Python and cannot travel to the GPU box (no /root/reference there), so this is the oracle
    port (pinned bit-for-bit to the reference by tests/golden), on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    faces = args.cpu_faces
    cpu_step_time(min(faces, 8), cores, reps=0)     # untimed: thread pool / allocator warm-up
    secs = []
    for _ in range(max(1, min(args.steps, 5))):
        sec, t_face, t_ema = cpu_step_time(faces, cores, reps=1)
        secs.append(sec)
    sec = float(np.median(secs))
    val = round(B_PER_GPU / sec, 2)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "faces/s", "n_gpus": args.gpus,
        "steps": len(secs), "warmup": 1, "ms_per_step": round(sec * 1e3, 2), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1] hot path on host cores (bounded sample)", "batch_per_gpu": B_PER_GPU,
                   "out_dim": OUT_DIM, "ncrops": N_LOCAL + 2},
        "cpu_baseline": {"value": val, "unit": "faces/s", "cores": cores, "kind": "port",
                         "sample": f"per-face parts on {faces} of 256 faces scaled x{B_PER_GPU // faces} + full EMA; "
                                   "oracle port of the reference (Python reference cannot travel to the GPU box)"},
        "e2e": {"value": val, "unit": "faces/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-faces", type=int, default=256, help="faces in the bounded CPU sample (256 = the full step)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
