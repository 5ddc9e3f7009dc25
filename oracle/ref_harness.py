"""Import harness for the UNMODIFIED reference (test infrastructure, not product code).

It is used by tests/golden/make_golden.py to generate the committed golden vectors, by the
tests that pin oracle/lafs_oracle.py against the real reference (build container only:
/root/reference) and by bench.py's reference arms (oracle/ref_step.py), which on the GPU box
import the git-ignored copy baseline/_ref/ that __graft_entry__.build() makes of the seven
files the path needs.  Nothing in the product package imports
this file; the -m gpu tests and smoke() do not use it.

Recipe follows SURVEY.md Appendix A: two import shims (IPython, timm.models.layers)
and an identity patch for the hard-coded .cuda() calls when no GPU is present.
"""
import contextlib
import io
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LOCAL_COPY = os.path.join(_REPO, "baseline", "_ref")
# the files of the reference the hot path needs
REF_FILES = ["lafs_train.py", "utils.py", "vision_transformer.py", "face_pre_pro/ViT_face.py",
             "face_pre_pro/mobilenet.py", "util/__init__.py", "util/mixup_my.py"]


def _find_root():
    for cand in (os.environ.get("LAFS_REFERENCE_ROOT"), "/root/reference", LOCAL_COPY):
        if cand and os.path.isfile(os.path.join(cand, "lafs_train.py")):
            return cand
    return "/root/reference"


REF_ROOT = _find_root()


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "lafs_train.py"))


def stage_local_copy(src="/root/reference") -> bool:
    """Copies the seven reference files, byte for byte, into baseline/_ref/ (git-ignored, NOT
    gpurun-ignored) so that bench.py's reference arms can run the unmodified reference on the GPU
    box.  No-op when `src` is absent (the GPU box)."""
    import shutil
    if not os.path.isfile(os.path.join(src, "lafs_train.py")):
        return False
    for rel in REF_FILES:
        dst = os.path.join(LOCAL_COPY, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(src, rel), dst)
    return True


_cached = None


def load():
    """Returns a namespace with the reference modules: .VF (face_pre_pro.ViT_face),
    .L (lafs_train), .vt (vision_transformer), .dutils (utils), .mixup (util.mixup_my)."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    import torch
    import torch.nn as nn
    import torch.distributed as dist

    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    ip = types.ModuleType("IPython")
    ip.embed = lambda *a, **k: None
    sys.modules.setdefault("IPython", ip)
    with contextlib.redirect_stdout(io.StringIO()):
        import vision_transformer as vt
        import utils as dutils
    tl = types.ModuleType("timm.models.layers")
    tl.DropPath, tl.trunc_normal_ = vt.DropPath, dutils.trunc_normal_
    tm = types.ModuleType("timm.models")
    tm.layers = tl
    t = types.ModuleType("timm")
    t.models = tm
    for k, v in {"timm": t, "timm.models": tm, "timm.models.layers": tl}.items():
        sys.modules.setdefault(k, v)
    with contextlib.redirect_stdout(io.StringIO()):
        from face_pre_pro import ViT_face as VF
        import lafs_train as L
        from util import mixup_my
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        nn.Module.cuda = lambda self, *a, **k: self
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        # both device types: the CPU arm all-reduces a CPU tensor, the eager-GPU report a CUDA one (lafs_train.py:675)
        dist.init_process_group("cpu:gloo,cuda:nccl" if torch.cuda.is_available() else "gloo", rank=0, world_size=1)
    _cached = types.SimpleNamespace(VF=VF, L=L, vt=vt, dutils=dutils, mixup=mixup_my)
    return _cached
