"""Import harness for the UNMODIFIED reference (test infrastructure, not product code).

Only usable where /root/reference exists (the build container).  It is used by
tests/golden/make_golden.py to generate the committed golden vectors and by
tests that pin oracle/lafs_oracle.py against the real reference.  Nothing in the
product package imports this file, and nothing on the GPU box needs it.

Recipe follows SURVEY.md Appendix A: two import shims (IPython, timm.models.layers)
and an identity patch for the hard-coded .cuda() calls when no GPU is present.
"""
import contextlib
import io
import os
import sys
import types

REF_ROOT = os.environ.get("LAFS_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "lafs_train.py"))


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
        dist.init_process_group("gloo", rank=0, world_size=1)
    _cached = types.SimpleNamespace(VF=VF, L=L, vt=vt, dutils=dutils, mixup=mixup_my)
    return _cached
