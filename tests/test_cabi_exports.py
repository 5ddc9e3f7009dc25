"""CPU: the C-ABI library builds/loads and exports every symbol include/lafs_b200.h declares;
argument validation that needs no GPU; the product path refuses CPU tensors."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from lafs_cvpr2024_b200 import _lib, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    return _lib


def declared_symbols():
    names = []
    for f in os.listdir(os.path.join(ROOT, "include")):
        src = open(os.path.join(ROOT, "include", f)).read()
        names += re.findall(r"LAFS_API\s+[\w\s\*]+?\b(lafs_\w+)\s*\(", src)
    return sorted(set(names))


def test_every_declared_symbol_is_exported_and_bound(lib):
    names = declared_symbols()
    assert len(names) >= 13
    h = lib.lib()
    for n in names:
        assert hasattr(h, n), f"{n} declared in include/ but not exported"
        assert n in lib.SIGNATURES, f"{n} has no ctypes signature"
    assert set(lib.SIGNATURES) == set(names)
    assert h.lafs_version() >= 100


def test_argument_errors_without_gpu(lib):
    h = lib.lib()
    assert h.lafs_dino_workspace_bytes(256, 65536, 6) > 0
    assert h.lafs_dino_workspace_bytes(256, 65536, 1) == 0
    assert h.lafs_dino_workspace_bytes(0, 65536, 6) == 0
    # null pointers / bad sizes are rejected before any CUDA call
    assert h.lafs_dino_fwd(None, None, None, 4, 1024, 6, 10.0, 25.0, 1, None, None, None, None, 0, None, 0.9, 0.1, None) == -1
    assert b"null" in h.lafs_last_error_string()
    assert h.lafs_gather_fwd(None, None, None, 1, 3, 112, 112, 196, 0, 0, None) == -1
    assert h.lafs_ema_multi(None, 0, 0.9, 0.1, 0, None) == 0       # empty list is a no-op
    assert h.lafs_ema_multi(None, 3, 0.9, 0.1, 0, None) == -1


def test_host_only_planning_entries(lib):
    """Sizing / layout entry points are pure host code: callable without a GPU."""
    import ctypes as C
    h = lib.lib()
    offs = (C.c_size_t * 5)()
    for world, B, D in ((2, 512, 512), (8, 1024, 512), (8, 1, 65536), (1, 7, 64)):
        total = h.lafs_xchg_bytes(world, B, D, offs)
        flags, slots, o_in, o_out, err = list(offs)
        assert flags == 0 and flags < err < slots < o_in < o_out < total
        assert slots % 256 == 0 and o_in % 256 == 0 and o_out % 256 == 0 and total % 256 == 0
        assert o_in - slots >= 2 * world * B * 16            # two parities of [world][B] float4 records
        assert o_out - o_in >= B * D * 4 and total - o_out >= B * D * 4
    assert h.lafs_xchg_bytes(9, 512, 512, offs) == 0 and h.lafs_xchg_bytes(0, 512, 512, offs) == 0
    assert h.lafs_xchg_bytes(2, 512, 512, None) > 0
    assert h.lafs_embed_bwd_workspace_bytes(100352, 768) >= 768 * 192 * 4
    assert h.lafs_embed_bwd_workspace_bytes(0, 768) == 0
    assert h.lafs_head_workspace_bytes(512, 93431, 512) > 0 and h.lafs_head_workspace_bytes(512, 93431, 100) == 0
    assert h.lafs_head_bwd_workspace_bytes(512, 93431, 512) >= 512 * 512 * 4
    # null pointers are rejected before any CUDA call
    assert h.lafs_embed_bwd_weight(None, None, 128, 768, None, None, 0, None) == -1
    assert h.lafs_embed_bwd_tokens(None, None, 128, 768, None, None) == -1


def test_no_cpu_fallback():
    import lafs_cvpr2024_b200 as P
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        P.ema_update_([torch.zeros(4)], [torch.zeros(4)], 0.99)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        P.extract_patches_pytorch_gridsample(torch.zeros(1, 3, 112, 112), torch.zeros(1, 196, 2),
                                             torch.tensor([8, 8]), 196)
    loss = P.DINOLoss(64, 4, 0.04, 0.07, 3, 10)
    assert list(loss.state_dict().keys()) == ["center"]          # checkpoint contract
    assert loss.center.shape == (1, 64)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        loss(torch.zeros(8, 64), torch.zeros(4, 64), 0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "lafs_cvpr2024_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
