"""CPU-only checks of the drop-in boundary: libllc.so builds for sm_100a, loads, and exports every
symbol include/llc.h declares (no compute calls: there is no GPU here and no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from lifelong_clip_b200 import _capi
    return _capi.load()


def declared_symbols():
    with open(os.path.join(ROOT, "include", "llc.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(llc_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from lifelong_clip_b200 import _capi
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in llc.h but not exported by libllc.so"
    assert sorted(_capi.SIGNATURES) == names, (
        set(names) ^ set(_capi.SIGNATURES))  # the ctypes table mirrors the header exactly


def test_version_and_error_slot(lib):
    assert lib.llc_version() == 100
    assert isinstance(lib.llc_last_error(), bytes)


def test_argument_errors_are_reported_before_any_launch(lib):
    from lifelong_clip_b200 import _capi
    # bad arguments are rejected on the host (rc < 0) with a message; nothing touches a device
    rc = lib.llc_label_remap(None, None, 0, None, -1, None)
    assert rc == -1 and b"llc_label_remap" in lib.llc_last_error()
    assert lib.llc_label_remap(None, None, 0, None, 0, None) == 0  # empty batch is a no-op
    e = _capi.GemmEpi()
    rc = lib.llc_gemm_bf16_tn(1 << 20, 40, 1 << 20, 40, 64, 64, 40, ctypes.byref(e), None)
    assert rc == -1 and b"multiple of 16" in lib.llc_last_error()
    cfg = _capi.VitCfg()
    cfg.image_size, cfg.patch, cfg.width, cfg.layers, cfg.heads = 224, 16, 768, 12, 11
    cfg.mlp_dim, cfg.embed_dim, cfg.lora_r, cfg.lora_scale = 3072, 512, 4, 0.25
    assert lib.llc_vit_arena_bytes(ctypes.byref(cfg), 4, 1) == 0   # heads*64 != width
    cfg.heads = 12
    nbytes = lib.llc_vit_arena_bytes(ctypes.byref(cfg), 32, 1)
    assert 2 ** 30 < nbytes < 4 * 2 ** 30   # ~1 GB per layer at 256 images -> ~1.6 GB at 32


def test_product_path_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from lifelong_clip_b200.adapter_clip import AdapterCLIP
    m = AdapterCLIP(vision_config=(32, 8, 128, 2, 64))
    with pytest.raises(RuntimeError, match="CUDA"):
        m.model.encode_image(torch.zeros(1, 3, 32, 32))


def test_no_product_module_imports_the_oracle():
    pkg = os.path.join(ROOT, "lifelong-clip_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            with open(os.path.join(pkg, fn)) as f:
                src = f.read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn
