"""The C-ABI shared library loads and exports every symbol include/cuda_audio_b200.h declares.
No compute calls: this runs without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cuda_audio_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ca_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_the_reference_boundary():
    names = declared_functions()
    for must in ("ca_create", "ca_destroy", "ca_load_ir", "ca_set_params", "ca_process", "ca_get_stats"):
        assert must in names
    assert len(names) >= 18


def test_library_exports_every_declared_symbol():
    import cuda_audio_b200 as m
    m.build()
    L = ctypes.CDLL(m.LIB_PATH)
    for name in declared_functions():
        assert hasattr(L, name), name
    assert sorted(m.EXPORTS) == declared_functions()
    assert L.ca_api_version() == 1


def test_struct_layouts_match_the_header():
    import cuda_audio_b200 as m
    # ca_config: 14 u32/i32 + 2*4 u32 + float = 23 words ; ca_params: 9 words
    assert ctypes.sizeof(m.Config) == 23 * 4
    assert ctypes.sizeof(m.Params) == 9 * 4
    cfg = m.default_config()
    assert cfg.struct_size == ctypes.sizeof(m.Config)
    assert (cfg.period, cfg.n_in, cfg.n_out) == (256, 2, 2)
    assert cfg.max_ir_frames == 512 * 256 - 1024  # CONV_DEFAULT_FFTSIZE - default nframes


def test_no_cpu_fallback_without_a_gpu():
    """On a box without CUDA the product must fail loudly, not fall back."""
    import cuda_audio_b200 as m
    from tests import conftest
    if conftest._has_gpu():
        pytest.skip("GPU present")
    with pytest.raises(m.CaError) as ei:
        m.Engine(period=64, max_ir_frames=128)
    assert ei.value.code == -2
    assert "no CPU fallback" in str(ei.value)


def test_product_does_not_reference_the_oracle():
    """Only tests/, smoke() and bench.py may touch oracle/."""
    pkg = os.path.join(ROOT, "cuda-audio_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".cpp", ".py", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in txt.lower() or f == "Makefile" and "oracle" not in txt, os.path.join(dp, f)


def test_auto_tiers_plan_is_valid():
    """ca_config_auto_tiers (host logic, no GPU): tier 0 == period, blocks grow by `growth`, every
    tier >= 1 starts at an IR offset >= its block size, the last tier covers the rest."""
    import cuda_audio_b200 as m
    for B, L, growth, maxb in [(64, 480000, 0, 0), (256, 192000, 0, 0), (32, 5000, 0, 0), (1024, 200000, 0, 0), (64, 700, 0, 0),
                               (256, 2880000, 0, 0), (256, 192000, 4, 4096), (64, 480000, 16, 16384), (128, 300, 0, 0)]:
        cfg = m.default_config(period=B, max_ir_frames=L)
        assert m.lib().ca_config_auto_tiers(ctypes.byref(cfg), growth, maxb) == 0
        off = 0
        assert cfg.tier_block[0] == B and 1 <= cfg.n_tiers <= 4
        for j in range(cfg.n_tiers):
            S, P = cfg.tier_block[j], cfg.tier_parts[j]
            assert S & (S - 1) == 0
            if j:
                assert off >= S and S > cfg.tier_block[j - 1] and 256 <= S <= (maxb or 16384)
            if j + 1 < cfg.n_tiers:
                assert P > 0
                off += S * P
            else:
                assert P == 0          # the engine extends the last tier to cover max_ir_frames
        assert off < L or cfg.n_tiers == 1
    cfg = m.default_config(period=100)
    assert m.lib().ca_config_auto_tiers(ctypes.byref(cfg), 0, 0) == -1
    cfg = m.default_config(period=64)
    assert m.lib().ca_config_auto_tiers(ctypes.byref(cfg), 3, 0) == -1
