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
    # ca_config: 14 u32/i32 + 2*4 u32 + float + voice_pool + schedule + io_chunks + sm_split = 27 words ; ca_params: 9 words
    assert ctypes.sizeof(m.Config) == 28 * 4
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


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """The boundary is a C ABI: a C99 translation unit (gcc -std=c99 -pedantic -Werror) includes the header,
    calls the configuration entry points and links against the shared library -- what a cgo / JNI / ctypes
    binding relies on.  Without a GPU ca_create must return CA_ERR_CUDA, never fall back."""
    import shutil
    import subprocess
    import cuda_audio_b200 as m
    m.build()
    src = tmp_path / "consumer.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "cuda_audio_b200.h"
int main(void)
{
    ca_config cfg;
    ca_engine *e = NULL;
    int rc;
    ca_config_init(&cfg);
    cfg.period = 64; cfg.max_ir_frames = 480000; cfg.flags = CA_FLAG_ASYNC_TIERS;
    rc = ca_config_auto_tiers(&cfg, 0, 0);
    printf("api=%d rc=%d tiers=%u", ca_api_version(), rc, cfg.n_tiers);
    { unsigned j; for (j = 0; j < cfg.n_tiers; j++) printf(" %ux%u", cfg.tier_block[j], cfg.tier_parts[j]); }
    rc = ca_create(&cfg, &e);
    printf(" create=%d (%s)\n", rc, ca_strerror(rc));
    if (rc == CA_OK) ca_destroy(e);
    return 0;
}
''')
    exe = tmp_path / "consumer"
    lib_dir = os.path.dirname(m.LIB_PATH)
    cc = shutil.which("gcc")
    r = subprocess.run([cc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                        "-L", lib_dir, "-lcuda_audio_b200", "-Wl,-rpath," + lib_dir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120).stdout
    assert "api=1 rc=0 tiers=4 64x9 512x7 4096x3 16384x0" in out, out   # one more tier-0 partition than the synchronous plan
    from tests import conftest
    assert ("create=0" in out) if conftest._has_gpu() else ("create=-2" in out), out


def test_invalid_configs_are_refused_before_any_cuda_call():
    """The C ABI reports errors, it never aborts (SURVEY 8b "error convention"; the reference asserts, conv.cu:152-194).
    Argument validation comes before the device check, so it is testable without a GPU: every malformed config is
    CA_ERR_INVALID (-1) with a reason in ca_last_error_string(), not CA_ERR_CUDA (-2) and not a crash."""
    import ctypes as C

    import cuda_audio_b200 as m
    L = m.lib()

    def create(**kw):
        cfg = m.default_config()
        for k, v in kw.items():
            setattr(cfg, k, v)
        h = C.c_void_p(0xdead)
        rc = L.ca_create(C.byref(cfg), C.byref(h))
        assert rc != 0 and not h.value, kw      # the out pointer is cleared on failure
        return rc, L.ca_last_error_string().decode()

    for bad in (dict(period=0), dict(period=48), dict(period=16), dict(period=2048), dict(n_in=0), dict(n_in=3), dict(n_out=0), dict(n_out=3),
                dict(n_instances=0), dict(n_instances=(1 << 20) + 1), dict(max_ir_frames=0), dict(n_ir_slots=0), dict(max_voices=5), dict(struct_size=8)):
        rc, why = create(**bad)
        assert rc == -1 and why, (bad, rc, why)
    assert L.ca_create(None, None) == -1
    cfg = m.default_config()
    assert L.ca_create(C.byref(cfg), None) == -1


def test_null_handles_are_errors_not_crashes():
    import ctypes as C

    import cuda_audio_b200 as m
    L = m.lib()
    p = m.Params()
    st = m.Stats()
    buf = (C.c_float * 8)()
    f32p = C.POINTER(C.c_float)
    none = C.c_void_p(None)
    assert L.ca_load_ir(none, 0, C.cast(buf, f32p), C.cast(buf, f32p), 8) == -1
    assert L.ca_set_params(none, 0, 0, C.byref(p)) == -1
    assert L.ca_get_params(none, 0, 0, C.byref(p)) == -1
    assert L.ca_set_glide(none, 0, 0, C.c_float(1.0)) == -1
    assert L.ca_set_active(none, 1) == -1
    assert L.ca_reset(none) == -1
    assert L.ca_process(none, C.cast(buf, f32p), C.cast(buf, f32p), 4) == -1
    assert L.ca_sync(none) == -1
    assert L.ca_get_stats(none, C.byref(st)) == -1
    assert L.ca_reset_stats(none) == -1
    assert L.ca_destroy(none) in (0, -1)     # destroying nothing is harmless
    assert L.ca_strerror(-1) and L.ca_strerror(-6) and L.ca_strerror(12345)


def test_group_entry_points_validate_without_a_gpu():
    """ca_group_* (one long IR over the GPUs of a node): malformed configs and null handles are CA_ERR_INVALID."""
    import ctypes as C

    import cuda_audio_b200 as m
    L = m.lib()

    def create(**kw):
        cfg = m.GroupConfig()
        L.ca_group_config_init(C.byref(cfg))
        assert cfg.struct_size == C.sizeof(m.GroupConfig)
        for k, v in kw.items():
            setattr(cfg, k, v)
        h = C.c_void_p(0xdead)
        rc = L.ca_group_create(C.byref(cfg), C.byref(h))
        assert not h.value
        return rc

    assert create(n_devices=0) == -1 and create(n_devices=9) == -1 and create(exchange=7) == -1 and create(struct_size=4) == -1
    from tests import conftest
    if not conftest._has_gpu():
        assert create() == -2 and "no CPU fallback" in L.ca_last_error_string().decode()
    none = C.c_void_p(None)
    p = m.Params()
    st = m.GroupStats()
    buf = (C.c_float * 8)()
    f32p = C.POINTER(C.c_float)
    assert L.ca_group_load_ir(none, 0, C.cast(buf, f32p), C.cast(buf, f32p), 8) == -1
    assert L.ca_group_set_params(none, 0, C.byref(p)) == -1
    assert L.ca_group_set_glide(none, 0, C.c_float(1.0)) == -1
    assert L.ca_group_process(none, C.cast(buf, C.c_void_p), C.cast(buf, C.c_void_p), 4) == -1
    assert L.ca_group_get_stats(none, C.byref(st)) == -1
    assert L.ca_group_reset_stats(none) == -1
    assert L.ca_group_destroy(none) in (0, -1)
