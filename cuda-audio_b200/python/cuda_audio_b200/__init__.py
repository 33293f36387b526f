"""ctypes binding of the C ABI (include/cuda_audio_b200.h) of the B200 convolution engine.

This is plumbing for tests/ and bench.py: the product is libcuda_audio_b200.so (CUDA kernels +
host runtime) and the C++ `Convolution` mirror under cuda-audio_b200/host/.  There is no CPU
fallback: if the shared library is missing or no CUDA device is usable, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(_PKG, "..", ".."))          # cuda-audio_b200/
LIB_PATH = os.environ.get("CA_B200_LIB") or os.path.join(ROOT, "libcuda_audio_b200.so")  # override: development A/B builds

CA_MAX_TIERS = 4
FLAG_GRAPH, FLAG_STREAMING, FLAG_L2_PERSIST, FLAG_PROFILE, FLAG_RAW_WET = 1, 2, 4, 8, 16
FLAG_ASYNC_TIERS = 32
FLAG_LEGACY_FFT = 64
FLAG_PERSISTENT = 128
FLAG_REF_QUIRKS = 256
SCHED_MAC_PERSISTENT, SCHED_MAC_PER_ITEM, SCHED_FUSED_TIER0, SCHED_NO_FUSED_TIER0, SCHED_PIPELINED, SCHED_NO_PDL, SCHED_ROWS8 = 1, 2, 4, 8, 16, 32, 64

EXPORTS = [
    "ca_api_version", "ca_strerror", "ca_last_error_string", "ca_config_init", "ca_config_auto_tiers", "ca_create", "ca_destroy",
    "ca_load_ir", "ca_load_ir_device", "ca_load_ir_interleaved_device", "ca_set_params", "ca_get_params", "ca_set_glide", "ca_set_active", "ca_reset",
    "ca_process", "ca_process_device", "ca_sync", "ca_stream", "ca_get_stats", "ca_reset_stats",
    "ca_host_alloc", "ca_host_free", "ca_measure_read_gbs", "ca_persist_stamps",
    "ca_group_config_init", "ca_group_create", "ca_group_destroy", "ca_group_load_ir", "ca_group_set_params", "ca_group_set_glide", "ca_group_reset",
    "ca_group_process", "ca_group_get_stats", "ca_group_reset_stats",
]
EXCHANGE_P2P, EXCHANGE_NCCL = 0, 1


class CaError(RuntimeError):
    def __init__(self, code, what, detail=""):
        super().__init__(f"{what}: error {code} ({detail})")
        self.code = code


class Config(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32), ("period", C.c_uint32),
                ("n_instances", C.c_uint32), ("n_in", C.c_uint32), ("n_out", C.c_uint32),
                ("max_ir_frames", C.c_uint32), ("n_ir_slots", C.c_uint32), ("flags", C.c_uint32),
                ("mac_split", C.c_uint32), ("max_voices", C.c_uint32), ("part_begin", C.c_uint32), ("part_count", C.c_uint32),
                ("n_tiers", C.c_uint32), ("tier_block", C.c_uint32 * CA_MAX_TIERS),
                ("tier_parts", C.c_uint32 * CA_MAX_TIERS), ("sample_rate", C.c_float), ("voice_pool", C.c_uint32), ("schedule", C.c_uint32), ("io_chunks", C.c_uint32), ("sm_split", C.c_uint32), ("ref_fft_size", C.c_uint32)]


class Params(C.Structure):
    """== Convolution::CC::value (conv.h:40-50)"""
    _fields_ = [("select", C.c_uint32), ("predelay", C.c_uint32), ("speed", C.c_uint32), ("vsteps", C.c_int32),
                ("dry", C.c_float), ("wet", C.c_float), ("panDry", C.c_float), ("panWet", C.c_float),
                ("level", C.c_float)]


class Stats(C.Structure):
    _fields_ = [("periods", C.c_uint64), ("xruns", C.c_uint64), ("mean_us", C.c_double), ("p50_us", C.c_double),
                ("p99_us", C.c_double), ("max_us", C.c_double), ("fwd_us", C.c_double), ("mac_us", C.c_double),
                ("inv_us", C.c_double), ("tiers_us", C.c_double), ("total_us", C.c_double), ("tier_fwd_us", C.c_double),
                ("tier_mac_us", C.c_double), ("tier_inv_us", C.c_double), ("gpu_launches", C.c_uint64),
                ("mac_bytes", C.c_uint64), ("mac_bytes_amortized", C.c_uint64), ("partitions", C.c_uint32),
                ("mac_split", C.c_uint32), ("device_bytes", C.c_uint64), ("n_tiers", C.c_uint32),
                ("tier_block", C.c_uint32 * CA_MAX_TIERS), ("tier_parts", C.c_uint32 * CA_MAX_TIERS),
                ("tier_offset", C.c_uint32 * CA_MAX_TIERS), ("tier0_fused", C.c_uint32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class GroupConfig(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("n_devices", C.c_uint32), ("devices", C.c_int32 * 8), ("period", C.c_uint32),
                ("n_in", C.c_uint32), ("n_out", C.c_uint32), ("max_ir_frames", C.c_uint32), ("n_ir_slots", C.c_uint32),
                ("flags", C.c_uint32), ("exchange", C.c_uint32), ("max_voices", C.c_uint32), ("sample_rate", C.c_float)]


class GroupStats(C.Structure):
    _fields_ = [("periods", C.c_uint64), ("mean_us", C.c_double), ("p50_us", C.c_double), ("p99_us", C.c_double), ("max_us", C.c_double),
                ("n_devices", C.c_uint32), ("exchange", C.c_uint32), ("part_begin", C.c_uint32 * 8), ("part_count", C.c_uint32 * 8),
                ("mac_split", C.c_uint32 * 8), ("mac_bytes", C.c_uint64 * 8), ("exchange_bytes_per_peer", C.c_uint64),
                ("gpu_launches", C.c_uint64), ("peer_timeout", C.c_int32)]


def build(force: bool = False) -> str:
    """Compile libcuda_audio_b200.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(ROOT, "csrc", f) for f in os.listdir(os.path.join(ROOT, "csrc"))]
    srcs.append(os.path.join(ROOT, "..", "include", "cuda_audio_b200.h"))
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", ROOT] + (["-B"] if force else []), stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CaError(-2, "libcuda_audio_b200.so is not built (run __graft_entry__.build()); no CPU fallback exists")
        L = C.CDLL(LIB_PATH)
        f32p = C.POINTER(C.c_float)
        vp = C.c_void_p
        L.ca_api_version.restype = C.c_int
        L.ca_strerror.restype = C.c_char_p
        L.ca_strerror.argtypes = [C.c_int]
        L.ca_last_error_string.restype = C.c_char_p
        L.ca_config_init.argtypes = [C.POINTER(Config)]
        L.ca_config_auto_tiers.argtypes = [C.POINTER(Config), C.c_uint32, C.c_uint32]
        L.ca_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
        L.ca_destroy.argtypes = [vp]
        L.ca_load_ir.argtypes = [vp, C.c_uint32, f32p, f32p, C.c_uint32]
        L.ca_load_ir_device.argtypes = [vp, C.c_uint32, vp, vp, C.c_uint32]
        L.ca_load_ir_interleaved_device.argtypes = [vp, C.c_uint32, vp, C.c_uint32]
        L.ca_set_params.argtypes = [vp, C.c_uint32, C.c_uint32, C.POINTER(Params)]
        L.ca_get_params.argtypes = [vp, C.c_uint32, C.c_uint32, C.POINTER(Params)]
        L.ca_set_glide.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_float]
        L.ca_set_active.argtypes = [vp, C.c_uint32]
        L.ca_reset.argtypes = [vp]
        L.ca_process.argtypes = [vp, vp, vp, C.c_uint32]
        L.ca_process_device.argtypes = [vp, vp, vp, C.c_uint32]
        L.ca_sync.argtypes = [vp]
        L.ca_stream.restype = vp
        L.ca_stream.argtypes = [vp]
        L.ca_get_stats.argtypes = [vp, C.POINTER(Stats)]
        L.ca_reset_stats.argtypes = [vp]
        L.ca_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
        L.ca_host_free.argtypes = [vp]
        L.ca_measure_read_gbs.argtypes = [C.c_int, C.c_size_t, C.c_int, C.POINTER(C.c_double)]
        L.ca_persist_stamps.argtypes = [vp, C.POINTER(C.c_uint64)]
        L.ca_group_config_init.argtypes = [C.POINTER(GroupConfig)]
        L.ca_group_create.argtypes = [C.POINTER(GroupConfig), C.POINTER(vp)]
        L.ca_group_destroy.argtypes = [vp]
        L.ca_group_load_ir.argtypes = [vp, C.c_uint32, f32p, f32p, C.c_uint32]
        L.ca_group_set_params.argtypes = [vp, C.c_uint32, C.POINTER(Params)]
        L.ca_group_set_glide.argtypes = [vp, C.c_uint32, C.c_float]
        L.ca_group_reset.argtypes = [vp]
        L.ca_group_process.argtypes = [vp, vp, vp, C.c_uint32]
        L.ca_group_get_stats.argtypes = [vp, C.POINTER(GroupStats)]
        L.ca_group_reset_stats.argtypes = [vp]
        _lib = L
    return _lib


def _check(rc, what):
    if rc != 0:
        L = lib()
        raise CaError(rc, what, f"{L.ca_strerror(rc).decode()}; {L.ca_last_error_string().decode()}")


def default_config(**kw) -> Config:
    cfg = Config()
    lib().ca_config_init(C.byref(cfg))
    for k, v in kw.items():
        if k in ("tier_block", "tier_parts"):
            for i, x in enumerate(v):
                getattr(cfg, k)[i] = x
        else:
            setattr(cfg, k, v)
    return cfg


def measure_read_gbs(nbytes: int, iters: int = 20, device: int = 0) -> float:
    g = C.c_double()
    _check(lib().ca_measure_read_gbs(device, nbytes, iters, C.byref(g)), "ca_measure_read_gbs")
    return g.value


class PinnedArray:
    """float32 numpy view over cudaMallocHost memory (so ca_process needs no staging copy)."""

    def __init__(self, shape):
        self.shape = tuple(shape)
        n = int(np.prod(self.shape))
        p = C.c_void_p()
        _check(lib().ca_host_alloc(C.byref(p), n * 4), "ca_host_alloc")
        self.ptr = p.value
        self.array = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(n,)).reshape(self.shape)

    def free(self):
        if self.ptr:
            lib().ca_host_free(self.ptr)
            self.ptr = None
            self.array = None


class Engine:
    """One engine = n_instances batched convolution instances (one reference `Convolution` each)."""

    def __init__(self, period=256, max_ir_frames=130048, n_instances=1, n_in=2, n_out=2, n_ir_slots=2, device=0,
                 flags=0, mac_split=0, part_begin=0, part_count=0, sample_rate=48000.0, tiers=None, max_voices=0, tier_growth=0,
                 tier_max_block=0, voice_pool=0, schedule=0, io_chunks=0, sm_split=0, ref_fft_size=0):
        kw = dict(period=period, max_ir_frames=max_ir_frames, n_instances=n_instances, n_in=n_in, n_out=n_out,
                  n_ir_slots=n_ir_slots, device=device, flags=flags, mac_split=mac_split, part_begin=part_begin,
                  part_count=part_count, sample_rate=sample_rate, max_voices=max_voices, voice_pool=voice_pool, schedule=schedule,
                  io_chunks=io_chunks, sm_split=sm_split, ref_fft_size=ref_fft_size)
        if tiers and tiers != "auto":
            kw.update(n_tiers=len(tiers), tier_block=[t[0] for t in tiers], tier_parts=[t[1] for t in tiers])
        self.cfg = default_config(**kw)
        if tiers == "auto":
            _check(lib().ca_config_auto_tiers(C.byref(self.cfg), tier_growth, tier_max_block), "ca_config_auto_tiers")
        h = C.c_void_p()
        _check(lib().ca_create(C.byref(self.cfg), C.byref(h)), "ca_create")
        self._h = h
        self.B, self.n_in, self.n_out, self.n_inst = period, n_in, n_out, n_instances
        self.n_active = n_instances

    def close(self):
        if getattr(self, "_h", None):
            lib().ca_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # --- IR bank ---
    def load_ir(self, slot, left, right=None):
        left = np.ascontiguousarray(left, np.float32)
        f32p = C.POINTER(C.c_float)
        rp = None
        if right is not None:
            right = np.ascontiguousarray(right, np.float32)
            assert len(right) == len(left)
            rp = right.ctypes.data_as(f32p)
        elif self.n_out == 2:
            rp = left.ctypes.data_as(f32p)
        _check(lib().ca_load_ir(self._h, slot, left.ctypes.data_as(f32p), rp, len(left)), "ca_load_ir")

    def load_ir_device(self, slot, d_left: int, d_right: int, frames: int):
        _check(lib().ca_load_ir_device(self._h, slot, d_left, d_right, frames), "ca_load_ir_device")

    # --- parameters ---
    def set_params(self, instance, inp, select=0, predelay=0, speed=100, vsteps=-1, dry=0.5, wet=0.5, panDry=0.0,
                   panWet=0.0, level=1.0):
        p = Params(select, predelay, speed, vsteps, dry, wet, panDry, panWet, level)
        _check(lib().ca_set_params(self._h, instance, inp, C.byref(p)), "ca_set_params")

    def get_params(self, instance, inp) -> Params:
        p = Params()
        _check(lib().ca_get_params(self._h, instance, inp, C.byref(p)), "ca_get_params")
        return p

    def set_glide(self, instance, inp, g):
        _check(lib().ca_set_glide(self._h, instance, inp, g), "ca_set_glide")

    def reset(self):
        """every active instance restarts like a new one (history dropped, wet glide from silence)"""
        _check(lib().ca_reset(self._h), "ca_reset")

    def set_active(self, n):
        _check(lib().ca_set_active(self._h, n), "ca_set_active")
        self.n_active = n

    # --- processing ---
    def process(self, x: np.ndarray) -> np.ndarray:
        """x: [n_active][n_in][B] float32 (host) -> [n_active][n_out][B]"""
        x = np.ascontiguousarray(x, np.float32)
        assert x.shape == (self.n_active, self.n_in, self.B), x.shape
        out = np.empty((self.n_active, self.n_out, self.B), np.float32)
        _check(lib().ca_process(self._h, x.ctypes.data, out.ctypes.data, self.B), "ca_process")
        return out

    def process_raw(self, in_ptr: int, out_ptr: int):
        _check(lib().ca_process(self._h, in_ptr, out_ptr, self.B), "ca_process")

    def process_device(self, d_in: int, d_out: int):
        _check(lib().ca_process_device(self._h, d_in, d_out, self.B), "ca_process_device")

    def sync(self):
        _check(lib().ca_sync(self._h), "ca_sync")

    @property
    def stream(self) -> int:
        return lib().ca_stream(self._h)

    def render(self, x: np.ndarray) -> np.ndarray:
        """x: [n_active][n_in][n] -> [n_active][n_out][n] (n truncated to whole periods)."""
        x = np.ascontiguousarray(x, np.float32)
        n = (x.shape[-1] // self.B) * self.B
        out = np.empty((self.n_active, self.n_out, n), np.float32)
        for t in range(n // self.B):
            out[:, :, t * self.B:(t + 1) * self.B] = self.process(x[:, :, t * self.B:(t + 1) * self.B])
        return out

    def stats(self) -> Stats:
        s = Stats()
        _check(lib().ca_get_stats(self._h, C.byref(s)), "ca_get_stats")
        return s

    def reset_stats(self):
        _check(lib().ca_reset_stats(self._h), "ca_reset_stats")

    def persist_stamps(self):
        a = (C.c_uint64 * 8)()
        _check(lib().ca_persist_stamps(self._h, a), "ca_persist_stamps")
        return list(a)


class Group:
    """One very long IR split by partition range across the GPUs of one node (ca_group, BASELINE configs[4])."""

    def __init__(self, devices, period=256, max_ir_frames=2880000, n_in=2, n_out=2, n_ir_slots=2, flags=0, exchange=EXCHANGE_P2P,
                 max_voices=0, sample_rate=48000.0):
        cfg = GroupConfig()
        lib().ca_group_config_init(C.byref(cfg))
        devices = list(devices)
        cfg.n_devices = len(devices)
        for i, d in enumerate(devices):
            cfg.devices[i] = d
        cfg.period, cfg.n_in, cfg.n_out, cfg.max_ir_frames, cfg.n_ir_slots = period, n_in, n_out, max_ir_frames, n_ir_slots
        cfg.flags, cfg.exchange, cfg.max_voices, cfg.sample_rate = flags, exchange, max_voices, sample_rate
        self.cfg = cfg
        h = C.c_void_p()
        _check(lib().ca_group_create(C.byref(cfg), C.byref(h)), "ca_group_create")
        self._h = h
        self.B, self.n_in, self.n_out = period, n_in, n_out

    def close(self):
        if getattr(self, "_h", None):
            lib().ca_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def load_ir(self, slot, left, right=None):
        left = np.ascontiguousarray(left, np.float32)
        right = left if right is None else np.ascontiguousarray(right, np.float32)
        f32p = C.POINTER(C.c_float)
        _check(lib().ca_group_load_ir(self._h, slot, left.ctypes.data_as(f32p), right.ctypes.data_as(f32p), len(left)), "ca_group_load_ir")

    def set_params(self, inp, select=0, predelay=0, speed=100, vsteps=-1, dry=0.5, wet=0.5, panDry=0.0, panWet=0.0, level=1.0):
        p = Params(select, predelay, speed, vsteps, dry, wet, panDry, panWet, level)
        _check(lib().ca_group_set_params(self._h, inp, C.byref(p)), "ca_group_set_params")

    def set_glide(self, inp, g):
        _check(lib().ca_group_set_glide(self._h, inp, g), "ca_group_set_glide")

    def process(self, x: np.ndarray) -> np.ndarray:
        """x: [n_in][B] float32 -> [n_out][B]"""
        x = np.ascontiguousarray(x, np.float32)
        assert x.shape == (self.n_in, self.B), x.shape
        out = np.empty((self.n_out, self.B), np.float32)
        _check(lib().ca_group_process(self._h, x.ctypes.data, out.ctypes.data, self.B), "ca_group_process")
        return out

    def process_raw(self, in_ptr: int, out_ptr: int):
        _check(lib().ca_group_process(self._h, in_ptr, out_ptr, self.B), "ca_group_process")

    def render(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, np.float32)
        n = (x.shape[-1] // self.B) * self.B
        out = np.empty((self.n_out, n), np.float32)
        for t in range(n // self.B):
            out[:, t * self.B:(t + 1) * self.B] = self.process(x[:, t * self.B:(t + 1) * self.B])
        return out

    def stats(self) -> GroupStats:
        s = GroupStats()
        _check(lib().ca_group_get_stats(self._h, C.byref(s)), "ca_group_get_stats")
        return s

    def reset(self):
        _check(lib().ca_group_reset(self._h), "ca_group_reset")

    def reset_stats(self):
        _check(lib().ca_group_reset_stats(self._h), "ca_group_reset_stats")
