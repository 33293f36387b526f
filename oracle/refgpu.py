"""ctypes front-end for oracle/_ref/libref_conv.so: the UNMODIFIED reference (conv.cu / wav.cu,
cuFFT path) compiled by oracle/ref_harness/Makefile.  TEST INFRASTRUCTURE: used by the -m gpu
parity tests, tests/golden/make_golden.py and `bench.py --impl reference`.  Needs a GPU to run.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libref_conv.so")

_libs = {}


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib(path: str = None):
    """The harness C API.  Default: the compiled reference; tests/dropin builds the same harness
    against the B200 engine's host mirror and passes that path instead."""
    path = path or LIB_PATH
    if path not in _libs:
        L = C.CDLL(path)
        f32p = C.POINTER(C.c_float)
        L.ref_device_count.restype = C.c_int
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.c_size_t, C.c_int]
        L.ref_prepare.argtypes = [C.c_void_p, C.c_size_t, f32p, f32p, C.c_size_t, C.c_size_t]
        L.ref_prepare_wav.argtypes = [C.c_void_p, C.c_size_t, C.c_char_p, C.c_size_t]
        L.ref_wav_decode.restype = C.c_long
        L.ref_wav_decode.argtypes = [C.c_char_p, f32p, f32p, C.c_size_t]
        L.ref_set_cc.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.c_long] + [C.c_float] * 5
        L.ref_midi_cc.argtypes = [C.c_void_p, C.c_int, C.c_uint8, C.c_uint8]
        L.ref_get_cc.argtypes = [C.c_void_p, C.c_int] + [C.POINTER(C.c_size_t)] * 4 + [f32p] * 5
        L.ref_process.argtypes = [C.c_void_p, f32p, f32p, f32p, f32p, C.c_size_t]
        L.ref_render.argtypes = [C.c_void_p, f32p, f32p, f32p, f32p, C.c_size_t, C.c_size_t]
        L.ref_avg_runtime_ms.restype = C.c_double
        L.ref_avg_runtime_ms.argtypes = [C.c_void_p]
        L.ref_bench.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_double)]
        _libs[path] = L
    return _libs[path]


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class RefGpu:
    """One reference `Convolution` object (conv.h:30-86) on the GPU."""

    def __init__(self, fft_size: int, device: int = 0, lib_path: str = None):
        self.N = fft_size
        self._lib = lib(lib_path)
        self._h = self._lib.ref_create(fft_size, device)
        if not self._h:
            raise RuntimeError("reference harness: no CUDA device")
        self._cc = [dict(select=0, predelay=0, speed=100, vsteps=-1, dry=0.5, wet=0.5, panDry=0.0, panWet=0.0, level=1.0)
                    for _ in range(2)]

    def prepare(self, idx, left, right, nframes=1024):
        left, right = _f32(left), _f32(right)
        rc = self._lib.ref_prepare(self._h, idx, _p(left), _p(right), len(left), nframes)
        assert rc == 0, rc

    def prepare_wav(self, idx, path, nframes=1024) -> int:
        return self._lib.ref_prepare_wav(self._h, idx, path.encode(), nframes)

    def set_cc(self, i, **kw):
        self._cc[i].update(kw)
        c = self._cc[i]
        self._lib.ref_set_cc(self._h, i, c["select"], c["predelay"], c["speed"], c["vsteps"], c["dry"], c["wet"],
                         c["panDry"], c["panWet"], c["level"])
        self._cc[i]["vsteps"] = -1

    def midi_cc(self, i, which, value):
        """which: 1 select, 2 predelay, 3 dry, 4 wet, 5 speed, 6 panDry, 7 panWet, 8 level"""
        self._lib.ref_midi_cc(self._h, i, which, value)

    def get_cc(self, i) -> dict:
        s = [C.c_size_t() for _ in range(4)]
        f = [C.c_float() for _ in range(5)]
        self._lib.ref_get_cc(self._h, i, *[C.byref(x) for x in s], *[C.byref(x) for x in f])
        return dict(select=s[0].value, predelay=s[1].value, speed=s[2].value, vsteps=s[3].value, dry=f[0].value,
                    wet=f[1].value, panDry=f[2].value, panWet=f[3].value, level=f[4].value)

    def process(self, in1, in2):
        in1, in2 = _f32(in1), _f32(in2)
        n = len(in1)
        L = np.empty(n, np.float32)
        R = np.empty(n, np.float32)
        self._lib.ref_process(self._h, _p(in1), _p(in2), _p(L), _p(R), n)
        return L, R

    def render(self, x1, x2, B):
        x1, x2 = _f32(x1), _f32(x2)
        periods = len(x1) // B
        L = np.empty(periods * B, np.float32)
        R = np.empty(periods * B, np.float32)
        self._lib.ref_render(self._h, _p(x1), _p(x2), _p(L), _p(R), B, periods)
        return L, R

    def avg_runtime_ms(self) -> float:
        return self._lib.ref_avg_runtime_ms(self._h)


def wav_decode(path: str, max_frames: int = 1 << 22, lib_path: str = None):
    L = np.empty(max_frames, np.float32)
    R = np.empty(max_frames, np.float32)
    n = lib(lib_path).ref_wav_decode(path.encode(), _p(L), _p(R), max_frames)
    n = min(n, max_frames)
    return L[:n].copy(), R[:n].copy()


def bench(instances, nframes: int, warmup: int, periods: int) -> np.ndarray:
    """K reference instances on K host threads, lock-stepped per period; returns wall us per period."""
    K = len(instances)
    arr = (C.c_void_p * K)(*[i._h for i in instances])
    wall = np.empty(periods, np.float64)
    instances[0]._lib.ref_bench(arr, K, nframes, warmup, periods, wall.ctypes.data_as(C.POINTER(C.c_double)))
    return wall
