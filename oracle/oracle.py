"""ctypes front-end for the CPU oracles (TEST INFRASTRUCTURE, not product code).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs import this module.  It wraps:

* ``RefConv``   -- oracle/refconv.c, the literal fp64 restatement of the reference's
                  ``Convolution`` (src/conv.cu:142-466), quirks included.
* ``Upols``     -- oracle/upols_cpu.c, the fp32 CPU uniform-partitioned overlap-save port
                  (bench cpu_baseline, kind "port").
* ``direct_conv`` / ``direct_conv_at`` -- fp64 time-domain convolution (ground truth).
* ``engine_truth`` -- the new-engine formula of SURVEY.md section 8a evaluated in fp64 with
                  FFT or direct convolution (static parameters).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("refconv.c", "upols_cpu.c")]
    stale = (not os.path.exists(_LIB)) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        f32p, f64p = C.POINTER(C.c_float), C.POINTER(C.c_double)
        L.refconv_create.restype = C.c_void_p
        L.refconv_create.argtypes = [C.c_size_t]
        L.refconv_destroy.argtypes = [C.c_void_p]
        L.refconv_set_cc.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.refconv_get_cc.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.refconv_prepare.argtypes = [C.c_void_p, C.c_size_t, f32p, f32p, C.c_size_t, C.c_size_t]
        L.refconv_process.argtypes = [C.c_void_p, f32p, f32p, f32p, f32p, C.c_size_t]
        L.refconv_handle_cc.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t]
        L.refconv_pcm16_to_float.argtypes = [C.c_void_p, f32p, C.c_size_t]
        L.refconv_pcm24_to_float.argtypes = [C.c_void_p, f32p, C.c_size_t]
        L.oracle_direct_conv_f64.argtypes = [f64p, C.c_size_t, f64p, C.c_size_t, f64p, C.c_size_t]
        L.oracle_direct_conv_f64_at.argtypes = [f64p, C.c_size_t, f64p, C.c_size_t, C.POINTER(C.c_int64), C.c_size_t, f64p]
        L.upols_create.restype = C.c_void_p
        L.upols_create.argtypes = [C.c_int] * 6
        L.upols_destroy.argtypes = [C.c_void_p]
        L.upols_P.argtypes = [C.c_void_p]
        L.upols_load_ir.argtypes = [C.c_void_p, C.c_int, f32p, f32p, C.c_int]
        L.upols_set_param.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_float] * 5 + [C.c_uint32, C.c_uint32, C.c_int32]
        L.upols_set_glide.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float]
        L.upols_process.argtypes = [C.c_void_p, f32p, f32p, C.c_int]
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t=C.c_float):
    return a.ctypes.data_as(C.POINTER(t))


class CC(C.Structure):
    """Mirror of Convolution::CC::value (conv.h:40-50) as refconv_cc."""
    _fields_ = [("select", C.c_size_t), ("predelay", C.c_size_t), ("speed", C.c_size_t), ("vsteps", C.c_size_t),
                ("dry", C.c_double), ("wet", C.c_double), ("panDry", C.c_double), ("panWet", C.c_double),
                ("level", C.c_double)]


class RefConv:
    """fp64 restatement of the reference Convolution (conv.cu)."""

    def __init__(self, fft_size: int):
        self.N = fft_size
        self._h = lib().refconv_create(fft_size)

    def close(self):
        if self._h:
            lib().refconv_destroy(self._h)
            self._h = None

    __del__ = close

    def set_cc(self, i, **kw):
        cc = CC()
        lib().refconv_get_cc(self._h, i, C.byref(cc))
        for k, v in kw.items():
            setattr(cc, k, v)
        lib().refconv_set_cc(self._h, i, C.byref(cc))

    def get_cc(self, i) -> CC:
        cc = CC()
        lib().refconv_get_cc(self._h, i, C.byref(cc))
        return cc

    def prepare(self, idx, left, right, nframes=1024):
        left, right = _f32(left), _f32(right)
        rc = lib().refconv_prepare(self._h, idx, _p(left), _p(right), len(left), nframes)
        assert rc == 0

    def process(self, in1, in2):
        in1, in2 = _f32(in1), _f32(in2)
        n = len(in1)
        L = np.empty(n, np.float32)
        R = np.empty(n, np.float32)
        rc = lib().refconv_process(self._h, _p(in1), _p(in2), _p(L), _p(R), n)
        assert rc == 0, rc
        return L, R

    def render(self, x1, x2, B):
        """Run len(x1)/B periods; returns (L, R)."""
        n = (len(x1) // B) * B
        L = np.empty(n, np.float32)
        R = np.empty(n, np.float32)
        for t in range(n // B):
            l, r = self.process(x1[t * B:(t + 1) * B], x2[t * B:(t + 1) * B])
            L[t * B:(t + 1) * B] = l
            R[t * B:(t + 1) * B] = r
        return L, R


def handle_cc(cc: CC, which: int, v: int, nb: int):
    lib().refconv_handle_cc(C.byref(cc), which, v, nb)


def pcm16_to_float(raw: np.ndarray) -> np.ndarray:
    raw = np.ascontiguousarray(raw, dtype=np.int16)
    out = np.empty(raw.size, np.float32)
    lib().refconv_pcm16_to_float(raw.ctypes.data, _p(out), raw.size)
    return out


def pcm24_to_float(raw_bytes: np.ndarray) -> np.ndarray:
    raw = np.ascontiguousarray(raw_bytes, dtype=np.uint8)
    out = np.empty(raw.size // 3, np.float32)
    lib().refconv_pcm24_to_float(raw.ctypes.data, _p(out), raw.size // 3)
    return out


def direct_conv(x, h, ny=None) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float64)
    h = np.ascontiguousarray(h, dtype=np.float64)
    ny = len(x) if ny is None else ny
    y = np.empty(ny, np.float64)
    lib().oracle_direct_conv_f64(_p(x, C.c_double), len(x), _p(h, C.c_double), len(h), _p(y, C.c_double), ny)
    return y


def direct_conv_at(x, h, idx) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float64)
    h = np.ascontiguousarray(h, dtype=np.float64)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    y = np.empty(len(idx), np.float64)
    lib().oracle_direct_conv_f64_at(_p(x, C.c_double), len(x), _p(h, C.c_double), len(h), _p(idx, C.c_int64), len(idx), _p(y, C.c_double))
    return y


def fft_conv(x, h, ny=None) -> np.ndarray:
    """fp64 FFT convolution (error ~1e-15) for sizes where direct conv is too slow."""
    x = np.asarray(x, np.float64)
    h = np.asarray(h, np.float64)
    ny = len(x) if ny is None else ny
    n = 1
    while n < len(x) + len(h):
        n *= 2
    y = np.fft.irfft(np.fft.rfft(x, n) * np.fft.rfft(h, n), n)
    return y[:ny]


def pan_gains(pan: float):
    """Pan law of conv.cu:386-389 / 418-421 -> (gainL, gainR)."""
    return (1 - pan if pan >= 0 else 1.0), (1 + pan if pan <= 0 else 1.0)


def engine_truth(x, irs, params, predelay=0, conv=fft_conv):
    """fp64 evaluation of the engine formula for STATIC parameters with the wet glide
    converged: out_o = clamp(sum_i panWet_io*level_i*wet_i*(x_i * h_io) delayed) + dry mix.

    x: [n_in][n] ; irs: [n_in][n_out][L] ; params: list per input of dict(wet,dry,level,panWet,panDry)
    """
    x = np.asarray(x, np.float64)
    n_in, n = x.shape
    n_out = len(irs[0])
    out = np.zeros((n_out, n))
    for o in range(n_out):
        wet = np.zeros(n)
        for i in range(n_in):
            p = params[i]
            pan = pan_gains(p.get("panWet", 0.0))[o] if n_out == 2 else 1.0
            y = conv(x[i], irs[i][o], n)
            wet += pan * p.get("level", 1.0) * p.get("wet", 1.0) * y
        if predelay:
            wet = np.concatenate([np.zeros(predelay), wet])[:n]
        wet = np.clip(wet, -1.0, 1.0)
        for i in range(n_in):
            p = params[i]
            pan = pan_gains(p.get("panDry", 0.0))[o] if n_out == 2 else 1.0
            wet = wet + p.get("dry", 0.0) * pan * p.get("level", 1.0) * x[i]
        out[o] = wet
    return out


class Upols:
    """fp32 CPU uniform-partitioned overlap-save port (cpu_baseline 'port')."""

    def __init__(self, B, max_ir_frames, n_in=2, n_out=2, n_inst=1, n_ir=2):
        self.B, self.n_in, self.n_out, self.n_inst = B, n_in, n_out, n_inst
        self._h = lib().upols_create(B, max_ir_frames, n_in, n_out, n_inst, n_ir)
        self.P = lib().upols_P(self._h)

    def close(self):
        if self._h:
            lib().upols_destroy(self._h)
            self._h = None

    __del__ = close

    def load_ir(self, slot, left, right=None):
        left = _f32(left)
        right = left if right is None else _f32(right)
        assert lib().upols_load_ir(self._h, slot, _p(left), _p(right), len(left)) == 0

    def set_param(self, inst, inp, wet=0.5, dry=0.5, level=1.0, panWet=0.0, panDry=0.0, predelay=0, select=0,
                  vsteps=-1, glide=None):
        lib().upols_set_param(self._h, inst, inp, wet, dry, level, panWet, panDry, predelay, select, vsteps)
        if glide is not None:
            lib().upols_set_glide(self._h, inst, inp, glide)

    def process(self, x):
        """x: [n_inst][n_in][B] -> [n_inst][n_out][B]"""
        x = _f32(x)
        out = np.empty((self.n_inst, self.n_out, self.B), np.float32)
        rc = lib().upols_process(self._h, _p(x), _p(out), self.B)
        assert rc == 0
        return out

    def render(self, x):
        """x: [n_inst][n_in][n] -> [n_inst][n_out][n]"""
        x = _f32(x)
        n = (x.shape[-1] // self.B) * self.B
        out = np.empty((self.n_inst, self.n_out, n), np.float32)
        for t in range(n // self.B):
            out[:, :, t * self.B:(t + 1) * self.B] = self.process(x[:, :, t * self.B:(t + 1) * self.B])
        return out


# ---------------------------------------------------------------------------------------------
# Synthetic signals of SURVEY.md section 8(d) (shared by tests, golden generation and bench)
# ---------------------------------------------------------------------------------------------
def synth_ir(frames: int, fs: float, seed: int, parity_safe: bool = True, dtype=np.float32) -> np.ndarray:
    """Exponentially-decaying Gaussian noise, T60 = 0.8 x length, unit energy.

    parity_safe=True applies the two linear corrections of SURVEY 8(c)(i):
    sum(h) = 0 and sum((-1)^n h) = 0, so the reference's DC / Nyquist quirks contribute nothing.
    """
    rng = np.random.default_rng(seed)
    n = np.arange(frames)
    t60 = 0.8 * frames / fs
    h = rng.standard_normal(frames) * np.exp(-6.91 * n / (t60 * fs))
    if parity_safe and frames >= 4:
        env = np.exp(-6.91 * n / (t60 * fs))
        alt = np.where(n % 2 == 0, 1.0, -1.0)
        # solve for a, b: h -= a*env + b*env*alt such that both sums vanish
        A = np.array([[env.sum(), (env * alt).sum()], [(env * alt).sum(), env.sum()]])
        rhs = np.array([h.sum(), (h * alt).sum()])
        a, b = np.linalg.solve(A, rhs)
        h = h - a * env - b * env * alt
    h = h / np.sqrt((h ** 2).sum())
    return h.astype(dtype)


def synth_audio(frames: int, seed: int, rms: float = 0.1, dtype=np.float32) -> np.ndarray:
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(frames) * rms
    return np.clip(x, -0.9, 0.9).astype(dtype)


def rel_l2(a, b) -> float:
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.sqrt(((a - b) ** 2).sum()) / max(np.sqrt((b ** 2).sum()), 1e-300))
