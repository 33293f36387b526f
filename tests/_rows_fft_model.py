"""numpy model of cuda-audio_b200/csrc/fft_rows.cuh: the 256-point row FFT (one warp per row, 8 points
per lane, 256 = 8 x 8 x 4, three shared-memory exchanges), the four-step decomposition M = M1 x 256 of the
long tiers, and the real-FFT split in POSITION order.  Every loop mirrors the kernel's thread/lane
mapping and shared-memory index functions one to one; tests/test_fft_model.py checks the result
against numpy and counts shared-memory bank conflicts per access.
TEST INFRASTRUCTURE (index-math model), not product code.
"""
import numpy as np

PAD = 4            # float2 slots of padding per 32 (exchange layouts)


def W(N, k, inv=False):
    return np.exp((2j if inv else -2j) * np.pi * (k % N) / N)


# ---- shared-memory index functions (float2 units) of the three exchanges ----
def e1(q, lane):             # after stage 1: value t_q[lane]
    return q * (32 + PAD) + lane


def e2(q, l0, r0):           # after stage 2: value g_{q,l0}[r0]; l0 is the slow index
    return l0 * (64 + PAD) + r0 * 8 + q


def e3(k):                   # natural order, padded every 32
    return k + (k >> 5) * PAD


ROW_SLOTS = 256 + 8 * PAD    # float2 slots of one row region


class Conflicts:
    """bank-conflict counter for 8-byte accesses: a warp access is served half-warp by half-warp; within a
    half-warp two lanes conflict when they touch different addresses in the same bank pair."""

    def __init__(self):
        self.worst = {}

    def access(self, name, idx_per_lane):
        worst = 1
        for h in range(2):
            lanes = idx_per_lane[16 * h:16 * h + 16]
            banks = {}
            for a in lanes:
                banks.setdefault(a % 16, set()).add(a)
            worst = max(worst, max(len(v) for v in banks.values()))
        self.worst[name] = max(self.worst.get(name, 1), worst)


def dft(v, inv):
    n = len(v)
    k = np.arange(n)
    return np.array([(v * W(n, k * q, inv)).sum() for q in range(n)])


def row_fft(x, inv=False, cf=None):
    """x: 256 complex in natural order -> 256 complex in natural order (unnormalised DFT / inverse DFT).
    Registers: lane holds v[b] = x[lane + 32 b]."""
    cf = cf or Conflicts()
    sm = np.zeros(ROW_SLOTS, complex)
    # stage 1: radix 8 over b, twiddle W_256^(lane q)
    for q in range(8):
        cf.access("e1.write", [e1(q, lane) for lane in range(32)])
    for lane in range(32):
        v = np.array([x[lane + 32 * b] for b in range(8)])
        y = dft(v, inv)
        for q in range(8):
            sm[e1(q, lane)] = y[q] * W(256, lane * q, inv)
    # stage 2: thread (q, l0) = (lane >> 2, lane & 3): radix 8 over l1, twiddle W_32^(l0 r0)
    for l1 in range(8):
        cf.access("e1.read", [e1(lane >> 2, (lane & 3) + 4 * l1) for lane in range(32)])
    regs = {}
    for lane in range(32):
        q, l0 = lane >> 2, lane & 3
        u = np.array([sm[e1(q, l0 + 4 * l1)] for l1 in range(8)])
        g = dft(u, inv)
        regs[lane] = [g[r0] * W(32, l0 * r0, inv) for r0 in range(8)]
    sm2 = np.zeros(ROW_SLOTS, complex)
    for r0 in range(8):
        cf.access("e2.write", [e2(lane >> 2, lane & 3, r0) for lane in range(32)])
    for lane in range(32):
        q, l0 = lane >> 2, lane & 3
        for r0 in range(8):
            sm2[e2(q, l0, r0)] = regs[lane][r0]
    # stage 3: thread handles (q, r0) groups g = lane and lane + 32 (g = r0 * 8 + q): radix 4 over l0
    for half in range(2):
        for l0 in range(4):
            idx = []
            for lane in range(32):
                g = lane + 32 * half
                idx.append(e2(g & 7, l0, g >> 3))
            cf.access("e2.read", idx)
    out = np.zeros(ROW_SLOTS, complex)
    for half in range(2):
        for r1 in range(4):
            idx = []
            for lane in range(32):
                g = lane + 32 * half
                idx.append(e3((g & 7) + 8 * (g >> 3) + 64 * r1))
            cf.access("e3.write", idx)
    for lane in range(32):
        for half in range(2):
            g = lane + 32 * half
            q, r0 = g & 7, g >> 3
            u = np.array([sm2[e2(q, l0, r0)] for l0 in range(4)])
            z = dft(u, inv)
            for r1 in range(4):
                out[e3(q + 8 * r0 + 64 * r1)] = z[r1]
    return np.array([out[e3(k)] for k in range(256)]), cf


def brev(x, bits):
    r = 0
    for i in range(bits):
        r |= ((x >> i) & 1) << (bits - 1 - i)
    return r


def zpos(k, s):
    """position of bin k in the long tiers' spectra (same as fft_cta.cuh): row = bitrev_s(k mod 2^s), column = k >> s"""
    return (brev(k & ((1 << s) - 1), s) << 8) | (k >> s) if s else k


def four_step_forward(z, s):
    """z: M = 256 * 2^s complex (natural order) -> spectrum in POSITION order.
    columns: A[k1][n2] = sum_n1 z[256 n1 + n2] W_M1^(n1 k1); rows: X[k1 + M1 k2] = sum_n2 A[k1][n2] W_M^(n2 k1) W_256^(n2 k2)."""
    M1 = 1 << s
    M = 256 * M1
    A = np.zeros((M1, 256), complex)
    for n2 in range(256):
        A[:, n2] = dft(z[n2::256], False)
    out = np.zeros(M, complex)
    for k1 in range(M1):
        lane = np.arange(256)
        row = A[k1] * W(M, lane * k1)
        X, _ = row_fft(row)
        out[(brev(k1, s) << 8):(brev(k1, s) << 8) + 256] = X
    return out


def four_step_inverse(Zp, s):
    """position-order spectrum -> natural-order time samples (unnormalised inverse DFT)."""
    M1 = 1 << s
    M = 256 * M1
    A = np.zeros((M1, 256), complex)
    for k1 in range(M1):
        row, _ = row_fft(Zp[(brev(k1, s) << 8):(brev(k1, s) << 8) + 256], inv=True)
        A[k1] = row * W(M, np.arange(256) * k1, inv=True)
    z = np.zeros(M, complex)
    for n2 in range(256):
        z[n2::256] = dft(A[:, n2], True)
    return z


def split_r2c_pos(Zp, s):
    """real-FFT split in position order: row k1 pairs with row (M1 - k1) % M1, column k2 with 255 - k2
    (k1 != 0) or (256 - k2) % 256 (k1 == 0); position 0 holds (DC, Nyquist)."""
    M1 = 1 << s
    M = 256 * M1
    X = np.zeros(M, complex)
    for k1 in range(M1):
        for k2 in range(256):
            k = k1 + M1 * k2
            if k == 0:
                z = Zp[0]
                X[0] = complex(z.real + z.imag, z.real - z.imag)
                continue
            kp1 = (M1 - k1) % M1
            kp2 = 255 - k2 if k1 else 256 - k2
            z = Zp[(brev(k1, s) << 8) | k2]
            zp = Zp[(brev(kp1, s) << 8) | kp2]
            w = W(2 * M, k)
            e = 0.5 * (z + np.conj(zp))
            m = (z - np.conj(zp)) * w
            X[(brev(k1, s) << 8) | k2] = e - 0.5j * m
    return X


def split_c2r_pos(Yp, s):
    M1 = 1 << s
    M = 256 * M1
    Z = np.zeros(M, complex)
    for k1 in range(M1):
        for k2 in range(256):
            k = k1 + M1 * k2
            if k == 0:
                y = Yp[0]
                Z[0] = complex(y.real + y.imag, y.real - y.imag)
                continue
            kp1 = (M1 - k1) % M1
            kp2 = 255 - k2 if k1 else 256 - k2
            y = Yp[(brev(k1, s) << 8) | k2]
            yp = Yp[(brev(kp1, s) << 8) | kp2]
            w = W(2 * M, k)
            Z[(brev(k1, s) << 8) | k2] = (y + np.conj(yp)) + 1j * (y - np.conj(yp)) * np.conj(w)
    return Z
