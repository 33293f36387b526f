"""numpy model of cuda-audio_b200/csrc/fft_rows16.cuh / kernels_rows16.cuh: the 256-point row FFT as 16 x 16 (two
rows per warp, one per half-warp, 16 points per lane, ONE shared-memory exchange) and the real-FFT split through
register shuffles.  Every loop mirrors the kernel's lane / register mapping one to one; tests/test_fft_model.py
checks the result against numpy and counts shared-memory bank conflicts of the exchange.
TEST INFRASTRUCTURE (index-math model), not product code.
"""
import numpy as np

from tests._rows_fft_model import W, dft

PITCH = 18                 # float2 per exchange row
SLOTS = 16 * PITCH         # float2 slots of one half-warp's region


def conflicts_8B(idx16):
    """worst multiplicity of an 8-byte access by 16 lanes (one wavefront): lanes conflict when they touch different
    addresses in the same bank pair (16 pairs of 4-byte banks)"""
    banks = {}
    for a in idx16:
        banks.setdefault(a % 16, set()).add(a)
    return max(len(v) for v in banks.values())


def conflicts_16B(idx8):
    """16-byte access by 8 lanes (one wavefront), index in float2 units (even)"""
    banks = {}
    for a in idx8:
        banks.setdefault((a // 2) % 8, set()).add(a)
    return max(len(v) for v in banks.values())


def fft256x2(x, inv=False):
    """x: 256 complex (one row) -> (regs, worst): regs[l][p] = X[l + 16 p]; lane l holds v[b] = x[l + 16 b]"""
    worst = 1
    S = np.zeros(SLOTS, complex)
    for q in range(16):
        worst = max(worst, conflicts_8B([q * PITCH + l for l in range(16)]))
    for l in range(16):
        y = dft(np.array([x[l + 16 * b] for b in range(16)]), inv)
        for q in range(16):
            S[q * PITCH + l] = y[q] * W(256, l * q, inv)      # table entry [q * 16 + l]
    regs = []
    for i in range(8):
        for half in range(2):
            worst = max(worst, conflicts_16B([l * PITCH + 2 * i for l in range(8 * half, 8 * half + 8)]))
    for l in range(16):                                        # lane l now plays q
        u = np.array([S[l * PITCH + j] for j in range(16)])
        regs.append(dft(u, inv))
    return regs, worst


def natural(regs):
    X = np.zeros(256, complex)
    for l in range(16):
        for p in range(16):
            X[l + 16 * p] = regs[l][p]
    return X


def pair(kind_a, lane, partner_half):
    """r16_pair(): (source lane of the shuffles, lane pairs inside itself)"""
    l = lane & 15
    src = ((lane & 16) | ((16 - l) & 15)) if kind_a else ((partner_half << 4) | (15 - l))
    return src, (kind_a and l == 0)


def x2_row(M1, pi, h):
    return (M1 // 2 if h else 0) if pi == 0 else (M1 - pi if h else pi)


def x2_partner_half(pi, h):
    return h if pi == 0 else h ^ 1


def warp_partners(M1, pi, regs_by_half):
    """partner value of every (lane, register) of a warp that holds rows x2_row(M1, pi, 0 / 1): what
    r16_partner() returns.  regs_by_half[h][l][p]."""
    out = [[[None] * 16 for _ in range(16)] for _ in range(2)]
    for lane in range(32):
        h, l = lane >> 4, lane & 15
        r = x2_row(M1, pi, h) if M1 > 1 else 0
        src, own = pair(r == 0, lane, x2_partner_half(pi, h))
        for p in range(16):
            out[h][l][p] = regs_by_half[h][l][(16 - p) & 15] if own else regs_by_half[src >> 4][src & 15][15 - p]
    return out
