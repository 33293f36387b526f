"""CPU check of the index math behind fft_warp.cuh (see tests/_warp_fft_model.py)."""
import numpy as np
import pytest

from tests import _warp_fft_model as wm


@pytest.mark.parametrize("R", [1, 2, 4, 8, 16, 32])
def test_warp_fft_layouts_and_split(R):
    M = 32 * R
    rng = np.random.default_rng(R)
    w = rng.standard_normal(2 * M)
    z = w[0::2] + 1j * w[1::2]
    Zs = wm.warp_fwd(z, R)
    Zref = np.fft.fft(z)
    k = np.array([[wm.brev5(l) + 32 * d for d in range(R)] for l in range(32)])
    assert np.allclose(Zs, Zref[k])                      # spectral layout: lane l, reg d <-> k = brev5(l) + 32 d
    Zp = wm.partner(Zs, R)
    assert np.allclose(Zp, Zref[(M - k) % M])            # conjugate-partner shuffle
    X = 0.5 * (Zs + np.conj(Zp)) - 0.5j * wm.W(2 * M, k) * (Zs - np.conj(Zp))
    Xref = np.fft.rfft(w)
    assert np.allclose(X, Xref[k])                       # R2C split
    packed = np.zeros(M, complex)
    packed[k] = X
    packed[0] = (Zref[0].real + Zref[0].imag) + 1j * (Zref[0].real - Zref[0].imag)
    assert np.isclose(packed[0].real, Xref[0].real) and np.isclose(packed[0].imag, Xref[M].real)
    Ys, Yp = packed[k], packed[(M - k) % M]
    Zc = (Ys + np.conj(Yp)) + 1j * np.conj(wm.W(2 * M, k)) * (Ys - np.conj(Yp))
    Zc[0, 0] = (packed[0].real + packed[0].imag) + 1j * (packed[0].real - packed[0].imag)
    zt = wm.warp_inv(Zc, R)
    y = np.zeros(2 * M)
    for a in range(32):
        for b in range(R):
            n = R * a + b
            y[2 * n], y[2 * n + 1] = zt[a, b].real, zt[a, b].imag
    assert np.allclose(y / (2 * M), w)                   # C2R: the 1/(2M) lives in the IR spectra


def test_overlap_save_partitioned_model():
    """Uniform partitioned overlap-save with packed spectra (bin 0 = DC/Nyquist as two REAL
    products) reproduces linear convolution -- the engine's math in numpy."""
    rng = np.random.default_rng(0)
    B, P = 32, 5
    h = rng.standard_normal(B * P - 7)
    x = rng.standard_normal(B * 20)
    hp = np.concatenate([h, np.zeros(B * P - len(h))])

    def pack(X):
        p = X[:B].copy()
        p[0] = X[0].real + 1j * X[B].real
        return p

    H = [pack(np.fft.rfft(np.concatenate([hp[k * B:(k + 1) * B], np.zeros(B)]))) / (2 * B) for k in range(P)]
    fdl = [np.zeros(B, complex) for _ in range(P)]
    prev = np.zeros(B)
    y = []
    for t in range(len(x) // B):
        cur = x[t * B:(t + 1) * B]
        fdl = [pack(np.fft.rfft(np.concatenate([prev, cur])))] + fdl[:-1]
        prev = cur
        Y = np.zeros(B, complex)
        for k in range(P):
            prod = fdl[k] * H[k]
            prod[0] = fdl[k][0].real * H[k][0].real + 1j * fdl[k][0].imag * H[k][0].imag
            Y += prod
        full = np.concatenate([[Y[0].real], Y[1:], [Y[0].imag]])
        y.append(np.fft.irfft(full, 2 * B)[B:] * 2 * B)
    y = np.concatenate(y)
    assert np.allclose(y, np.convolve(x, h)[:len(x)])


@pytest.mark.parametrize("s", [0, 1, 3, 5])
def test_cta_fft_layout(s):
    """fft_cta.cuh: DIF stages + 256-point blocks leave bin k at zpos(k); the inverse undoes it."""
    M = 256 << s
    rng = np.random.default_rng(s)
    z = rng.standard_normal(M) + 1j * rng.standard_normal(M)
    sm = wm.cta_fwd(z, s)
    ref = np.fft.fft(z)
    pos = np.array([wm.zpos(k, s) for k in range(M)])
    assert sorted(pos) == list(range(M))
    assert np.allclose(sm[pos], ref)
    back = wm.cta_inv(sm, s)
    assert np.allclose(back / M, z)


# ---- row-FFT family (fft_rows.cuh / kernels_rows.cuh) ---------------------------------------------
def test_row_fft_256_matches_numpy_and_is_bank_conflict_free():
    from tests import _rows_fft_model as R
    rng = np.random.default_rng(0)
    x = rng.standard_normal(256) + 1j * rng.standard_normal(256)
    X, cf = R.row_fft(x)
    assert np.abs(X - np.fft.fft(x)).max() < 1e-12
    assert all(v == 1 for v in cf.worst.values()), cf.worst
    Xi, _ = R.row_fft(x, inv=True)
    assert np.abs(Xi - np.fft.ifft(x) * 256).max() < 1e-12
    assert 8 * 36 <= R.ROW_SLOTS and 4 * 68 <= R.ROW_SLOTS and R.e3(255) < R.ROW_SLOTS


@pytest.mark.parametrize("s", [0, 1, 3, 5])
def test_four_step_position_order_and_split(s):
    from tests import _rows_fft_model as R
    rng = np.random.default_rng(s)
    M = 256 << s
    w = rng.standard_normal(2 * M)
    z = w[0::2] + 1j * w[1::2]
    Zp = R.four_step_forward(z, s)
    pos = np.array([R.zpos(k, s) for k in range(M)])
    assert np.abs(Zp[pos] - np.fft.fft(z)).max() < 1e-10
    Xp = R.split_r2c_pos(Zp, s)
    Xr = np.fft.rfft(w)
    want = Xr[:M].copy()
    want[0] = complex(Xr[0].real, Xr[M].real)
    assert np.abs(Xp[pos] - want).max() < 1e-10
    back = R.four_step_inverse(R.split_c2r_pos(Xp, s), s) / (2 * M)
    assert np.abs(back - z).max() < 1e-12


@pytest.mark.parametrize("M1", [32, 64])
def test_two_stage_column_dft_and_row_pairing(M1):
    """k_tcols_*: n1 = j + 8 b, radix M1/8 over b, twiddle W_M1^(jq), radix 8 over j, k1 = q + (M1/8) r;
    k_trows_*: every row appears once and its split partner sits in the neighbouring warp (or is itself)."""
    from tests import _rows_fft_model as R
    rng = np.random.default_rng(M1)
    Mb = M1 // 8
    x = rng.standard_normal(M1) + 1j * rng.standard_normal(M1)
    S = np.zeros((Mb, 8), complex)
    for j in range(8):
        y = R.dft(np.array([x[j + 8 * b] for b in range(Mb)]), False)
        for q in range(Mb):
            S[q, j] = y[q] * R.W(M1, j * q)
    A = np.zeros(M1, complex)
    for q in range(Mb):
        u = R.dft(S[q], False)
        for r in range(8):
            A[q + Mb * r] = u[r]
    assert np.abs(A - np.fft.fft(x)).max() < 1e-12

    def row_of(rg, warp):
        p, second = 4 * rg + (warp >> 1), warp & 1
        return (M1 // 2 if second else 0) if p == 0 else (M1 - p if second else p)

    rows = []
    for rg in range(M1 // 8):
        for warp in range(8):
            r = row_of(rg, warp)
            rows.append(r)
            assert (M1 - r) % M1 in (r, row_of(rg, warp ^ 1))
    assert sorted(rows) == list(range(M1))


# ---- 16 x 16 row-FFT family (fft_rows16.cuh / kernels_rows16.cuh) ----------------------------------
def test_row_fft_16x16_matches_numpy_and_is_bank_conflict_free():
    from tests import _rows16_fft_model as R
    rng = np.random.default_rng(1)
    x = rng.standard_normal(256) + 1j * rng.standard_normal(256)
    regs, worst = R.fft256x2(x)
    assert np.abs(R.natural(regs) - np.fft.fft(x)).max() < 1e-12
    assert worst == 1
    regs, _ = R.fft256x2(x, inv=True)
    assert np.abs(R.natural(regs) - np.fft.ifft(x) * 256).max() < 1e-12
    assert R.SLOTS >= 256


@pytest.mark.parametrize("M1", [1, 2, 8, 16, 64])
def test_shuffle_split_partners_are_the_conjugate_bins(M1):
    """every (lane, register) of every warp receives bin M - k of its own bin k = k1 + M1 k2 (k2 = l + 16 p)"""
    from tests import _rows16_fft_model as R
    M = 256 * M1
    Z = np.arange(M, dtype=float) + 1j * 0.0      # the value of bin k is k itself
    def row_regs(r):
        return [[Z[r + M1 * (l + 16 * p)] for p in range(16)] for l in range(16)]
    seen = set()
    for pi in range(max(1, M1 // 2)):
        rows = [R.x2_row(M1, pi, h) if M1 > 1 else 0 for h in range(2)]
        parts = R.warp_partners(M1, pi, [row_regs(rows[0]), row_regs(rows[1])])
        for h in range(2):
            for l in range(16):
                for p in range(16):
                    k = rows[h] + M1 * (l + 16 * p)
                    assert parts[h][l][p].real == (M - k) % M, (M1, pi, h, l, p)
            seen.add(rows[h])
    assert seen == set(range(M1))
