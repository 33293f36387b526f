"""numpy model of the one-warp FFT used by the CUDA kernels (cuda-audio_b200/csrc/fft_warp.cuh):
lane/register layouts, DIF-across-lanes butterflies with bit-reversed lane order, the partner
(conjugate-bin) shuffle and the real-FFT split / inverse split formulas.  The CUDA code is a
transcription of these loops; tests/test_fft_model.py checks them against numpy.fft."""
import numpy as np
def brev5(x): return int('{:05b}'.format(x)[::-1],2)
def W(N,k): return np.exp(-2j*np.pi*k/N)
def warp_fwd(z, R):
    M=32*R; reg=np.zeros((32,R),complex)
    for a in range(32):
        for b in range(R): reg[a,b]=z[R*a+b]
    # step1 DIF across lanes
    for s in range(5):
        half=16>>s; new=reg.copy()
        for l in range(32):
            p=reg[l^half]; v=reg[l]; j=l&(half-1)
            if l&half: new[l]=(p-v)*W(2*half,j)
            else: new[l]=v+p
        reg=new
    for l in range(32):
        c=brev5(l)
        for b in range(R): reg[l,b]*=W(M,b*c)
    out=np.zeros_like(reg)
    for l in range(32):
        for d in range(R):
            out[l,d]=sum(reg[l,b]*W(R,b*d) for b in range(R))
    return out  # lane l reg d -> k=brev5(l)+32d
def warp_inv(Zs, R):
    M=32*R; reg=np.zeros((32,R),complex)
    for l in range(32):
        for b in range(R):
            reg[l,b]=sum(Zs[l,d]*np.conj(W(R,b*d)) for d in range(R))
    for l in range(32):
        c=brev5(l)
        for b in range(R): reg[l,b]*=np.conj(W(M,b*c))
    for s in range(5):
        half=1<<s; vt=reg.copy()
        for l in range(32):
            if l&half: vt[l]=reg[l]*np.conj(W(2*half,l&(half-1)))
        new=vt.copy()
        for l in range(32):
            p=vt[l^half]
            new[l]=(p-vt[l]) if (l&half) else (vt[l]+p)
        reg=new
    return reg # lane a reg b -> n=R*a+b
def partner(Zs,R):
    # returns conj-partner array Zp[l,d]=Z[M-k]
    Zp=np.zeros_like(Zs)
    for l in range(32):
        c=brev5(l); lp=brev5((32-c)&31)
        for d in range(R):
            if c==0: Zp[l,d]=Zs[l,(R-d)%R]
            else: Zp[l,d]=Zs[lp,R-1-d]
    return Zp


# ---- CTA-level FFT (fft_cta.cuh): s radix-2 DIF stages in "shared memory" + 256-point blocks ----
def brev(x, bits):
    return int('{:0{w}b}'.format(x, w=bits)[::-1], 2) if bits else 0


def zpos(k, s):
    return (brev(k & ((1 << s) - 1), s) << 8) | (k >> s)


def cta_fwd(z, s):
    M = 256 << s
    sm = np.array(z, complex)
    for st in range(s):
        half = M >> (st + 1)
        new = sm.copy()
        for j in range(M // 2):
            pos = j & (half - 1)
            i0 = ((j // half) * 2 * half) | pos
            i1 = i0 + half
            a, b = sm[i0], sm[i1]
            new[i0] = a + b
            new[i1] = (a - b) * W(M, pos << st)
        sm = new
    for blk in range(1 << s):
        sm[blk * 256:(blk + 1) * 256] = np.fft.fft(sm[blk * 256:(blk + 1) * 256])  # the warp FFT, natural order
    return sm


def cta_inv(sm, s):
    M = 256 << s
    sm = np.array(sm, complex)
    for blk in range(1 << s):
        sm[blk * 256:(blk + 1) * 256] = np.fft.ifft(sm[blk * 256:(blk + 1) * 256]) * 256
    for st in range(s - 1, -1, -1):
        half = M >> (st + 1)
        new = sm.copy()
        for j in range(M // 2):
            pos = j & (half - 1)
            i0 = ((j // half) * 2 * half) | pos
            i1 = i0 + half
            a, b = sm[i0], sm[i1] * np.conj(W(M, pos << st))
            new[i0] = a + b
            new[i1] = a - b
        sm = new
    return sm
