/*
 * oracle/upols_cpu.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * (1) fp64 direct (time-domain) convolution: the mathematical ground truth the
 *     north-star names ("FP64 direct-convolution oracle", tolerance 1e-4).
 * (2) A self-contained fp32 CPU uniform-partitioned overlap-save convolver
 *     with its own real FFT (FFTW is not installed in this image): the
 *     "CPU partitioned convolution" baseline of BASELINE.md section 4(b),
 *     i.e. bench.py's cpu_baseline of kind "port".  It restates the NEW
 *     engine's math ("New-engine math that replaces
 *     a6-a14") with the reference's parameter semantics:
 *       wet glide           conv.cu:15-32,339-353   (g += (wet-g)/(vsteps+5))
 *       pan law / level     conv.cu:386-401
 *       predelay (input 0)  conv.cu:411-415
 *       clamp(+-1) wet      conv.cu:98
 *       dry mix after clamp conv.cu:418-427
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this library; the product never links or calls it.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------ */
/* (1) fp64 direct convolution: y[n] = sum_m h[m] x[n-m], n in [0, ny)        */
/* ------------------------------------------------------------------------ */
void oracle_direct_conv_f64(const double *x, size_t nx, const double *h, size_t nh, double *y,
                            size_t ny)
{
#pragma omp parallel for schedule(dynamic, 256)
    for (long n = 0; n < (long)ny; n++) {
        size_t m0 = (size_t)n >= nx ? (size_t)n - nx + 1 : 0;
        size_t m1 = (size_t)n < nh - 1 ? (size_t)n : nh - 1;
        double acc = 0.0;
        for (size_t m = m0; m <= m1 && m < nh; m++) acc += h[m] * x[n - m];
        y[n] = acc;
    }
}

/* sampled variant for very long IRs: y[idx[j]] only */
void oracle_direct_conv_f64_at(const double *x, size_t nx, const double *h, size_t nh,
                               const int64_t *idx, size_t nidx, double *y)
{
#pragma omp parallel for schedule(dynamic, 16)
    for (long j = 0; j < (long)nidx; j++) {
        size_t n = (size_t)idx[j];
        size_t m0 = n >= nx ? n - nx + 1 : 0;
        size_t m1 = n < nh - 1 ? n : nh - 1;
        double acc = 0.0;
        for (size_t m = m0; m <= m1 && m < nh; m++) acc += h[m] * x[n - m];
        y[j] = acc;
    }
}

/* ------------------------------------------------------------------------ */
/* (2) fp32 CPU uniform-partitioned overlap-save                              */
/* ------------------------------------------------------------------------ */
typedef struct {
    int M;            /* complex FFT size = B */
    float *wr, *wi;   /* W_M^k, k < M/2 */
    float *sr, *si;   /* W_2M^k, k < M   (real-FFT split twiddles) */
    int *rev;
} rfft_plan;

static void plan_init(rfft_plan *p, int M)
{
    p->M = M;
    p->wr = (float *)malloc(sizeof(float) * (M / 2 + 1));
    p->wi = (float *)malloc(sizeof(float) * (M / 2 + 1));
    p->sr = (float *)malloc(sizeof(float) * M);
    p->si = (float *)malloc(sizeof(float) * M);
    p->rev = (int *)malloc(sizeof(int) * M);
    for (int k = 0; k < M / 2; k++) {
        p->wr[k] = (float)cos(-2.0 * M_PI * k / M);
        p->wi[k] = (float)sin(-2.0 * M_PI * k / M);
    }
    for (int k = 0; k < M; k++) {
        p->sr[k] = (float)cos(-M_PI * k / M);
        p->si[k] = (float)sin(-M_PI * k / M);
    }
    int bits = 0;
    while ((1 << bits) < M) bits++;
    for (int i = 0; i < M; i++) {
        int r = 0;
        for (int b = 0; b < bits; b++) if (i & (1 << b)) r |= 1 << (bits - 1 - b);
        p->rev[i] = r;
    }
}
static void plan_free(rfft_plan *p) { free(p->wr); free(p->wi); free(p->sr); free(p->si); free(p->rev); }

/* in-place complex radix-2 DIT on split arrays, sign = -1 fwd, +1 inverse */
static void cfft(const rfft_plan *p, float *re, float *im, int inverse)
{
    int M = p->M;
    for (int i = 0; i < M; i++) {
        int j = p->rev[i];
        if (i < j) { float t = re[i]; re[i] = re[j]; re[j] = t; t = im[i]; im[i] = im[j]; im[j] = t; }
    }
    for (int len = 2; len <= M; len <<= 1) {
        int half = len >> 1, step = M / len;
        for (int i = 0; i < M; i += len)
            for (int k = 0; k < half; k++) {
                float wr = p->wr[k * step], wi = inverse ? -p->wi[k * step] : p->wi[k * step];
                float ur = re[i + k], ui = im[i + k];
                float xr = re[i + k + half], xi = im[i + k + half];
                float vr = xr * wr - xi * wi, vi = xr * wi + xi * wr;
                re[i + k] = ur + vr; im[i + k] = ui + vi;
                re[i + k + half] = ur - vr; im[i + k + half] = ui - vi;
            }
    }
}

/* real FFT of 2M reals -> packed M complex (bin 0 = (DC, Nyquist)), split layout */
static void rfft_fwd(const rfft_plan *p, const float *w, float *Xr, float *Xi, float *tr, float *ti)
{
    int M = p->M;
    for (int n = 0; n < M; n++) { tr[n] = w[2 * n]; ti[n] = w[2 * n + 1]; }
    cfft(p, tr, ti, 0);
    Xr[0] = tr[0] + ti[0];
    Xi[0] = tr[0] - ti[0];
    for (int k = 1; k < M; k++) {
        float zr = tr[k], zi = ti[k], pr = tr[M - k], pi = -ti[M - k]; /* conj Z[M-k] */
        float er = 0.5f * (zr + pr), ei = 0.5f * (zi + pi);
        float dr = zr - pr, di = zi - pi;                 /* Z - conj Zp */
        /* -(j/2) * W * d */
        float wr = p->sr[k], wi = p->si[k];
        float mr = dr * wr - di * wi, mi = dr * wi + di * wr;
        Xr[k] = er + 0.5f * mi;
        Xi[k] = ei - 0.5f * mr;
    }
}

/* inverse: packed spectrum (pre-scaled by 1/(2M)) -> 2M reals */
static void rfft_inv(const rfft_plan *p, const float *Yr, const float *Yi, float *y, float *tr, float *ti)
{
    int M = p->M;
    tr[0] = Yr[0] + Yi[0];
    ti[0] = Yr[0] - Yi[0];
    for (int k = 1; k < M; k++) {
        float ar = Yr[k], ai = Yi[k], pr = Yr[M - k], pi = -Yi[M - k];
        float sr = ar + pr, si = ai + pi;
        float dr = ar - pr, di = ai - pi;
        float wr = p->sr[k], wi = -p->si[k]; /* conj W_2M^k */
        float mr = dr * wr - di * wi, mi = dr * wi + di * wr;
        /* Z = s + j*m */
        tr[k] = sr - mi;
        ti[k] = si + mr;
    }
    cfft(p, tr, ti, 1);
    for (int n = 0; n < M; n++) { y[2 * n] = tr[n]; y[2 * n + 1] = ti[n]; }
}

#define UP_RING 16384 /* predelay ring per input: >= 8192 + 2*B */

typedef struct {
    float wet, dry, level, panWet, panDry;
    uint32_t predelay, select, vsteps;
    float g; /* glide state */
} upols_in;

typedef struct {
    int B, P, n_in, n_out, n_inst, n_ir;
    rfft_plan plan;
    float *Hr, *Hi;   /* [n_ir][n_out][P][B] */
    float *Xr, *Xi;   /* [n_inst][n_in][P][B] ring; newest at slot head, older at head+1.. */
    float *ring;      /* [n_inst][n_in][UP_RING] delayed + scaled input stream */
    upols_in *par;    /* [n_inst][n_in] */
    int head;
    uint64_t t;       /* period counter */
} upols;

upols *upols_create(int B, int max_ir_frames, int n_in, int n_out, int n_inst, int n_ir)
{
    upols *u = (upols *)calloc(1, sizeof(upols));
    u->B = B; u->P = (max_ir_frames + B - 1) / B; u->n_in = n_in; u->n_out = n_out;
    u->n_inst = n_inst; u->n_ir = n_ir;
    plan_init(&u->plan, B);
    size_t hs = (size_t)n_ir * n_out * u->P * B, xs = (size_t)n_inst * n_in * u->P * B;
    u->Hr = (float *)calloc(hs, sizeof(float)); u->Hi = (float *)calloc(hs, sizeof(float));
    u->Xr = (float *)calloc(xs, sizeof(float)); u->Xi = (float *)calloc(xs, sizeof(float));
    u->ring = (float *)calloc((size_t)n_inst * n_in * UP_RING, sizeof(float));
    u->par = (upols_in *)calloc((size_t)n_inst * n_in, sizeof(upols_in));
    for (int i = 0; i < n_inst * n_in; i++) {
        upols_in *p = &u->par[i];
        p->wet = 0.5f; p->dry = 0.5f; p->level = 1.0f; p->select = 0; p->g = 0.0f;
    }
    return u;
}

void upols_destroy(upols *u)
{
    if (!u) return;
    plan_free(&u->plan);
    free(u->Hr); free(u->Hi); free(u->Xr); free(u->Xi); free(u->ring); free(u->par); free(u);
}

int upols_P(const upols *u) { return u->P; }

/* stereo (or mono) IR -> bank slot; chan[o] = time-domain IR for output o */
int upols_load_ir(upols *u, int slot, const float *left, const float *right, int frames)
{
    if (slot < 0 || slot >= u->n_ir) return -1;
    int B = u->B, P = u->P;
    if (frames > P * B) frames = P * B;
    float *w = (float *)calloc(2 * B, sizeof(float));
    float *tr = (float *)malloc(sizeof(float) * B), *ti = (float *)malloc(sizeof(float) * B);
    const float scale = 1.0f / (2.0f * B);
    for (int o = 0; o < u->n_out; o++) {
        const float *h = o == 0 ? left : right;
        for (int k = 0; k < P; k++) {
            memset(w, 0, sizeof(float) * 2 * B);
            for (int n = 0; n < B; n++) {
                int s = k * B + n;
                w[n] = s < frames ? h[s] * scale : 0.0f;
            }
            size_t off = (((size_t)slot * u->n_out + o) * P + k) * B;
            rfft_fwd(&u->plan, w, u->Hr + off, u->Hi + off, tr, ti);
        }
    }
    free(w); free(tr); free(ti);
    return 0;
}

void upols_set_param(upols *u, int inst, int input, float wet, float dry, float level, float panWet,
                     float panDry, uint32_t predelay, uint32_t select, int32_t vsteps)
{
    upols_in *p = &u->par[inst * u->n_in + input];
    p->wet = wet; p->dry = dry; p->level = level; p->panWet = panWet; p->panDry = panDry;
    p->predelay = predelay; p->select = select;
    if (vsteps >= 0) p->vsteps = (uint32_t)vsteps;
}

/* force the glide state (tests use it to skip the 80-period fade-in) */
void upols_set_glide(upols *u, int inst, int input, float g) { u->par[inst * u->n_in + input].g = g; }

static inline float clampf(float v) { return v < -1.f ? -1.f : (v > 1.f ? 1.f : v); }

/* one period for every instance; in: [n_inst][n_in][B], out: [n_inst][n_out][B] */
int upols_process(upols *u, const float *in, float *out, int nframes)
{
    if (nframes != u->B) return -1;
    const int B = u->B, P = u->P, n_in = u->n_in, n_out = u->n_out;
    const int head = u->head = (u->head + P - 1) % P; /* ring runs backwards: slot head+k = X[t-k] */
    const uint64_t t = u->t;
    const uint32_t pd0_mask = UP_RING - 1;
#pragma omp parallel
    {
        float *w = (float *)malloc(sizeof(float) * 2 * B);
        float *tr = (float *)malloc(sizeof(float) * B), *ti = (float *)malloc(sizeof(float) * B);
        float *Yr = (float *)malloc(sizeof(float) * B), *Yi = (float *)malloc(sizeof(float) * B);
        float *y = (float *)malloc(sizeof(float) * 2 * B);
#pragma omp for schedule(static)
        for (int s = 0; s < u->n_inst; s++) {
            /* predelay of input 0 applies to every input (conv.cu:412,415) */
            uint32_t pd = u->par[s * n_in].predelay;
            for (int i = 0; i < n_in; i++) {
                upols_in *p = &u->par[s * n_in + i];
                p->g = p->g + (p->wet - p->g) / (float)(p->vsteps + 5);
                if (p->vsteps > 0) p->vsteps--;
                float g = p->g * p->level;
                float *ring = u->ring + ((size_t)s * n_in + i) * UP_RING;
                const float *x = in + ((size_t)s * n_in + i) * B;
                uint32_t base = (uint32_t)((t * B) & pd0_mask);
                for (int n = 0; n < B; n++) ring[(base + pd + n) & pd0_mask] += g * x[n];
                uint32_t prev = (uint32_t)(((t + (UP_RING / B) - 1) * B) & pd0_mask);
                for (int n = 0; n < B; n++) { w[n] = ring[(prev + n) & pd0_mask]; w[B + n] = ring[(base + n) & pd0_mask]; }
                for (int n = 0; n < B; n++) ring[(prev + n) & pd0_mask] = 0.0f; /* consumed */
                size_t off = (((size_t)s * n_in + i) * P + head) * B;
                rfft_fwd(&u->plan, w, u->Xr + off, u->Xi + off, tr, ti);
            }
            for (int o = 0; o < n_out; o++) {
                memset(Yr, 0, sizeof(float) * B); memset(Yi, 0, sizeof(float) * B);
                for (int i = 0; i < n_in; i++) {
                    const upols_in *p = &u->par[s * n_in + i];
                    float pan = o == 0 ? (p->panWet >= 0 ? 1 - p->panWet : 1) : (p->panWet <= 0 ? 1 + p->panWet : 1);
                    if (n_out == 1) pan = 1.0f;
                    const float *Hr = u->Hr + ((size_t)p->select * n_out + o) * P * B;
                    const float *Hi = u->Hi + ((size_t)p->select * n_out + o) * P * B;
                    const float *Xr = u->Xr + ((size_t)s * n_in + i) * P * B;
                    const float *Xi = u->Xi + ((size_t)s * n_in + i) * P * B;
                    float a0 = 0.f, a1 = 0.f;
                    float *restrict accR = w, *restrict accI = w + B;
                    memset(w, 0, sizeof(float) * 2 * B);
                    for (int k = 0; k < P; k++) {
                        int slot = head + k; if (slot >= P) slot -= P;
                        const float *restrict xr = Xr + (size_t)slot * B, *restrict xi = Xi + (size_t)slot * B;
                        const float *restrict hr = Hr + (size_t)k * B, *restrict hi = Hi + (size_t)k * B;
                        a0 += xr[0] * hr[0]; a1 += xi[0] * hi[0];
#pragma omp simd
                        for (int b = 0; b < B; b++) {
                            accR[b] += xr[b] * hr[b] - xi[b] * hi[b];
                            accI[b] += xr[b] * hi[b] + xi[b] * hr[b];
                        }
                    }
                    w[0] = a0; w[B] = a1; /* packed DC / Nyquist are two real products */
                    for (int b = 0; b < B; b++) { Yr[b] += pan * w[b]; Yi[b] += pan * w[B + b]; }
                }
                rfft_inv(&u->plan, Yr, Yi, y, tr, ti);
                float *dst = out + ((size_t)s * n_out + o) * B;
                for (int n = 0; n < B; n++) {
                    float v = clampf(y[B + n]);
                    for (int i = 0; i < n_in; i++) {
                        const upols_in *p = &u->par[s * n_in + i];
                        float pan = o == 0 ? (p->panDry >= 0 ? 1 - p->panDry : 1) : (p->panDry <= 0 ? 1 + p->panDry : 1);
                        if (n_out == 1) pan = 1.0f;
                        v += in[((size_t)s * n_in + i) * B + n] * (p->dry * pan * p->level);
                    }
                    dst[n] = v;
                }
            }
        }
        free(w); free(tr); free(ti); free(Yr); free(Yi); free(y);
    }
    u->t++;
    return 0;
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
