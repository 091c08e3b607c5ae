// refchan.cu — reference-mode multichannel front end (BASELINE configs[3]).
//
// The reference reaches channel k of a wideband capture only as
//     process(frequency_shift(x, f_k, fs))                    (/root/reference/tetraear/signal/processor.py:85-100, :245-262)
// i.e. M passes over the input, each an NCO (complex128) followed by scipy.signal.decimate(., q) = zero-phase
// cheby1(8) in float64.  The zero-phase response is g[d] = 2 Re sum_i alpha_i p_i^|d| (see dsp_fast.cu), so for every
// channel at once
//     y_k[j] = sum_t g[n_j - t] x[t] e^{-j w_k t}  =  e^{-j w_k n_j} sum_d g[d] e^{+j w_k d} x[n_j - d],   n_j = q j,
// and with w_k = 2 pi k / M the inner sum is a polyphase filterbank: fold the window into M branches
//     v_j[r] = sum_l g[|r + M l|] x[n_j - r - M l],   Y_j[k] = sum_r v_j[r] e^{+j 2 pi k r / M}   (size-M inverse DFT)
// — the input is read once per frame for all M channels.  g is truncated where it has decayed to 1e-10 of its peak
// (|d| < P/2 * M, P = 24 taps per branch for q = 83, M = 800).  What this leaves out is scipy's treatment of the chunk
// ends, which depends on the channel because the reference shifts BEFORE it pads: the odd extension of the shifted
// signal (27 samples), its constant continuation (steady-state initial conditions) and the right-edge transient of
// the backward pass.  refchan_edge_kernel adds exactly those terms, in float64, to the ~P/2*M/q outputs at each end
// of every channel.  Result: decimate(frequency_shift(x, f_k, fs), q) for all k to ~1e-6 of full scale.
#include <cmath>
#include <complex>
#include <cstring>
#include <mutex>
#include <vector>
#include "common.h"
#include "kernels.h"
#include "packed.h"
#include "decim_common.h"

namespace tb200 {

constexpr int RC_PMAX = 32;          // prototype taps per polyphase branch (upper bound)
constexpr int RC_EDGE = 27;          // scipy sosfiltfilt edge (3 * ntaps) of the order-8 decimator
constexpr int RC_NFIX = 160;         // outputs that receive the right-edge transient
constexpr int RC_LT = 192;           // table of lambda^j in the edge kernel (>= RC_NFIX and the ~L/q + 2 edge outputs)

struct RefChanEdge {
    double p[4][2], lam[4][2];       // poles (upper half plane), lambda = p^q
    double g_tail[4][2];             // 2 alpha_i / (1 - p_i): tail sum  Gt(D) = Re sum_i g_tail_i p_i^(D+1)
    double ka[4][2], kM[4][4][2], kMc[4][4][2], pc[4][2];    // right-edge transient (dsp_fast.cu make_fast_params)
    double gL[4][2], gR[4][2];       // g_tail_i p_i^(RC_EDGE + 1)  and  g_tail_i p_i^(n + RC_EDGE - q (n_dec - 1)): tail sums at output 0 / n_dec - 1
    int L;                           // truncation length of g (multiple of M)
};

// twiddles of the in-register transforms: e^{+j 2 pi i / 32} (i < 32), e^{+j 2 pi i / 5} (i < 5), e^{+j 2 pi i / 25} (i < 25)
__constant__ float2 c_w32[32];
__constant__ float2 c_w5[5];
__constant__ float2 c_w25[25];

__device__ __forceinline__ float2 cmulf2(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 caddf2(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csubf2(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// inverse 5-point DFT of (a0..a4), direct
__device__ __forceinline__ void idft5(const float2 (&a)[5], float2 (&o)[5])
{
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        float2 acc = a[0];
#pragma unroll
        for (int n = 1; n < 5; ++n) acc = caddf2(acc, cmulf2(a[n], c_w5[(k * n) % 5]));
        o[k] = acc;
    }
}
// General form, any M = N1 x 2^k: direct small DFTs from the twiddle tables (M = 800 has its own kernel below).
__global__ void __launch_bounds__(256)
refchan_kernel(const float2 *__restrict__ x, long n, int M, int N1, int log2_n2, int F, int P, int l0, int hop, long n_frames,
               const float *__restrict__ proto, const float2 *__restrict__ tw, float2 *__restrict__ out, long out_stride)
{
    extern __shared__ float2 rc_smem[];
    const int N2 = 1 << log2_n2;
    const int APITCH = N2 + 1;
    float2 *v = rc_smem;                      // [F][M]
    float2 *a = v + (long)F * M;              // [F][N1][APITCH]
    float2 *w1 = a + (long)F * N1 * APITCH;   // [N1] exp(+j 2 pi i / N1)
    float2 *w2 = w1 + N1;                     // [N2]
    float2 *wm = w2 + N2;                     // [M]  exp(+j 2 pi i / M)
    const long m0 = (long)blockIdx.x * F;
    for (int i = threadIdx.x; i < N1 + N2 + M; i += blockDim.x) w1[i] = tw[i];
    // ---- folded window: v[f][r] = sum_l h[r + M l] x[n_j - r - M l],  l = l0 .. l0 + P - 1   (proto[(l - l0) M + r]);
    // reference mode: h[d] = g[|d|], l0 = -P/2; the causal etsi prototype: l0 = 0.
    // A thread keeps one r for all F frames: the tap is loaded once per F multiply-adds.
    for (int r = threadIdx.x; r < M; r += blockDim.x) {
        float2 acc[8];
#pragma unroll
        for (int f = 0; f < 8; ++f) acc[f] = make_float2(0.f, 0.f);
        for (int l = l0; l < l0 + P; ++l) {
            const float h = __ldg(proto + (l - l0) * M + r);
            const long base = m0 * hop - r - (long)l * M;
#pragma unroll
            for (int f = 0; f < 8; ++f) {
                const long idx = base + (long)f * hop;
                if (f < F && idx >= 0 && idx < n) {
                    const float2 sv = x[idx];
                    acc[f].x = fmaf(h, sv.x, acc[f].x); acc[f].y = fmaf(h, sv.y, acc[f].y);
                }
            }
        }
#pragma unroll
        for (int f = 0; f < 8; ++f) if (f < F) v[f * M + r] = (m0 + f < n_frames) ? acc[f] : make_float2(0.f, 0.f);
    }
    __syncthreads();
    // ---- stage A: for every r2, a size-N1 DFT over r1 (r = N2 r1 + r2), times exp(+j 2 pi k1 r2 / M)
    for (int jj = threadIdx.x; jj < M; jj += blockDim.x) {
        const int k1 = jj >> log2_n2, r2 = jj & (N2 - 1);
        const float2 t = wm[(k1 * r2) % M];
        for (int f = 0; f < F; ++f) {
            const float2 *vf = v + (long)f * M + r2;
            float2 acc = make_float2(0.f, 0.f);
            int e = 0;                                        // (k1 r1) mod N1
            for (int r1 = 0; r1 < N1; ++r1) {
                float2 s = vf[r1 << log2_n2], w = w1[e];
                acc.x = fmaf(s.x, w.x, fmaf(-s.y, w.y, acc.x)); acc.y = fmaf(s.x, w.y, fmaf(s.y, w.x, acc.y));
                e += k1; if (e >= N1) e -= N1;
            }
            a[((long)f * N1 + k1) * APITCH + r2] = make_float2(acc.x * t.x - acc.y * t.y, acc.x * t.y + acc.y * t.x);
        }
    }
    __syncthreads();
    // ---- stage B: size-N2 DFT over r2; k = k1 + N1 k2; rotation e^{-j 2 pi k n_j / M}; store [k][m0 .. m0+F)
    for (int k = threadIdx.x; k < M; k += blockDim.x) {
        const int k2 = k / N1, k1 = k - k2 * N1;
        float2 *o = out + (long)k * out_stride + m0;
        const int kh = (int)(((long)k * hop) % M);
        for (int f = 0; f < F; ++f) {
            const long j = m0 + f;
            if (j >= n_frames) break;
            const float2 *af = a + ((long)f * N1 + k1) * APITCH;
            float2 acc = make_float2(0.f, 0.f);
            int e = 0;                                        // (k2 r2) mod N2
            for (int r2 = 0; r2 < N2; ++r2) {
                float2 s = af[r2], w = w2[e];
                acc.x = fmaf(s.x, w.x, fmaf(-s.y, w.y, acc.x)); acc.y = fmaf(s.x, w.y, fmaf(s.y, w.x, acc.y));
                e = (e + k2) & (N2 - 1);
            }
            const float2 rot = wm[(int)(((long)kh * (j % M)) % M)];     // e^{+j 2 pi (k hop j) / M}; conjugate applied
            o[f] = make_float2(acc.x * rot.x + acc.y * rot.y, acc.y * rot.x - acc.x * rot.y);
        }
    }
}

// ---- M = 800 = 25 x 32 ---------------------------------------------------------------------------------------------
// A CTA folds the windows of 8 consecutive frames and transforms them.  The window of a frame is P * 800 samples (150 KB
// for P = 24) and the windows of the 8 frames overlap almost entirely, so the input is staged in shared memory one
// prototype block l at a time: the samples that block needs for all 8 frames are the 800 + 7 hop contiguous samples
//     x[hop m0 - 799 - 800 l  ..  hop (m0 + 7) - 800 l],
// brought in by one TMA bulk copy per PAIR of blocks (their ranges overlap: 800 + 800 + 7 hop samples) into a double buffer,
// one pair ahead of the one being accumulated.  The 800-point
// inverse DFT runs in place on the folded windows V[r1][f][r2] (r = 32 r1 + r2, pitch 33) as four small-radix passes —
// 5-point over a (r1 = 5a + b), 5-point over b, 8-point over i and 4-point over g (r2 = 4i + g) — each a thread per
// (row, column) task, so no thread holds more than 8 points and all 416 threads have work.  Output k = k1 + 25 k2.
constexpr int PB_M = 800, PB_F = 8, PB_PITCH = 33, PB_ROWS = 25 * PB_F, PB_THREADS = 416, PB_RS = 400;

// inverse 8-point DFT in registers, radix 2: v[] holds the input in bit-reversed order, the output in natural order
__device__ __forceinline__ void ifft8_regs(float2 (&v)[8])
{
#pragma unroll
    for (int S = 1; S <= 3; ++S) {
        const int m = 1 << S, hm = m >> 1, tstep = 32 >> S;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int b = (i / hm) * m, j = i % hm;
            const float2 t = j ? cmulf2(c_w32[j * tstep], v[b + j + hm]) : v[b + j + hm];
            v[b + j + hm] = csubf2(v[b + j], t);
            v[b + j] = caddf2(v[b + j], t);
        }
    }
}

template <int HOP>                                   // frame spacing as a constant (83: reference mode at 20 MS/s, 400: etsi bank), 0 = run time
__global__ void __launch_bounds__(PB_THREADS, 2)
polybank800_kernel(const float2 *__restrict__ x, long n, int P, int l0, int hop_rt, long n_frames,
                   const float *__restrict__ proto, const float2 *__restrict__ wm_g, float2 *__restrict__ out, long out_stride)
{
    extern __shared__ float2 pb_smem[];
    __shared__ __align__(8) uint64_t full[2];
    float2 *V = pb_smem;                              // [r1 or rho][f][PB_PITCH]
    float2 *wm = V + PB_ROWS * PB_PITCH;              // [800] exp(+j 2 pi i / 800)
    float2 *seg = wm + PB_M;                          // [2][SEGP + 800]: two block pairs
    // block l of the prototype needs x[s0 .. s0 + SEG), s0 = hop m0 - 799 - 800 l: odd, so the staged range starts one
    // sample earlier (16-byte aligned for the bulk copy) and is padded to an even length
    const int hop = HOP ? HOP : hop_rt;
    const int SEG = PB_M + (PB_F - 1) * hop, SEGP = (SEG + 2) & ~1;
    const int tid = threadIdx.x;
    const long m0 = (long)blockIdx.x * PB_F;
    const long s_base = (long)hop * m0 - PB_M;        // first staged sample of block l: s_base - 800 l
    const bool aligned = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) mbar_init(&full[i], 1);
    }
    __syncthreads();
    // Blocks are staged in PAIRS: the ranges of blocks l and l + 1 are 800 samples apart and overlap, so the pair is the
    // contiguous run of SEGP + 800 samples starting at s_base - 800 (l + 1) — one bulk copy and one __syncthreads per two
    // blocks.  Pairs pi whose whole run lies inside the chunk (pb_lo <= pi <= pb_hi) come by TMA, the others by plain
    // loads with zeros outside [0, n).
    const int SEGQ = SEGP + PB_M;                     // samples per pair
    const int n_pairs = P >> 1;                       // P is even (launch check)
    const long s_pair0 = s_base - (long)PB_M * (l0 + 1);        // first sample of pair 0; pair pi starts 1600 pi earlier
    int pb_lo = 1, pb_hi = 0;
    if (aligned) {
        auto fdiv = [](long a, long b) -> long { return a / b - ((a % b != 0) && (a < 0)); };
        pb_hi = (int)max(min(fdiv(s_pair0, 2 * PB_M), 1L << 20), -(1L << 20));
        pb_lo = (int)max(min(-fdiv(n - s_pair0 - SEGQ, 2 * PB_M), 1L << 20), -(1L << 20));
    }
    auto bulk = [&](int pi) -> bool { return pi >= pb_lo && pi <= pb_hi; };
    auto stage = [&](int pi, int buf) {
        const long s0 = s_pair0 - 2L * PB_M * pi;
        float2 *dst = seg + buf * SEGQ;
        if (bulk(pi)) {
            if (tid == 0) {
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                tma_load_1d(dst, x + s0, (unsigned)SEGQ * 8u, &full[buf]);
            }
        } else {
            for (int i = tid; i < SEGQ; i += PB_THREADS) {
                const long t = s0 + i;
                dst[i] = (t >= 0 && t < n) ? x[t] : make_float2(0.f, 0.f);
            }
        }
    };
    stage(0, 0);
    for (int i = tid; i < PB_M; i += PB_THREADS) wm[i] = wm_g[i];
    // ---- folded windows: v_f[r] = sum_l h[r + 800 l] x[hop (m0 + f) - r - 800 l]; sample (r, f) of block l is staged at
    // index 800 - r + hop f of the block's range.  Branches hop apart see the same samples one frame later, so a thread
    // takes the pair (rA, rA + hop) and all 8 frames: 9 shared loads feed 16 packed multiply-adds.  (If
    // hop * ceil(ceil(800 / hop) / 2) exceeds the CTA, a thread takes branches r and r + 400 instead: 16 loads.)
    const int nI = (PB_M + hop - 1) / hop, groups = (nI + 1) / 2;
    const bool toep = hop * groups <= PB_THREADS;
    int rA, rB;
    if (toep) { const int grp = tid / hop; rA = tid - grp * hop + 2 * grp * hop; rB = rA + hop; if (grp >= groups) rA = rB = PB_M; }
    else { rA = tid < PB_RS ? tid : PB_M; rB = tid < PB_RS ? tid + PB_RS : PB_M; }
    const bool vA = rA < PB_M, vB = rB < PB_M;
    float2 acc[2][PB_F];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int f = 0; f < PB_F; ++f) acc[i][f] = make_float2(0.f, 0.f);
    unsigned phase = 0;                               // bit b: parity the next bulk completion of buffer b will have
    // taps are fetched one pair ahead: their latency hides behind the wait.  Threads without a valid second branch
    // read tap 0 and the word in front of the stage: their acc[1] is never stored.
    const float *pA = proto + (vA ? rA : 0), *pB = proto + (vB ? rB : 0);
    float hA_n[2] = {__ldg(pA), __ldg(pA + PB_M)}, hB_n[2] = {__ldg(pB), __ldg(pB + PB_M)};
    const int offA = PB_M - rA, offB = PB_M - rB;
    for (int pi = 0; pi < n_pairs; ++pi) {
        const int buf = pi & 1;
        const float hA[2] = {hA_n[0], hA_n[1]}, hB[2] = {hB_n[0], hB_n[1]};
        if (pi + 1 < n_pairs) {
            pA += 2 * PB_M; pB += 2 * PB_M;
            hA_n[0] = __ldg(pA); hA_n[1] = __ldg(pA + PB_M); hB_n[0] = __ldg(pB); hB_n[1] = __ldg(pB + PB_M);
        }
        if (bulk(pi)) { mbar_wait(&full[buf], (phase >> buf) & 1u); phase ^= 1u << buf; }
        __syncthreads();                              // pair pi is visible; everyone is done with the other buffer
        if (pi + 1 < n_pairs) stage(pi + 1, buf ^ 1);
        if (!vA) continue;
#pragma unroll
        for (int half = 0; half < 2; ++half) {        // block l0 + 2 pi + half lies 800 (1 - half) samples into the pair
            const float2 *sg = seg + buf * SEGQ + (half ? 0 : PB_M);
            if (toep) {
                float2 prev = sg[offB];
#pragma unroll
                for (int f = 0; f < PB_F; ++f) {
                    const float2 cur = sg[offA + hop * f];
                    acc[0][f] = ffma2(splat(hA[half]), cur, acc[0][f]);
                    acc[1][f] = ffma2(splat(hB[half]), prev, acc[1][f]);
                    prev = cur;
                }
            } else {
#pragma unroll
                for (int f = 0; f < PB_F; ++f) {
                    acc[0][f] = ffma2(splat(hA[half]), sg[offA + hop * f], acc[0][f]);
                    acc[1][f] = ffma2(splat(hB[half]), sg[offB + hop * f], acc[1][f]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int r = i ? rB : rA;
        if (r < PB_M) {
            const int r1 = r >> 5, r2 = r & 31;
#pragma unroll
            for (int f = 0; f < PB_F; ++f) V[(r1 * PB_F + f) * PB_PITCH + r2] = (m0 + f < n_frames) ? acc[i][f] : make_float2(0.f, 0.f);
        }
    }
    __syncthreads();
    // ---- pass 1 (in place): 5-point inverse DFT over a, r1 = 5a + b; result c goes to row 5c + b with the twiddle w25^(b c)
    for (int c = tid; c < PB_F * 32 * 5; c += PB_THREADS) {
        const int r2 = c & 31, b = (c >> 5) % 5, f = c / 160;
        float2 *col = V + (b * PB_F + f) * PB_PITCH + r2;
        float2 in[5], o[5];
#pragma unroll
        for (int a = 0; a < 5; ++a) in[a] = col[5 * a * PB_F * PB_PITCH];
        idft5(in, o);
#pragma unroll
        for (int cc = 0; cc < 5; ++cc) col[5 * cc * PB_F * PB_PITCH] = cc ? cmulf2(o[cc], c_w25[b * cc]) : o[cc];
    }
    __syncthreads();
    // ---- pass 2 (in place): 5-point inverse DFT over b of rows 5c + b; k1 = c + 5d goes to row 5c + d, times e^{+j 2 pi k1 r2 / 800}
    for (int c = tid; c < PB_F * 32 * 5; c += PB_THREADS) {
        const int r2 = c & 31, cc = (c >> 5) % 5, f = c / 160;
        float2 *col = V + (5 * cc * PB_F + f) * PB_PITCH + r2;
        float2 in[5], o[5];
#pragma unroll
        for (int b = 0; b < 5; ++b) in[b] = col[b * PB_F * PB_PITCH];
        idft5(in, o);
#pragma unroll
        for (int d = 0; d < 5; ++d) col[d * PB_F * PB_PITCH] = cmulf2(o[d], wm[(cc + 5 * d) * r2]);
    }
    __syncthreads();
    // ---- pass 3 (in place): 8-point inverse DFT over i, r2 = 4i + g; result u goes to column 4u + g with the twiddle w32^(g u)
    for (int c = tid; c < PB_ROWS * 4; c += PB_THREADS) {
        const int row = c % PB_ROWS, g = c / PB_ROWS;
        float2 *rp = V + row * PB_PITCH + g;
        float2 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[((i & 1) << 2) | (i & 2) | ((i & 4) >> 2)] = rp[4 * i];
        ifft8_regs(v);
#pragma unroll
        for (int u = 0; u < 8; ++u) rp[4 * u] = (g && u) ? cmulf2(v[u], c_w32[g * u]) : v[u];
    }
    __syncthreads();
    // ---- pass 4: 4-point inverse DFT over g of columns 4u + g: k2 = u + 8v; channel k = k1 + 25 k2; rotation e^{-j 2 pi k hop j / 800}
    for (int c = tid; c < PB_ROWS * 8; c += PB_THREADS) {
        const int row = c % PB_ROWS, u = c / PB_ROWS;
        const int f = row & (PB_F - 1), rho = row >> 3;
        const long j = m0 + f;
        if (j >= n_frames) continue;
        const float2 *rp = V + row * PB_PITCH + 4 * u;
        const float2 z0 = rp[0], z1 = rp[1], z2 = rp[2], z3 = rp[3];
        const float2 s02 = caddf2(z0, z2), d02 = csubf2(z0, z2), s13 = caddf2(z1, z3), d13 = csubf2(z1, z3);
        float2 X[4];
        X[0] = caddf2(s02, s13);
        X[1] = make_float2(d02.x - d13.y, d02.y + d13.x);          // d02 + j d13
        X[2] = csubf2(s02, s13);
        X[3] = make_float2(d02.x + d13.y, d02.y - d13.x);          // d02 - j d13
        const int k1 = rho / 5 + 5 * (rho % 5);
        const unsigned hj = (unsigned)(((long)hop * j) % PB_M);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const int k = k1 + 25 * (u + 8 * v);
            const float2 rot = wm[((unsigned)k * hj) % PB_M];         // conjugate applied
            out[(long)k * out_stride + j] = make_float2(X[v].x * rot.x + X[v].y * rot.y, X[v].y * rot.x - X[v].x * rot.y);
        }
    }
}

struct cdd2 { double x, y; };
__device__ __forceinline__ cdd2 zmul(cdd2 a, cdd2 b) { return cdd2{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ __forceinline__ cdd2 zpow(cdd2 p, long e)
{
    cdd2 r = cdd2{1.0, 0.0};
    while (e > 0) { if (e & 1) r = zmul(r, p); p = zmul(p, p); e >>= 1; }
    return r;
}
// sample m of the frequency-shifted chunk, numpy's evaluation order: x[m] * exp(W * (m / fs)), W = fl(-2 pi f)
__device__ __forceinline__ cdd2 shifted(const float2 *__restrict__ x, long m, double W, double fs)
{
    double s, c;
    sincos(__dmul_rn(W, (double)m / fs), &s, &c);
    const double xr = (double)x[m].x, xi = (double)x[m].y;
    return cdd2{xr * c - xi * s, xr * s + xi * c};
}

// ---- G_P[k] = sum over the chunk's last K samples of P^(n + 26 - t) x[t] e^{-j 2 pi k t / M}, for the 8 poles P (the four
// upper-half-plane poles of the decimator and their conjugates) and all M channels at once.  It is what the right-edge
// transient of the backward pass needs per channel (F_i[M-1] of the frequency-shifted, padded chunk = the 27 extension
// samples + this sum) and, taken channel by channel as the reference's arithmetic would, it was 70 % of the edge kernel:
// M x K complex multiply-adds with a sincos each.  For all channels together it is again a fold and a DFT:
//     v_P[r] = sum_{t = r mod M} P^(n + 26 - t) x[t],   G_P[k] = sum_r v_P[r] e^{-j 2 pi k r / M}      (8 M-point DFTs).
// Two small kernels, float64 throughout: the fold (thread = (pole, r), up to K / M independent loads) and the DFTs
// (8 lanes per output bin, 100 terms each, shuffle tree).
__global__ void __launch_bounds__(128)
refchan_fold_kernel(const float2 *__restrict__ x, long n, int M, int K, RefChanEdge E, cdd2 *__restrict__ V)
{
    const int pole = blockIdx.x, i = pole & 3;
    const int r = blockIdx.y * blockDim.x + threadIdx.x;
    if (r >= M) return;
    const cdd2 P = cdd2{E.p[i][0], pole < 4 ? E.p[i][1] : -E.p[i][1]};
    const cdd2 PM = zpow(P, M);
    const long t_lo = n + 26 - (long)(K - 1) > 0 ? n + 26 - (long)(K - 1) : 0;
    long t = (n - 1) - ((n - 1 - r) % M);                      // largest t <= n - 1 with t = r (mod M)
    if (r > n - 1) t = -1;                                      // chunk shorter than M: no such sample
    cdd2 pw = zpow(P, n + 26 - t), acc = cdd2{0.0, 0.0};
    while (t >= t_lo) {
        float2 s[4];                                            // four loads in flight per round
#pragma unroll
        for (int u = 0; u < 4; ++u) { const long tu = t - (long)u * M; s[u] = tu >= t_lo ? x[tu] : make_float2(0.f, 0.f); }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const cdd2 term = zmul(pw, cdd2{(double)s[u].x, (double)s[u].y});
            acc.x += term.x; acc.y += term.y;
            pw = zmul(pw, PM);
        }
        t -= 4L * M;
    }
    V[(long)pole * M + r] = acc;
}

__global__ void __launch_bounds__(128)
refchan_gdft_kernel(int M, const cdd2 *__restrict__ V, cdd2 *__restrict__ G)
{
    extern __shared__ cdd2 ts_smem[];
    cdd2 *v = ts_smem;                 // [M]
    cdd2 *w = v + M;                   // [M] exp(-j 2 pi i / M)
    const int pole = blockIdx.x;
    for (int r = threadIdx.x; r < M; r += blockDim.x) {
        v[r] = V[(long)pole * M + r];
        double sn, cs;
        sincospi(-2.0 * (double)r / (double)M, &sn, &cs);
        w[r] = cdd2{cs, sn};
    }
    __syncthreads();
    const int j = threadIdx.x & 7;                             // lane of the bin's group: terms r = j, j + 8, ...
    const int k = blockIdx.y * 16 + (threadIdx.x >> 3);
    cdd2 acc = cdd2{0.0, 0.0};
    if (k < M) {
        int e = (int)(((long)k * j) % M);                      // (k r) mod M
        const int de = (int)(((long)k * 8) % M);
        for (int r = j; r < M; r += 8) {
            const cdd2 a = v[r], b = w[e];
            acc.x = fma(a.x, b.x, fma(-a.y, b.y, acc.x));
            acc.y = fma(a.x, b.y, fma(a.y, b.x, acc.y));
            e += de; if (e >= M) e -= M;
        }
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
    }
    if (k < M && j == 0) G[(long)pole * M + k] = acc;
}

// One CTA per channel: the terms of scipy's padded, zero-phase filtering that lie outside the chunk.
__global__ void __launch_bounds__(128, 4)
refchan_edge_kernel(const float2 *__restrict__ x, long n, int q, double fs, const double *__restrict__ freq, long n_dec,
                    const double *__restrict__ gfull, RefChanEdge E, const cdd2 *__restrict__ Gsum, int n_ch,
                    float2 *__restrict__ out, long out_stride)
{
    __shared__ cdd2 extL[RC_EDGE + 1], extR[RC_EDGE + 1];      // ext[t], t = -1..-27 (index 1..27) / t = n..n+26 (index 0..26)
    __shared__ cdd2 part[4][2];                                 // F_i[M-1] [mode][I/Q]
    __shared__ cdd2 kappa[4][2];
    __shared__ cdd2 lamp[4][RC_LT];
    const int ch = blockIdx.x, tid = threadIdx.x;
    const double f = freq[ch];
    const double W = __dmul_rn(-6.283185307179586, f);
    float2 *o = out + (long)ch * out_stride;
    if (tid <= RC_EDGE && tid >= 1) {
        const cdd2 e0 = shifted(x, 0, W, fs), v = shifted(x, tid, W, fs);
        extL[tid] = cdd2{2.0 * e0.x - v.x, 2.0 * e0.y - v.y};
    }
    if (tid < RC_EDGE) {
        const cdd2 e1 = shifted(x, n - 1, W, fs), v = shifted(x, n - 2 - tid, W, fs);
        extR[tid] = cdd2{2.0 * e1.x - v.x, 2.0 * e1.y - v.y};
    }
    // ---- F_i[M-1] = sum_k p_i^k ext[M-1-k] (M-1 <-> t = n + 26): the 27 extension samples here, the chunk's own samples
    // from refchan_gdft_kernel: with A = G_{p_i}[ch], B = conj(G_{conj p_i}[ch]) the I part is (A + B) / 2, the Q part (A - B) / 2j
    __syncthreads();
    if (tid < 8) {
        const int i = tid >> 1, c = tid & 1;
        const cdd2 p = cdd2{E.p[i][0], E.p[i][1]};
        cdd2 Fv = cdd2{0.0, 0.0};
        for (int k = RC_EDGE - 1; k >= 0; --k) {                 // Horner over ext[n + 26 - k], k = 26 .. 0
            Fv = zmul(p, Fv);
            Fv.x += c ? extR[RC_EDGE - 1 - k].y : extR[RC_EDGE - 1 - k].x;        // ext[n + 26 - k]
        }
        const cdd2 A = Gsum[(long)i * n_ch + ch], Bc = Gsum[(long)(i + 4) * n_ch + ch];
        const cdd2 B = cdd2{Bc.x, -Bc.y};
        if (c == 0) { Fv.x += 0.5 * (A.x + B.x); Fv.y += 0.5 * (A.y + B.y); }
        else        { Fv.x += 0.5 * (A.y - B.y); Fv.y -= 0.5 * (A.x - B.x); }
        part[i][c] = Fv;
    }
    __syncthreads();
    if (tid < 8) {
        const int i = tid >> 1, c = tid & 1;
        const double xl = c ? extR[RC_EDGE - 1].y : extR[RC_EDGE - 1].x;
        cdd2 k = cdd2{E.ka[i][0] * xl, E.ka[i][1] * xl};
        for (int j = 0; j < 4; ++j) {
            const cdd2 Fj = part[j][c];
            const cdd2 t1 = zmul(cdd2{E.kM[i][j][0], E.kM[i][j][1]}, Fj);
            const cdd2 t2 = zmul(cdd2{E.kMc[i][j][0], E.kMc[i][j][1]}, cdd2{Fj.x, -Fj.y});
            k.x += t1.x + t2.x; k.y += t1.y + t2.y;
        }
        kappa[i][c] = k;
    }
    __syncthreads();
    // lambda_i^j, j < RC_LT: every power the two ends need is a host constant times one of these — the tail sum of g
    // beyond distance D, Gt(D) = sum_{d > D} g[d] = Re sum_i g_tail_i p_i^(D+1), has D + 1 = RC_EDGE + 1 + q j at the left end
    // and n + RC_EDGE - q (n_dec - 1) + q jj at the right end; the transient decays as lambda^jj
    for (int e = tid; e < 4 * RC_LT; e += blockDim.x) {
        const int i = e / RC_LT, j = e - i * RC_LT;
        lamp[i][j] = zpow(cdd2{E.lam[i][0], E.lam[i][1]}, j);
    }
    __syncthreads();
    auto g_tail = [&](const double (*g0)[2], long j) -> double {
        double sacc = 0.0;
        for (int i = 0; i < 4; ++i) {
            const cdd2 pw = j < RC_LT ? lamp[i][j] : zpow(cdd2{E.lam[i][0], E.lam[i][1]}, j);
            sacc += (g0[i][0] * pw.x - g0[i][1] * pw.y);
        }
        return sacc;
    };
    const long n_edge = (long)E.L / q + 2;
    // ---- left end: sum_{t<0} g[n_j - t] ext[t]  (27 extension samples + the constant continuation ext[-27])
    for (long j = tid; j < n_edge && j < n_dec; j += blockDim.x) {
        const long nj = j * q;
        double ax = 0.0, ay = 0.0;
        for (int u = 1; u <= RC_EDGE; ++u) {
            const long d = nj + u;
            if (d < E.L) { const double g = gfull[d]; ax += g * extL[u].x; ay += g * extL[u].y; }
        }
        const double gt = g_tail(E.gL, j);
        ax += gt * extL[RC_EDGE].x; ay += gt * extL[RC_EDGE].y;
        float2 v = o[j];
        v.x += (float)ax; v.y += (float)ay;
        o[j] = v;
    }
    __syncthreads();
    // ---- right end: sum_{t>=n} g[t - n_j] ext[t] (+ constant continuation ext[n+26]) - transient of the backward pass
    for (long jj = tid; jj < n_edge && jj < n_dec; jj += blockDim.x) {
        const long j = n_dec - 1 - jj;
        const long nj = j * q;
        double ax = 0.0, ay = 0.0;
        for (int u = 0; u < RC_EDGE; ++u) {
            const long d = n + u - nj;
            if (d < E.L) { const double g = gfull[d]; ax += g * extR[u].x; ay += g * extR[u].y; }
        }
        const double gt = g_tail(E.gR, jj);
        ax += gt * extR[RC_EDGE - 1].x; ay += gt * extR[RC_EDGE - 1].y;
        if (jj < RC_NFIX) {
            // y[k] -= 2 Re sum_i kappa_i pc_i lambda_i^(n_dec-1-k), I and Q separately (real filter, complex state)
            for (int i = 0; i < 4; ++i) {
                const cdd2 gl = zmul(cdd2{E.pc[i][0], E.pc[i][1]}, lamp[i][jj]);
                ax -= 2.0 * (gl.x * kappa[i][0].x - gl.y * kappa[i][0].y);
                ay -= 2.0 * (gl.x * kappa[i][1].x - gl.y * kappa[i][1].y);
            }
        }
        float2 v = o[j];
        v.x += (float)ax; v.y += (float)ay;
        o[j] = v;
    }
}

// ---- host: the zero-phase response of scipy.signal.decimate's float64 design for factor q ----------------------
typedef std::complex<double> cd;
static void putz(double dst[2], cd v) { dst[0] = v.real(); dst[1] = v.imag(); }

int channelizer_factor(int M, int *N1, int *N2);

static int refchan_design(const DecimDesign &d, int q, long n, long n_dec, int M, std::vector<double> &g, RefChanEdge &E)
{
    cd p8[8], r8[8];
    const double gain = d.sos64[0];
    for (int s = 0; s < 4; ++s) {
        double a1 = d.sos64[6 * s + 4], a2 = d.sos64[6 * s + 5];
        double disc = a1 * a1 - 4 * a2;
        if (!(disc < 0)) return fail(TB200_E_ARG, "refchan: real pole pair in section %d", s);
        p8[s] = cd(-a1 / 2, std::sqrt(-disc) / 2);
        p8[s + 4] = std::conj(p8[s]);
    }
    cd prodp(1.0, 0.0);
    for (int i = 0; i < 8; ++i) prodp *= p8[i];
    const double dd = gain / prodp.real();
    for (int i = 0; i < 8; ++i) {
        cd num = gain * std::pow(1.0 + 1.0 / p8[i], 8), den(1.0, 0.0);
        for (int j = 0; j < 8; ++j) if (j != i) den *= (1.0 - p8[j] / p8[i]);
        r8[i] = num / den;
    }
    cd alpha[4];
    double rmax = 0.0;
    for (int j = 0; j < 4; ++j) {
        cd acc(dd, 0.0);
        for (int i = 0; i < 8; ++i) acc += r8[i] / (1.0 - p8[i] * p8[j]);
        alpha[j] = r8[j] * acc;
        rmax = std::max(rmax, std::abs(p8[j]));
    }
    // truncate g where the slowest mode has decayed to 1e-10: a whole number of M-blocks
    long L = (long)std::ceil(std::log(1e-10) / std::log(rmax));
    int half = (int)((L + M - 1) / M);
    if (2 * half > RC_PMAX) return fail(TB200_E_ARG, "refchan: q=%d needs %d taps per branch (max %d)", q, 2 * half, RC_PMAX);
    L = (long)half * M;
    g.assign((size_t)L + 64, 0.0);
    for (int j = 0; j < 4; ++j) {
        cd pw(1.0, 0.0);
        for (long k = 0; k < (long)g.size(); ++k) { g[k] += 2.0 * (alpha[j] * pw).real(); pw *= p8[j]; }
    }
    memset(&E, 0, sizeof(E));
    E.L = (int)L;
    for (int j = 0; j < 4; ++j) {
        putz(E.p[j], p8[j]);
        putz(E.lam[j], std::pow(p8[j], q));
        putz(E.g_tail[j], 2.0 * alpha[j] / (1.0 - p8[j]));
        const cd gt = 2.0 * alpha[j] / (1.0 - p8[j]);
        putz(E.gL[j], gt * std::pow(p8[j], (double)(RC_EDGE + 1)));
        putz(E.gR[j], gt * std::pow(p8[j], (double)(n + RC_EDGE - (long)q * (n_dec - 1))));
    }
    // right-edge transient constants (same algebra as make_fast_params, kept in double)
    cd H1(dd, 0.0);
    for (int j = 0; j < 8; ++j) H1 += r8[j] / (1.0 - p8[j]);
    const long Mx = n + 54, n_last = 27 + (long)q * (n_dec - 1);
    for (int i = 0; i < 4; ++i) {
        cd ka = (H1 - dd) / (1.0 - p8[i]);
        cd kM[8];
        for (int j = 0; j < 8; ++j) {
            ka -= r8[j] * p8[j] / ((1.0 - p8[i] * p8[j]) * (1.0 - p8[j]));
            kM[j] = -r8[j] / (1.0 - p8[i]) + r8[j] * p8[j] / (1.0 - p8[i] * p8[j]);
        }
        putz(E.ka[i], r8[i] * ka);
        for (int j = 0; j < 4; ++j) { putz(E.kM[i][j], r8[i] * kM[j]); putz(E.kMc[i][j], r8[i] * kM[j + 4]); }
        putz(E.pc[i], std::pow(p8[i], (double)(Mx - n_last)));
    }
    return TB200_OK;
}

// workspace layout of h->rc_tab: [proto: RC_PMAX * M floats][twiddles: (M + 512) float2][g: RC_PMAX * M / 2 + 64 doubles][freq: M doubles]
// [G: 8 * M complex128][V: 8 * M complex128] (refchan_fold_kernel / refchan_gdft_kernel)
int launch_refchan(tb200_handle h, const float2 *x, long n, int M, int q, double fs, const DecimDesign &d, float2 *out,
                   long out_stride, cudaStream_t st)
{
    if (!h->rc_tab) return fail(TB200_E_ARG, "refchan: handle was created without max_channels");
    if (M < 2 || M > h->cfg.max_channels) return fail(TB200_E_ARG, "refchan: n_channels %d exceeds the handle's max_channels", M);
    if (n <= 2 * RC_EDGE + 2) return fail(TB200_E_ARG, "refchan: input too short");
    int N1, N2;
    int rc = channelizer_factor(M, &N1, &N2);
    if (rc) return rc;
    const long n_dec = (n + q - 1) / q;
    if (out_stride < n_dec) return fail(TB200_E_ARG, "refchan: out_stride %ld < %ld outputs", out_stride, n_dec);
    std::vector<double> g;
    RefChanEdge E;
    rc = refchan_design(d, q, n, n_dec, M, g, E);
    if (rc) return rc;
    const int P = 2 * (E.L / M);
    float *proto = reinterpret_cast<float *>(h->rc_tab);
    float2 *tw = reinterpret_cast<float2 *>(proto + (size_t)RC_PMAX * h->cfg.max_channels);
    double *gdev = reinterpret_cast<double *>(tw + h->cfg.max_channels + 512);
    double *fdev = gdev + (size_t)RC_PMAX * h->cfg.max_channels / 2 + 64;
    if (!(h->rc_M == M && h->rc_q == q && h->rc_n == n && h->rc_fs == fs)) {
        std::vector<float> hp((size_t)P * M);
        for (int l = -P / 2; l < P / 2; ++l)
            for (int r = 0; r < M; ++r) {
                long dist = std::labs((long)r + (long)M * l);
                hp[(size_t)(l + P / 2) * M + r] = dist < E.L ? (float)g[dist] : 0.f;
            }
        std::vector<float2> twv(N1 + N2 + M);
        for (int i = 0; i < N1; ++i) twv[i] = make_float2((float)std::cos(2 * M_PI * i / N1), (float)std::sin(2 * M_PI * i / N1));
        for (int i = 0; i < N2; ++i) twv[N1 + i] = make_float2((float)std::cos(2 * M_PI * i / N2), (float)std::sin(2 * M_PI * i / N2));
        for (int i = 0; i < M; ++i) twv[N1 + N2 + i] = make_float2((float)std::cos(2 * M_PI * i / M), (float)std::sin(2 * M_PI * i / M));
        std::vector<double> fr(M);
        for (int k = 0; k < M; ++k) fr[k] = (k < M / 2 ? k : k - M) * (fs / M);      // centre frequency of DFT bin k
        TB_CUDA(cudaMemcpyAsync(proto, hp.data(), hp.size() * sizeof(float), cudaMemcpyHostToDevice, st));
        TB_CUDA(cudaMemcpyAsync(tw, twv.data(), twv.size() * sizeof(float2), cudaMemcpyHostToDevice, st));
        TB_CUDA(cudaMemcpyAsync(gdev, g.data(), ((size_t)E.L + 32) * sizeof(double), cudaMemcpyHostToDevice, st));
        TB_CUDA(cudaMemcpyAsync(fdev, fr.data(), fr.size() * sizeof(double), cudaMemcpyHostToDevice, st));
        TB_CUDA(cudaStreamSynchronize(st));            // host vectors go out of scope
        h->rc_M = M; h->rc_q = q; h->rc_n = n; h->rc_fs = fs;
    }
    rc = launch_polyphase_bank(h, x, n, M, N1, N2, P, -P / 2, q, n_dec, proto, tw, out, out_stride, st);
    if (rc) return rc;
    cdd2 *Gsum = reinterpret_cast<cdd2 *>(fdev + h->cfg.max_channels);
    {
        const int K = 128 * ((E.L + 127) / 128);                       // samples of the padded signal the transient looks back on
        cdd2 *Vfold = Gsum + 8 * (size_t)h->cfg.max_channels;
        refchan_fold_kernel<<<dim3(8, (M + 127) / 128), 128, 0, st>>>(x, n, M, K, E, Vfold);
        TB_LAUNCH_CHECK(h);
        const size_t smem = 2 * (size_t)M * sizeof(cdd2);
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(refchan_gdft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return fail(TB200_E_CUDA, "cudaFuncSetAttribute(refchan_gdft): %s", cudaGetErrorString(e));
        }
        refchan_gdft_kernel<<<dim3(8, (M + 15) / 16), 128, smem, st>>>(M, Vfold, Gsum);
        TB_LAUNCH_CHECK(h);
    }
    refchan_edge_kernel<<<M, 128, 0, st>>>(x, n, q, fs, fdev, n_dec, gdev, E, Gsum, M, out, out_stride);
    TB_LAUNCH_CHECK(h);
    return TB200_OK;
}

// The folded-window + size-M inverse DFT + rotation kernel for a prototype of P taps per branch starting at block l0,
// frames `hop` samples apart: y_k[j] = e^{-j 2 pi k hop j / M} sum_r e^{+j 2 pi k r / M} sum_l h[r + M l] x[hop j - r - M l].
// Shared by the reference-mode bank (above) and the etsi channelizer (channelizer.cu: hop = M / 2, l0 = 0).
int launch_polyphase_bank(tb200_handle h, const float2 *x, long n, int M, int N1, int N2, int P, int l0, int hop, long n_frames,
                          const float *proto, const float2 *tw, float2 *out, long out_stride, cudaStream_t st)
{
    if (N1 == 25 && N2 == 32 && (P & 1) == 0) {
        const size_t smem = sizeof(float2) * ((size_t)PB_ROWS * PB_PITCH + PB_M + 2 * ((size_t)((PB_M + (PB_F - 1) * hop + 2) & ~1) + PB_M));
        if (smem <= 200 * 1024) {
            static std::once_flag tw_once[64];
            static cudaError_t tw_err[64];
            const int dv = h->cfg.device & 63;
            std::call_once(tw_once[dv], [&] {
                float2 w32[32], w5[5], w25[25];
                for (int i = 0; i < 32; ++i) w32[i] = make_float2((float)std::cos(2 * M_PI * i / 32), (float)std::sin(2 * M_PI * i / 32));
                for (int i = 0; i < 5; ++i) w5[i] = make_float2((float)std::cos(2 * M_PI * i / 5), (float)std::sin(2 * M_PI * i / 5));
                for (int i = 0; i < 25; ++i) w25[i] = make_float2((float)std::cos(2 * M_PI * i / 25), (float)std::sin(2 * M_PI * i / 25));
                cudaError_t e1 = cudaMemcpyToSymbol(c_w32, w32, sizeof(w32));
                cudaError_t e2 = cudaMemcpyToSymbol(c_w5, w5, sizeof(w5));
                cudaError_t e3 = cudaMemcpyToSymbol(c_w25, w25, sizeof(w25));
                tw_err[dv] = e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3);
            });
            if (tw_err[dv] != cudaSuccess) return fail(TB200_E_CUDA, "polyphase bank: twiddle upload: %s", cudaGetErrorString(tw_err[dv]));
            auto go = [&](auto kern) -> int {
                cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) return fail(TB200_E_CUDA, "cudaFuncSetAttribute(polybank800): %s", cudaGetErrorString(e));
                const unsigned blocks = (unsigned)((n_frames + PB_F - 1) / PB_F);
                kern<<<blocks, PB_THREADS, smem, st>>>(x, n, P, l0, hop, n_frames, proto, tw + N1 + N2, out, out_stride);
                return TB200_OK;
            };
            const int rc = hop == 83 ? go(polybank800_kernel<83>) : hop == 400 ? go(polybank800_kernel<400>) : go(polybank800_kernel<0>);
            if (rc) return rc;
            TB_LAUNCH_CHECK(h);
            return TB200_OK;
        }
    }
    int F = 8;
    int log2_n2 = 0;
    while ((1 << log2_n2) < N2) ++log2_n2;
    auto need = [&](int f) { return sizeof(float2) * ((size_t)f * M + (size_t)f * N1 * (N2 + 1) + N1 + N2 + M); };
    size_t smem = need(F);
    while (smem > 100 * 1024 && F > 1) { F /= 2; smem = need(F); }
    if (smem > 200 * 1024) return fail(TB200_E_ARG, "polyphase bank: M=%d too large for shared memory", M);
    cudaError_t e = cudaFuncSetAttribute(refchan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(TB200_E_CUDA, "cudaFuncSetAttribute(refchan): %s", cudaGetErrorString(e));
    unsigned blocks = (unsigned)((n_frames + F - 1) / F);
    refchan_kernel<<<blocks, 256, smem, st>>>(x, n, M, N1, log2_n2, F, P, l0, hop, n_frames, proto, tw, out, out_stride);
    TB_LAUNCH_CHECK(h);
    return TB200_OK;
}

}  // namespace tb200
