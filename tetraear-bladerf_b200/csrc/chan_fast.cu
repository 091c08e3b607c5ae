// chan_fast.cu — time-parallel NCO + zero-phase Butterworth channel filter ("FAST" mode) for
// SignalProcessor.frequency_shift / filter_signal as chained by process()
// (/root/reference/tetraear/signal/processor.py:85-100, :51-83, :260-264).
//
// Same arithmetic model as the reference — float64, butter(4) run forward and backward over the
// odd-extended chunk with steady-state initial conditions (scipy filtfilt) — evaluated in the
// two-sided one-pole form (see dsp_fast.cu): with F_i[n] = sum_k p_i^k x[n-k] and
// B_i[n] = sum_k p_i^k x[n+k] over the padded signal,
//     y[n] = (c0 - 2 sum Re a_i) x[n] + 2 Re sum_i a_i (F_i[n] + B_i[n]) - right-edge transient.
// Each lane owns R = 8 consecutive samples, runs both one-pole recurrences from a zero state,
// warp Kogge-Stone scans give the in-pass carries and the neighbour warp's end state the
// cross-pass carry (|p|max^256 < 1e-15).  All float64: agrees with scipy to ~1e-13.
#include <algorithm>
#include <complex>
#include <mutex>
#include <vector>
#include "common.h"
#include "kernels.h"
#include "packed.h"

namespace tb200 {

constexpr int CF_R = 8;
constexpr int CF_P = 32 * CF_R;
constexpr int CF_W = 8;            // warps per CTA: warp 0 = both halo half-passes, warps 1..7 = storing passes
constexpr int CF_FIX = 320;        // samples that receive the right-edge term (0.88^320 ~ 2e-18)

struct ChanFastParams {
    double p[2][2];
    double ppow[6][2][2];          // (p^R)^(2^s)  (s = 5 only for the edge term)
    double c0p;                    // c0 - 2 sum Re(alpha)
    double a[2][2];                // alpha_i
    double aF[CF_R][2][2];         // alpha_i p^(r+1)
    double aB[CF_R][2][2];         // alpha_i p^(R-r)
    double inv1mp[2][2];
    double ka[2][2], kM[2][2][2], kMc[2][2][2];
    double pc[2][2];               // p^16 : exponent of the last output (M - n_last)
    double a2[2][2];               // 2 alpha_i
    double aF2[CF_R][2][2];        // 2 alpha_i p^(r+1)
    double aB2[CF_R][2][2];        // 2 alpha_i p^(R-r)
    double p7[2][2], p15[2][2];
    double lanepow[32][2][2];      // (p^R)^l: weight of the cross-pass carry at lane l
    double G[CF_R];                // zero-state response inside a lane: y[r] = sum_j G[|r - j|] x[j]  (G[0] includes c0p)
    double PW[2][CF_R][2];         // p_i^k, k < R: lane aggregates  sum_j p^(R-1-j) x[j]  /  sum_j p^j x[j]
    int nstage[2];                 // Kogge-Stone stages whose multiplier (p^R)^(2^s) is above 1e-17
};

struct cdd { double x, y; };
__device__ __forceinline__ cdd cmuld(cdd a, cdd b) { return cdd{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ __forceinline__ cdd cfmad(cdd a, cdd b, cdd c) { return cdd{fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y))}; }
__device__ __forceinline__ cdd dshfl_up(cdd v, int d) { return cdd{__shfl_up_sync(0xffffffffu, v.x, d), __shfl_up_sync(0xffffffffu, v.y, d)}; }
__device__ __forceinline__ cdd dshfl_down(cdd v, int d) { return cdd{__shfl_down_sync(0xffffffffu, v.x, d), __shfl_down_sync(0xffffffffu, v.y, d)}; }
__device__ __forceinline__ cdd dshfl(cdd v, int l) { return cdd{__shfl_sync(0xffffffffu, v.x, l), __shfl_sync(0xffffffffu, v.y, l)}; }
__device__ __forceinline__ cdd ldd(const double c[2]) { return cdd{c[0], c[1]}; }

// working precision of the main kernel: double (default) or float (TB200 "fast32" channel filter)
template <typename T> struct cx { T x, y; };
template <typename T> __device__ __forceinline__ cx<T> cxmul(cx<T> a, cx<T> b) { return cx<T>{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
template <typename T> __device__ __forceinline__ cx<T> cxfma(cx<T> a, cx<T> b, cx<T> c) { return cx<T>{a.x * b.x - a.y * b.y + c.x, a.x * b.y + a.y * b.x + c.y}; }
template <typename T> __device__ __forceinline__ cx<T> cxup(cx<T> v, int d) { return cx<T>{__shfl_up_sync(0xffffffffu, v.x, d), __shfl_up_sync(0xffffffffu, v.y, d)}; }
template <typename T> __device__ __forceinline__ cx<T> cxdown(cx<T> v, int d) { return cx<T>{__shfl_down_sync(0xffffffffu, v.x, d), __shfl_down_sync(0xffffffffu, v.y, d)}; }
template <typename T> __device__ __forceinline__ cx<T> cxld(const double c[2]) { return cx<T>{(T)c[0], (T)c[1]}; }

// sample i of the chunk after the optional NCO, complex128
template <typename TI>
__device__ __forceinline__ cdd nco_sample(const TI *__restrict__ x, long i, double W, double fs)
{
    double xr = (double)x[2 * i], xi = (double)x[2 * i + 1];
    if (W == 0.0) return cdd{xr, xi};
    double t = (double)i / fs, s, c;
    sincos(__dmul_rn(W, t), &s, &c);
    return cdd{__dsub_rn(__dmul_rn(xr, c), __dmul_rn(xi, s)), __dadd_rn(__dmul_rn(xr, s), __dmul_rn(xi, c))};
}

// padded signal at ext position t (edge 15), constants beyond both ends
template <typename TI>
__device__ __forceinline__ cdd chan_padded(const TI *__restrict__ x, long n, long t, double W, double fs)
{
    const long edge = 15;
    if (t >= edge && t < edge + n) return nco_sample<TI>(x, t - edge, W, fs);
    long ie, ii;
    if (t < edge) { ie = 0; ii = t < 0 ? edge : edge - t; }
    else { long u = t - edge - n; ie = n - 1; ii = u < edge ? n - 2 - u : n - 1 - edge; }
    if (W == 0.0 && sizeof(TI) == 4) {      // complex64 reaches filtfilt un-promoted: odd_ext rounds in float32
        float er = (float)x[2 * ie], ei = (float)x[2 * ie + 1], vr = (float)x[2 * ii], vi = (float)x[2 * ii + 1];
        return cdd{(double)__fsub_rn(__fmul_rn(2.0f, er), vr), (double)__fsub_rn(__fmul_rn(2.0f, ei), vi)};
    }
    cdd e = nco_sample<TI>(x, ie, W, fs), v = nco_sample<TI>(x, ii, W, fs);
    return cdd{__dsub_rn(__dmul_rn(2.0, e.x), v.x), __dsub_rn(__dmul_rn(2.0, e.y), v.y)};
}

constexpr int CG_NS = CF_W - 1;       // storing passes per CTA (warps 1..7); warp 0 runs both halo half-passes
constexpr int CG_NP = CF_W + 1;       // passes per tile: left halo, CG_NS storing, right halo

// one direction over the lane's CF_R samples (both modes, I and Q): zero-state recurrences, Kogge-Stone scan of the
// lane aggregates, in-warp carry applied to y; the pass-end state goes to carry[I/Q][mode]
template <typename T, bool FWD>
__device__ __forceinline__ void cg_dir(const cx<T> (&xs)[CF_R], T (&yI)[CF_R], T (&yQ)[CF_R], int lane, const ChanFastParams &K,
                                       cx<T> (*carry)[2], cx<T> *FeI, cx<T> *FeQ, bool edge_lane)
{
    typedef cx<T> C;
    const C p0 = cxld<T>(K.p[0]), p1 = cxld<T>(K.p[1]), a0 = cxld<T>(K.a[0]), a1 = cxld<T>(K.a[1]);
    C sI0 = C{T(0), T(0)}, sQ0 = sI0, sI1 = sI0, sQ1 = sI0;
#pragma unroll
    for (int i = 0; i < CF_R; ++i) {
        const int r = FWD ? i : CF_R - 1 - i;
        sI0 = cxmul<T>(p0, sI0); sI0.x += xs[r].x;
        sQ0 = cxmul<T>(p0, sQ0); sQ0.x += xs[r].y;
        sI1 = cxmul<T>(p1, sI1); sI1.x += xs[r].x;
        sQ1 = cxmul<T>(p1, sQ1); sQ1.x += xs[r].y;
        yI[r] += T(2) * (a0.x * sI0.x - a0.y * sI0.y + a1.x * sI1.x - a1.y * sI1.y);
        yQ[r] += T(2) * (a0.x * sQ0.x - a0.y * sQ0.y + a1.x * sQ1.x - a1.y * sQ1.y);
        if (FWD && r == 6) { FeI[0] = sI0; FeI[1] = sI1; FeQ[0] = sQ0; FeQ[1] = sQ1; }      // zero-state F at r = 6
    }
    C sI[2] = {sI0, sI1}, sQ[2] = {sQ0, sQ1};
#pragma unroll
    for (int m = 0; m < 2; ++m) {
#pragma unroll
        for (int st = 0; st < 5; ++st) {
            C oI = FWD ? cxup<T>(sI[m], 1 << st) : cxdown<T>(sI[m], 1 << st);
            C oQ = FWD ? cxup<T>(sQ[m], 1 << st) : cxdown<T>(sQ[m], 1 << st);
            C g = cxld<T>(K.ppow[st][m]);
            if (FWD ? (lane < (1 << st)) : (lane + (1 << st) >= 32)) g = C{T(0), T(0)};
            sI[m] = cxfma<T>(g, oI, sI[m]);
            sQ[m] = cxfma<T>(g, oQ, sQ[m]);
        }
        C cI = FWD ? cxup<T>(sI[m], 1) : cxdown<T>(sI[m], 1);
        C cQ = FWD ? cxup<T>(sQ[m], 1) : cxdown<T>(sQ[m], 1);
        if (lane == (FWD ? 0 : 31)) cI = cQ = C{T(0), T(0)};
        if (lane == (FWD ? 31 : 0)) { carry[0][m] = sI[m]; carry[1][m] = sQ[m]; }
        if (FWD && edge_lane) {                                   // + p^7 * (carry from lane 0 of this warp)
            C p7 = cxld<T>(K.p[m]);
            const C pm = p7;
            for (int i = 0; i < 6; ++i) p7 = cxmul<T>(p7, pm);
            FeI[m] = cxfma<T>(p7, cI, FeI[m]);
            FeQ[m] = cxfma<T>(p7, cQ, FeQ[m]);
        }
#pragma unroll
        for (int r = 0; r < CF_R; ++r) {
            const C af = FWD ? cxld<T>(K.aF[r][m]) : cxld<T>(K.aB[r][m]);
            yI[r] += T(2) * (af.x * cI.x - af.y * cI.y);
            yQ[r] += T(2) * (af.x * cQ.x - af.y * cQ.y);
        }
    }
}

template <typename TI, typename T>
__device__ __forceinline__ void cg_load(const TI *__restrict__ xc, long n, long m0, double W, double fs, cx<T> (&xs)[CF_R])
{
    typedef cx<T> C;
    if (sizeof(TI) == 4 && W == 0.0 && m0 >= 0 && m0 + CF_R <= n && ((reinterpret_cast<uintptr_t>(xc + 2 * m0) & 15) == 0)) {
        // interior lane, complex64, no NCO: 8 samples = 64 contiguous bytes = four 16-byte loads (tiles are end-aligned)
        const float4 *v4 = reinterpret_cast<const float4 *>(xc + 2 * m0);
#pragma unroll
        for (int r = 0; r < CF_R / 2; ++r) {
            float4 v = v4[r];
            xs[2 * r] = C{(T)v.x, (T)v.y};
            xs[2 * r + 1] = C{(T)v.z, (T)v.w};
        }
    } else {
#pragma unroll
        for (int r = 0; r < CF_R; ++r) { cdd v = chan_padded<TI>(xc, n, 15 + m0 + r, W, fs); xs[r] = C{(T)v.x, (T)v.y}; }
    }
}

template <typename TI, typename T>
__global__ void __launch_bounds__(CF_W * 32, 2)
chan_fast_kernel(const TI *__restrict__ x, long n, long in_stride, int tiles_per_chunk,
                 const double *__restrict__ freq_off, double fs, ChanFastParams K,
                 double *__restrict__ filtered, long out_stride, double *__restrict__ phase_partial, int sps, int step)
{
    typedef cx<T> C;
    __shared__ C carryF[CG_NP][2][2];         // [pass][I/Q][mode]
    __shared__ C carryB[CG_NP][2][2];
    __shared__ C kappa[2][2];                 // [I/Q][mode] right-edge transient amplitudes of this chunk (last tile only)
    __shared__ T pws[CG_NS * CF_P];           // |y|^2 of the tile's storing range, for the timing-phase sums
    const int chunk = blockIdx.x / tiles_per_chunk;
    const int tile = blockIdx.x - chunk * tiles_per_chunk;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const TI *xc = x + (long)chunk * in_stride * 2;
    const double f = freq_off ? freq_off[chunk] : 0.0;
    const double W = (f != 0.0) ? __dmul_rn(-6.283185307179586, f) : 0.0;
    // tiles are aligned to the END of the chunk (the first may be partial), so the last CF_FIX samples and the right
    // extension belong to one tile, which also derives and applies the right-edge term.  Warp 0 runs the two halo
    // half-passes (causal over pass 0, anti-causal over pass CG_NP-1), warps 1..7 one storing pass each.
    const bool last_tile = tile == tiles_per_chunk - 1;
    const long mt0 = n - (long)(tiles_per_chunk - tile) * CG_NS * CF_P;          // first sample of the storing range
    const bool store = w > 0;
    const int passB = store ? w : CG_NP - 1;
    const long m0 = mt0 - CF_P + (long)w * CF_P + (long)lane * CF_R;
    const long m0B = mt0 - CF_P + (long)passB * CF_P + (long)lane * CF_R;
    // ext[M-1] is sample m = n + 14: lane 1, r = 6 of the right-halo pass of the last tile (run by warp 0)
    const bool edge_lane = last_tile && w == 0 && lane == 1;
    C FeI[2], FeQ[2];

    C xs[CF_R];
    cg_load<TI, T>(xc, n, m0, W, fs, xs);
    T yI[CF_R], yQ[CF_R];
    const T c0p = (T)K.c0p;
#pragma unroll
    for (int r = 0; r < CF_R; ++r) { yI[r] = c0p * xs[r].x; yQ[r] = c0p * xs[r].y; }
    cg_dir<T, true>(xs, yI, yQ, lane, K, carryF[w], FeI, FeQ, false);
    if (!store) {
        cg_load<TI, T>(xc, n, m0B, W, fs, xs);
        if (last_tile) cg_dir<T, true>(xs, yI, yQ, lane, K, carryF[CG_NP - 1], FeI, FeQ, edge_lane);   // F[M-1] lives in the right halo
    }
    cg_dir<T, false>(xs, yI, yQ, lane, K, carryB[passB], FeI, FeQ, false);
    __syncthreads();
    if (last_tile) {
        if (edge_lane) {
            // F_j[M-1] = local + p^15 * (end state of the last storing pass);  kappa_i = ka_i x_last + sum_j kM F_j + kMc conj F_j
            for (int m = 0; m < 2; ++m) {
                C p15 = cxld<T>(K.p[m]);
                const C pm = p15;
                for (int i = 0; i < 14; ++i) p15 = cxmul<T>(p15, pm);
                FeI[m] = cxfma<T>(p15, carryF[CG_NP - 2][0][m], FeI[m]);
                FeQ[m] = cxfma<T>(p15, carryF[CG_NP - 2][1][m], FeQ[m]);
            }
            for (int i = 0; i < 2; ++i) {
                C kI = C{(T)K.ka[i][0] * xs[6].x, (T)K.ka[i][1] * xs[6].x}, kQ = C{(T)K.ka[i][0] * xs[6].y, (T)K.ka[i][1] * xs[6].y};
                for (int j = 0; j < 2; ++j) {
                    kI = cxfma<T>(cxld<T>(K.kM[i][j]), FeI[j], kI);
                    kI = cxfma<T>(cxld<T>(K.kMc[i][j]), C{FeI[j].x, -FeI[j].y}, kI);
                    kQ = cxfma<T>(cxld<T>(K.kM[i][j]), FeQ[j], kQ);
                    kQ = cxfma<T>(cxld<T>(K.kMc[i][j]), C{FeQ[j].x, -FeQ[j].y}, kQ);
                }
                kappa[0][i] = kI; kappa[1][i] = kQ;
            }
        }
        __syncthreads();
    }
    if (store) {
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            C LF = C{T(1), T(0)}, LB = C{T(1), T(0)};
#pragma unroll
            for (int s = 0; s < 5; ++s) {
                C g = cxld<T>(K.ppow[s][m]);
                if (lane & (1 << s)) LF = cxmul<T>(LF, g);
                if ((31 - lane) & (1 << s)) LB = cxmul<T>(LB, g);
            }
            C gFI = cxmul<T>(LF, carryF[w - 1][0][m]), gFQ = cxmul<T>(LF, carryF[w - 1][1][m]);
            C gBI = cxmul<T>(LB, carryB[w + 1][0][m]), gBQ = cxmul<T>(LB, carryB[w + 1][1][m]);
#pragma unroll
            for (int r = 0; r < CF_R; ++r) {
                const C af = cxld<T>(K.aF[r][m]), ab = cxld<T>(K.aB[r][m]);
                yI[r] += T(2) * (af.x * gFI.x - af.y * gFI.y + ab.x * gBI.x - ab.y * gBI.y);
                yQ[r] += T(2) * (af.x * gFQ.x - af.y * gFQ.y + ab.x * gBQ.x - ab.y * gBQ.y);
            }
        }
        if (last_tile && n - 1 - m0 < CF_FIX + CF_R) {
            // y[m] -= 2 Re sum_i kappa_i p_i^(16 + mm),  mm = n-1-m   (only the last CF_FIX samples matter)
#pragma unroll
            for (int r = 0; r < CF_R; ++r) {
                const long mm = n - 1 - (m0 + r);
                if (mm < 0 || mm >= CF_FIX) continue;
                const int a = (int)(mm / CF_R), b = (int)(mm - (long)a * CF_R);
                for (int m = 0; m < 2; ++m) {
                    C g = cxld<T>(K.pc[m]);
                    for (int sft = 0; sft < 6; ++sft) if (a & (1 << sft)) g = cxmul<T>(g, cxld<T>(K.ppow[sft][m]));
                    const C pm = cxld<T>(K.p[m]);
                    for (int i = 0; i < b; ++i) g = cxmul<T>(g, pm);
                    yI[r] -= T(2) * (kappa[0][m].x * g.x - kappa[0][m].y * g.y);
                    yQ[r] -= T(2) * (kappa[1][m].x * g.x - kappa[1][m].y * g.y);
                }
            }
        }
        double2 *out = reinterpret_cast<double2 *>(filtered) + (long)chunk * out_stride;
        T *pw = pws + (w - 1) * CF_P + lane * CF_R;
#pragma unroll
        for (int r = 0; r < CF_R; ++r) {
            const long m = m0 + r;
            const bool ok = m >= 0 && m < n;
            if (ok) out[m] = make_double2((double)yI[r], (double)yQ[r]);
            if (phase_partial) pw[r] = ok ? yI[r] * yI[r] + yQ[r] * yQ[r] : T(0);
        }
    }
    // ---- per-phase power partial sums for the timing pick (extract_symbols, processor.py:196-210): 16 threads per
    // phase walk the tile's |y|^2 in shared memory (fixed order), a 16-lane shuffle tree finishes the sum.
    if (phase_partial) {
        __syncthreads();
        const int grp = threadIdx.x >> 4, seg = threadIdx.x & 15;
        const int ph = grp * step;
        double t = 0.0;
        if (ph < sps) {
            const long lo = mt0 > 0 ? mt0 : 0, hi = mt0 + (long)CG_NS * CF_P;
            const long first = lo + ((ph - lo % sps) % sps + sps) % sps;     // first sample of this phase in the tile
            const long kmax = (n - ph) / sps;                                // symbols of this phase: k < kmax
            long k = first / sps + seg;
            for (long m = first + (long)seg * sps; m < hi && k < kmax; m += 16L * sps, k += 16) t += (double)pws[m - mt0];
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (ph < sps && seg == 0) phase_partial[((long)chunk * tiles_per_chunk + tile) * TB200_MAX_PHASES + ph] = t;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// float64 kernel of the "fast" mode.  Same tiling as chan_fast_kernel<TI, double> above (which stays for float working
// precision on complex128 input) with the instruction count roughly halved: every recurrence step and every output
// update is a chain of DFMAs on pre-doubled weights, the halo warp carries no outputs, the in-pass (lane) carry and the
// cross-pass (warp) carry are summed first and applied to the 8 outputs once, lane powers come from a shared-memory
// table and the timing-phase sums index in 32 bits without divisions.
__device__ __forceinline__ void d_step(cdd &s, const cdd p, const double x)          // s = p s + x
{
    const double re = fma(p.x, s.x, fma(-p.y, s.y, x));
    const double im = fma(p.x, s.y, p.y * s.x);
    s.x = re; s.y = im;
}
__device__ __forceinline__ double d_reacc(const cdd a, const cdd s, const double y) { return fma(a.x, s.x, fma(-a.y, s.y, y)); }   // y + Re(a s)

// The part of a lane's work that involves only its own CF_R samples is written as plain sums with constant weights
// instead of the one-pole recurrences: the 8 outputs' zero-state response y[r] = sum_j G[|r-j|] x[j] (a symmetric
// real Toeplitz product) and the complex lane aggregates sum_j p^(7-j) x[j] / sum_j p^j x[j] of both directions and
// modes.  Same arithmetic to rounding, a third fewer DFMAs, and no serial dependency chain of 16 DFMAs per
// direction — the kernel is latency bound on exactly those chains.
__device__ __forceinline__ void c64_outputs(const double (&xs)[CF_R], double (&y)[CF_R], const ChanFastParams &K)
{
#pragma unroll
    for (int r = 0; r < CF_R; ++r) {
        double acc = K.G[0] * xs[r];
#pragma unroll
        for (int j = 0; j < CF_R; ++j) if (j != r) acc = fma(K.G[j > r ? j - r : r - j], xs[j], acc);
        y[r] = acc;
    }
}
template <bool FWD>
__device__ __forceinline__ void c64_aggregate(const double (&xs)[CF_R], const ChanFastParams &K, cdd (&s)[2])
{
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        double re = 0.0, im = 0.0;
#pragma unroll
        for (int j = 0; j < CF_R; ++j) {
            const int k = FWD ? CF_R - 1 - j : j;
            re = fma(K.PW[m][k][0], xs[j], re);
            im = fma(K.PW[m][k][1], xs[j], im);
        }
        s[m] = cdd{re, im};
    }
}

// Kogge-Stone scan of the lane aggregates of one direction; returns the lane's in-pass carry-in (exclusive), the
// pass-end state goes to carry[mode].  EDGE: Fe = zero-state F at r = 6 of this lane + p^7 * carry-in (right-halo pass).
template <bool FWD, bool EDGE>
__device__ __forceinline__ void c64_scan(const double (&xs)[CF_R], cdd (&s)[2], int lane, const ChanFastParams &K,
                                         cdd *carry, cdd (&cin)[2], cdd *Fe, bool edge_lane)
{
    if (EDGE) {
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            double re = 0.0, im = 0.0;
#pragma unroll
            for (int j = 0; j < CF_R - 1; ++j) { re = fma(K.PW[m][CF_R - 2 - j][0], xs[j], re); im = fma(K.PW[m][CF_R - 2 - j][1], xs[j], im); }
            Fe[m] = cdd{re, im};
        }
    }
#pragma unroll
    for (int m = 0; m < 2; ++m) {
#pragma unroll
        for (int st = 0; st < 5; ++st) {
            const cdd o = FWD ? dshfl_up(s[m], 1 << st) : dshfl_down(s[m], 1 << st);
            if (FWD ? (lane >= (1 << st)) : (lane + (1 << st) < 32)) s[m] = cfmad(ldd(K.ppow[st][m]), o, s[m]);
        }
        cin[m] = FWD ? dshfl_up(s[m], 1) : dshfl_down(s[m], 1);
        if (lane == (FWD ? 0 : 31)) cin[m] = cdd{0.0, 0.0};
        if (lane == (FWD ? 31 : 0)) carry[m] = s[m];
        if (EDGE && edge_lane) Fe[m] = cfmad(ldd(K.p7[m]), cin[m], Fe[m]);     // + p^7 * (carry from lane 0 of this warp)
    }
}

// component c (0 = I, 1 = Q) of the lane's CF_R samples
template <typename TI>
__device__ __noinline__ double c64_padded(const TI *__restrict__ xc, long n, long t, double W, double fs, int c)
{
    const cdd v = chan_padded<TI>(xc, n, t, W, fs);
    return c ? v.y : v.x;
}
template <typename TI>
__device__ __forceinline__ void c64_load(const TI *__restrict__ xc, long n, long m0, double W, double fs, int c, double (&xs)[CF_R])
{
    if (W == 0.0 && m0 >= 0 && m0 + CF_R <= n) {
#pragma unroll
        for (int r = 0; r < CF_R; ++r) xs[r] = (double)xc[2 * (m0 + r) + c];
    } else {
#pragma unroll
        for (int r = 0; r < CF_R; ++r) xs[r] = c64_padded<TI>(xc, n, 15 + m0 + r, W, fs, c);
    }
}

constexpr int C64_TILE = CG_NS * CF_P;                       // samples a CTA stores
__host__ __device__ constexpr int c64_pad(int i) { return i + (i >> 4); }    // staging index: lane stride 8 doubles -> no bank conflicts

// CTA = 16 warps: warp (pass, component).  I and Q never mix in the filter, so each warp carries one real signal: half
// the registers of a (pass, I+Q) warp, twice the warps in flight for the same work — the kernel is bound by the latency
// of dependent DFMAs, not by their count.  Results meet in shared memory: coalesced double2 stores and |y|^2.
template <typename TI>
__global__ void __launch_bounds__(2 * CF_W * 32, 2)
chan_fast64_kernel(const TI *__restrict__ x, long n, long in_stride, int tiles_per_chunk,
                   const double *__restrict__ freq_off, double fs, ChanFastParams K,
                   double *__restrict__ filtered, long out_stride, double *__restrict__ phase_partial, int sps, int step)
{
    __shared__ cdd carryF[CG_NP][2][2];       // [pass][I/Q][mode]
    __shared__ cdd carryB[CG_NP][2][2];
    __shared__ cdd kappa[2][2];               // [I/Q][mode] right-edge transient amplitudes of this chunk (last tile only)
    __shared__ cdd lanepow[32][2];
    __shared__ double ys[2][c64_pad(C64_TILE) + 1];          // outputs of the tile per component, then |y|^2 in ys[0]
    if (threadIdx.x < 64) lanepow[threadIdx.x >> 1][threadIdx.x & 1] = ldd(K.lanepow[threadIdx.x >> 1][threadIdx.x & 1]);
    const int chunk = blockIdx.x / tiles_per_chunk;
    const int tile = blockIdx.x - chunk * tiles_per_chunk;
    const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w = wi >> 1, c = wi & 1;        // pass, component
    const TI *xc = x + (long)chunk * in_stride * 2;
    const double f = freq_off ? freq_off[chunk] : 0.0;
    const double W = (f != 0.0) ? __dmul_rn(-6.283185307179586, f) : 0.0;
    // tiling: see chan_fast_kernel
    const bool last_tile = tile == tiles_per_chunk - 1;
    const long mt0 = n - (long)(tiles_per_chunk - tile) * C64_TILE;
    const bool store = w > 0;
    const int passB = store ? w : CG_NP - 1;
    const long m0 = mt0 - CF_P + (long)w * CF_P + (long)lane * CF_R;
    const long m0B = mt0 - CF_P + (long)passB * CF_P + (long)lane * CF_R;
    const bool edge_lane = last_tile && w == 0 && lane == 1;
    cdd Fe[2], cF[2], cB[2];

    // A causal pass that lies entirely left of the odd extension (sample index < -15) sees the constant continuation
    // only: its end state is the steady state c / (1 - p) and nothing of it is stored — the first, partial tile of a
    // chunk skips those passes (tiles are aligned to the END of the chunk: 13108 samples = 3.2 of the first tile's 7).
    const bool const_pass = mt0 - CF_P + (long)w * CF_P + CF_P - 1 < -15;
    double xs[CF_R], y[CF_R];
    if (const_pass) {
        if (lane == 31) {
            const double c0 = c64_padded<TI>(xc, n, -1, W, fs, c);
            carryF[w][c][0] = cdd{K.inv1mp[0][0] * c0, K.inv1mp[0][1] * c0};
            carryF[w][c][1] = cdd{K.inv1mp[1][0] * c0, K.inv1mp[1][1] * c0};
        }
#pragma unroll
        for (int r = 0; r < CF_R; ++r) { xs[r] = 0.0; y[r] = 0.0; }
    } else c64_load<TI>(xc, n, m0, W, fs, c, xs);
    cdd sF[2], sB[2];
    if (store && const_pass) {
        cF[0] = cF[1] = cB[0] = cB[1] = cdd{0.0, 0.0};
    } else if (store) {
        c64_outputs(xs, y, K);
        c64_aggregate<true>(xs, K, sF);
        c64_aggregate<false>(xs, K, sB);
        c64_scan<true, false>(xs, sF, lane, K, carryF[w][c], cF, Fe, false);
        c64_scan<false, false>(xs, sB, lane, K, carryB[w][c], cB, Fe, false);
    } else {
        if (!const_pass) {
            c64_aggregate<true>(xs, K, sF);
            c64_scan<true, false>(xs, sF, lane, K, carryF[0][c], cF, Fe, false);
        }
        c64_load<TI>(xc, n, m0B, W, fs, c, xs);
        if (last_tile) {                               // F[M-1] lives in the right halo
            c64_aggregate<true>(xs, K, sF);
            c64_scan<true, true>(xs, sF, lane, K, carryF[CG_NP - 1][c], cF, Fe, edge_lane);
        }
        c64_aggregate<false>(xs, K, sB);
        c64_scan<false, false>(xs, sB, lane, K, carryB[CG_NP - 1][c], cB, Fe, false);
    }
    __syncthreads();
    if (last_tile) {
        if (edge_lane) {
            // F_j[M-1] = local + p^15 * (end state of the last storing pass);  kappa_i = ka_i x_last + sum_j kM F_j + kMc conj F_j
            for (int m = 0; m < 2; ++m) Fe[m] = cfmad(ldd(K.p15[m]), carryF[CG_NP - 2][c][m], Fe[m]);
            for (int i = 0; i < 2; ++i) {
                cdd k = cdd{K.ka[i][0] * xs[6], K.ka[i][1] * xs[6]};
                for (int j = 0; j < 2; ++j) {
                    k = cfmad(ldd(K.kM[i][j]), Fe[j], k);
                    k = cfmad(ldd(K.kMc[i][j]), cdd{Fe[j].x, -Fe[j].y}, k);
                }
                kappa[c][i] = k;
            }
        }
        __syncthreads();
    }
    if (store && !const_pass) {
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            // total carry into this lane: in-pass (from the scan) + lane power * end state of the neighbouring pass
            const cdd gF = cfmad(lanepow[lane][m], carryF[w - 1][c][m], cF[m]);
            const cdd gB = cfmad(lanepow[31 - lane][m], carryB[w + 1][c][m], cB[m]);
#pragma unroll
            for (int r = 0; r < CF_R; ++r) y[r] = d_reacc(ldd(K.aF2[r][m]), gF, d_reacc(ldd(K.aB2[r][m]), gB, y[r]));
        }
        const int mi = (int)m0, ni = (int)n;                     // n < 2^30 (launch_t): 32-bit index tests
        if (last_tile && ni - 1 - mi < CF_FIX + CF_R) {
            // y[m] -= 2 Re sum_i kappa_i p_i^(16 + mm),  mm = n-1-m   (only the last CF_FIX samples matter)
#pragma unroll
            for (int r = 0; r < CF_R; ++r) {
                const int mm = ni - 1 - (mi + r);
                if (mm < 0 || mm >= CF_FIX) continue;
                const int a = mm / CF_R, b = mm - a * CF_R;
                for (int m = 0; m < 2; ++m) {
                    cdd g = ldd(K.pc[m]);
                    for (int sft = 0; sft < 6; ++sft) if (a & (1 << sft)) g = cmuld(g, ldd(K.ppow[sft][m]));
                    const cdd pm = ldd(K.p[m]);
                    for (int i = 0; i < b; ++i) g = cmuld(g, pm);
                    y[r] -= 2.0 * (kappa[c][m].x * g.x - kappa[c][m].y * g.y);
                }
            }
        }
        const int i0 = (w - 1) * CF_P + lane * CF_R;
#pragma unroll
        for (int r = 0; r < CF_R; ++r) ys[c][c64_pad(i0 + r)] = y[r];
    }
    __syncthreads();
    // ---- coalesced complex128 stores; |y|^2 replaces the I component in shared memory
    {
        double2 *out = reinterpret_cast<double2 *>(filtered) + (long)chunk * out_stride;
        const int ni = (int)n, t0 = (int)mt0;
        for (int i = threadIdx.x; i < C64_TILE; i += 2 * CF_W * 32) {
            const int m = t0 + i, ip = c64_pad(i);
            const double yi = ys[0][ip], yq = ys[1][ip];
            const bool ok = m >= 0 && m < ni;
            if (ok) out[m] = make_double2(yi, yq);
            if (phase_partial) ys[0][ip] = ok ? fma(yi, yi, yq * yq) : 0.0;
        }
    }
    // ---- per-phase power partial sums for the timing pick (extract_symbols, processor.py:196-210): 16 threads per
    // phase walk the tile's |y|^2 in shared memory (fixed order), a 16-lane shuffle tree finishes the sum.
    if (phase_partial) {
        __syncthreads();
        const int grp = threadIdx.x >> 4, seg = threadIdx.x & 15;
        const int ph = grp * step;
        double t = 0.0;
        if (ph < sps) {
            const int ni = (int)n, t0 = (int)mt0, lo = t0 > 0 ? t0 : 0, hi = t0 + C64_TILE;
            int d = ph - lo % sps;
            if (d < 0) d += sps;                                             // first sample of this phase in the tile: lo + d
            // symbols of this phase are k < (n - ph) / sps, i.e. samples m = ph + sps k with m + sps <= n
            for (int m = lo + d + seg * sps; m < hi && m + sps <= ni; m += 16 * sps) t += ys[0][c64_pad(m - t0)];
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (ph < sps && seg == 0) phase_partial[((long)chunk * tiles_per_chunk + tile) * TB200_MAX_PHASES + ph] = t;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// float32 packed variant (TB200_OPT_CHAN_F32, complex64 input).  Same two-sided modal form; the I and Q components
// share every coefficient, so all recurrences, scans and carries run on float2 (I, Q) pairs with fma.rn.f32x2
// (SASS FFMA2), Kogge-Stone stages whose multiplier (p^R)^(2^s) is below 1e-9 are skipped, the lane powers of the
// cross-warp carry come from a shared-memory table, and the 13 timing-phase power sums are taken from a
// shared-memory copy of |y|^2 by 16 threads per phase (fixed order) instead of 13 warp reductions per warp.
struct ChanFast32Params {
    float p[2][2];
    float ppow[6][2][2];           // (p^R)^(2^s)
    float c0p;
    float a2[2][2];                // 2 alpha_i
    float aF2[CF_R][2][2];         // 2 alpha_i p^(r+1)
    float aB2[CF_R][2][2];         // 2 alpha_i p^(R-r)
    float lanepow[32][2][2];       // (p^R)^l
    float ka[2][2], kM[2][2][2], kMc[2][2][2];
    float pc[2][2];                // p^16
    float p7[2][2], p15[2][2];
    int nstage[2];                 // Kogge-Stone stages that matter for mode i
    float G[CF_R];                 // y[r] = sum_j G[|r - j|] x[j]: zero-state response inside a lane (G[0] includes c0p)
    float inv1mp[2][2];            // 1 / (1 - p_i): steady state of a constant input
    float PW[2][CF_R][2];          // p_i^k, k < R
};

constexpr int C32_NS = CG_NS;
constexpr int C32_NP = CG_NP;

// in-lane work as constant-weight sums (see c64_outputs): outputs' zero-state response and the lane aggregates
__device__ __forceinline__ void c32_outputs(const float2 (&xs)[CF_R], float2 (&y)[CF_R], const ChanFast32Params &K)
{
#pragma unroll
    for (int r = 0; r < CF_R; ++r) {
        float2 acc = make_float2(K.G[0] * xs[r].x, K.G[0] * xs[r].y);
#pragma unroll
        for (int j = 0; j < CF_R; ++j) if (j != r) acc = ffma2(splat(K.G[j > r ? j - r : r - j]), xs[j], acc);
        y[r] = acc;
    }
}
template <bool FWD, int NT>
__device__ __forceinline__ void c32_aggregate(const float2 (&xs)[CF_R], const ChanFast32Params &K, cp (&s)[2])
{
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        float2 re = make_float2(0.f, 0.f), im = re;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int k = FWD ? NT - 1 - j : j;
            re = ffma2(splat(K.PW[m][k][0]), xs[j], re);
            im = ffma2(splat(K.PW[m][k][1]), xs[j], im);
        }
        s[m].re = re; s[m].im = im;
    }
}

// one direction: Kogge-Stone scan of the lane aggregates, in-warp carry applied to y; the pass-end state goes to
// carry[0..1].  EDGE: Fe = zero-state F at r = 6 of this lane + p^7 * carry-in (right-halo pass of the last tile).
template <bool FWD, bool EDGE>
__device__ __forceinline__ void c32_dir(const float2 (&xs)[CF_R], float2 (&y)[CF_R], int lane, const ChanFast32Params &K,
                                        cp *carry, cp *Fe, bool edge_lane)
{
    const float2 zero2 = make_float2(0.f, 0.f);
    cp s[2];
    c32_aggregate<FWD, CF_R>(xs, K, s);
    if (EDGE) {
        cp f6[2];
        c32_aggregate<true, CF_R - 1>(xs, K, f6);
        Fe[0] = f6[0]; Fe[1] = f6[1];
    }
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        cp a = s[m];
#pragma unroll
        for (int st = 0; st < 5; ++st) {
            // (a uniform test on K.nstage[m] would skip the last stages of the fast-decaying mode, but the branch
            // costs more issue slots than the 10 instructions it saves)
            cp o;
            o.re = FWD ? shfl2_up(a.re, 1 << st) : shfl2_down(a.re, 1 << st);
            o.im = FWD ? shfl2_up(a.im, 1 << st) : shfl2_down(a.im, 1 << st);
            cf g = ld2(K.ppow[st][m]);
            if (FWD ? (lane < (1 << st)) : (lane + (1 << st) >= 32)) g.x = g.y = 0.f;
            a = cpfma(g, o, a);
        }
        cp cin;
        cin.re = FWD ? shfl2_up(a.re, 1) : shfl2_down(a.re, 1);
        cin.im = FWD ? shfl2_up(a.im, 1) : shfl2_down(a.im, 1);
        if (lane == (FWD ? 0 : 31)) cin.re = cin.im = zero2;
        if (lane == (FWD ? 31 : 0)) carry[m] = a;
        if (EDGE && edge_lane) Fe[m] = cpfma(ld2(K.p7[m]), cin, Fe[m]);   // + p^7 * (carry from lane 0 of this warp)
#pragma unroll
        for (int r = 0; r < CF_R; ++r) y[r] = re_acc(FWD ? ld2(K.aF2[r][m]) : ld2(K.aB2[r][m]), cin, y[r]);
    }
}

// out-of-line: edge lanes and NCO'd chunks only; keeps the hot path's code small (instruction fetch)
__device__ __noinline__ float2 c32_padded(const float *__restrict__ xc, long n, long t, double W, double fs)
{
    cdd v = chan_padded<float>(xc, n, t, W, fs);
    return make_float2((float)v.x, (float)v.y);
}

// NCO of an interior lane (frequency_shift, processor.py:85-100): phase of the lane's first sample exactly as the
// reference forms it (float64: fl(-2 pi f) * fl(i / fs)), then a float32 rotation by the per-sample step for the other
// seven samples (error ~4e-7 of full scale; two double sincos per lane instead of eight).
__device__ __noinline__ float4 c32_nco(long m0, double W, double fs)
{
    double s0, c0, sd, cdv;
    sincos(__dmul_rn(W, (double)m0 / fs), &s0, &c0);
    sincos(W / fs, &sd, &cdv);
    return make_float4((float)c0, (float)s0, (float)cdv, (float)sd);
}

__device__ __forceinline__ void c32_load(const float *__restrict__ xc, long n, long m0, double W, double fs, float2 (&xs)[CF_R])
{
    if (m0 >= 0 && m0 + CF_R <= n && ((reinterpret_cast<uintptr_t>(xc + 2 * m0) & 15) == 0)) {
        // interior lane: 8 samples = 64 contiguous bytes = four 16-byte loads (tiles are end-aligned)
        const float4 *v4 = reinterpret_cast<const float4 *>(xc + 2 * m0);
#pragma unroll
        for (int r = 0; r < CF_R / 2; ++r) {
            const float4 v = v4[r];
            xs[2 * r] = make_float2(v.x, v.y);
            xs[2 * r + 1] = make_float2(v.z, v.w);
        }
        if (W != 0.0) {
            const float4 rot = c32_nco(m0, W, fs);
            float c = rot.x, sn = rot.y;
#pragma unroll
            for (int r = 0; r < CF_R; ++r) {
                const float xr = xs[r].x, xi = xs[r].y;
                xs[r] = make_float2(xr * c - xi * sn, xr * sn + xi * c);
                const float c2 = c * rot.z - sn * rot.w;
                sn = sn * rot.z + c * rot.w;
                c = c2;
            }
        }
    } else {
#pragma unroll
        for (int r = 0; r < CF_R; ++r) xs[r] = c32_padded(xc, n, 15 + m0 + r, W, fs);
    }
}

__global__ void __launch_bounds__(CF_W * 32, 4)
chan_fast32_kernel(const float *__restrict__ x, long n, long in_stride, int tiles_per_chunk,
                   const double *__restrict__ freq_off, double fs, ChanFast32Params K,
                   double *__restrict__ filtered, long out_stride, double *__restrict__ phase_partial, int sps, int step)
{
    __shared__ cp carryF[C32_NP][2];
    __shared__ cp carryB[C32_NP][2];
    __shared__ cp kappa[2];
    __shared__ cf lanepow[32][2];
    __shared__ float pws[C32_NS * CF_P + C32_NS * CF_P / 32];      // index i + (i >> 5): lane stride 8 floats -> no bank conflicts
    if (threadIdx.x < 64) lanepow[threadIdx.x >> 1][threadIdx.x & 1] = ld2(K.lanepow[threadIdx.x >> 1][threadIdx.x & 1]);
    const int chunk = blockIdx.x / tiles_per_chunk;
    const int tile = blockIdx.x - chunk * tiles_per_chunk;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float *xc = x + (long)chunk * in_stride * 2;
    const double f = freq_off ? freq_off[chunk] : 0.0;
    const double W = (f != 0.0) ? __dmul_rn(-6.283185307179586, f) : 0.0;
    // tiles are aligned to the END of the chunk (the first may be partial): the last CF_FIX samples and the right
    // extension belong to one tile, which also derives and applies the right-edge term
    const bool last_tile = tile == tiles_per_chunk - 1;
    const long mt0 = n - (long)(tiles_per_chunk - tile) * C32_NS * CF_P;          // first sample of the storing range
    const bool store = w > 0;
    const int passB = store ? w : C32_NP - 1;
    const long m0 = mt0 - CF_P + (long)w * CF_P + (long)lane * CF_R;             // causal pass of this warp
    const long m0B = mt0 - CF_P + (long)passB * CF_P + (long)lane * CF_R;        // anti-causal pass
    // ext[M-1] is sample m = n + 14: lane 1, r = 6 of the right-halo pass of the last tile (run by warp 0)
    const bool edge_lane = last_tile && w == 0 && lane == 1;

    // a causal pass entirely left of the odd extension sees the constant continuation only (see chan_fast64_kernel)
    const bool const_pass = mt0 - CF_P + (long)w * CF_P + CF_P - 1 < -15;
    float2 xs[CF_R], y[CF_R];
    cp Fe[2];
    if (const_pass) {
        if (lane == 31) {
            const float2 c0 = c32_padded(xc, n, -1, W, fs);
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                carryF[w][m].re = make_float2(K.inv1mp[m][0] * c0.x, K.inv1mp[m][0] * c0.y);
                carryF[w][m].im = make_float2(K.inv1mp[m][1] * c0.x, K.inv1mp[m][1] * c0.y);
            }
        }
#pragma unroll
        for (int r = 0; r < CF_R; ++r) { xs[r] = make_float2(0.f, 0.f); y[r] = make_float2(0.f, 0.f); }
    } else {
        c32_load(xc, n, m0, W, fs, xs);
        if (store) c32_outputs(xs, y, K);
        else {
#pragma unroll
            for (int r = 0; r < CF_R; ++r) y[r] = make_float2(0.f, 0.f);        // the halo warp stores nothing
        }
        c32_dir<true, false>(xs, y, lane, K, carryF[w], Fe, false);
    }
    if (!store) {
        c32_load(xc, n, m0B, W, fs, xs);
        if (last_tile) c32_dir<true, true>(xs, y, lane, K, carryF[C32_NP - 1], Fe, edge_lane);   // F[M-1] lives in the right halo
    }
    if (!(store && const_pass)) c32_dir<false, false>(xs, y, lane, K, carryB[passB], Fe, false);
    __syncthreads();
    if (last_tile) {
        if (edge_lane) {
            // F_j[M-1] = local + p^15 * (end state of the last storing pass);  kappa_i = ka_i x_last + sum_j kM F_j + kMc conj F_j
#pragma unroll
            for (int m = 0; m < 2; ++m) Fe[m] = cpfma(ld2(K.p15[m]), carryF[C32_NP - 2][m], Fe[m]);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                cp k;
                k.re = make_float2(K.ka[i][0] * xs[6].x, K.ka[i][0] * xs[6].y);
                k.im = make_float2(K.ka[i][1] * xs[6].x, K.ka[i][1] * xs[6].y);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    k = cpfma(ld2(K.kM[i][j]), Fe[j], k);
                    cp cj; cj.re = Fe[j].re; cj.im = make_float2(-Fe[j].im.x, -Fe[j].im.y);
                    k = cpfma(ld2(K.kMc[i][j]), cj, k);
                }
                kappa[i] = k;
            }
        }
        __syncthreads();
    }
    if (store && !const_pass) {                        // (a constant pass covers only negative sample indices: nothing to store or sum)
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            const cp gF = cpmul(lanepow[lane][m], carryF[w - 1][m]);
            const cp gB = cpmul(lanepow[31 - lane][m], carryB[w + 1][m]);
#pragma unroll
            for (int r = 0; r < CF_R; ++r) y[r] = re_acc(ld2(K.aB2[r][m]), gB, re_acc(ld2(K.aF2[r][m]), gF, y[r]));
        }
        if (last_tile && n - 1 - m0 < CF_FIX + CF_R) {
            // y[m] -= 2 Re sum_i kappa_i p_i^(16 + mm),  mm = n-1-m   (only the last CF_FIX samples matter)
#pragma unroll
            for (int r = 0; r < CF_R; ++r) {
                const long mm = n - 1 - (m0 + r);
                if (mm < 0 || mm >= CF_FIX) continue;
                const int a = (int)(mm / CF_R), b = (int)(mm - (long)a * CF_R);
                for (int m = 0; m < 2; ++m) {
                    cf g = ld2(K.pc[m]);
                    for (int sft = 0; sft < 6; ++sft) if (a & (1 << sft)) g = cmulf(g, ld2(K.ppow[sft][m]));
                    const cf pm = ld2(K.p[m]);
                    for (int i = 0; i < b; ++i) g = cmulf(g, pm);
                    y[r] = re_acc(cf{-2.0f * g.x, -2.0f * g.y}, kappa[m], y[r]);
                }
            }
        }
        double2 *out = reinterpret_cast<double2 *>(filtered) + (long)chunk * out_stride + m0;
        const int mi = (int)m0, ni = (int)n;                     // n < 2^30 (launch_t): 32-bit index tests
        float *pw = pws + (w - 1) * CF_P + lane * CF_R + (((w - 1) * CF_P + lane * CF_R) >> 5);     // 8 samples of a lane share i >> 5
        if (mi >= 0 && mi + CF_R <= ni) {
#pragma unroll
            for (int r = 0; r < CF_R; ++r) {
                out[r] = make_double2((double)y[r].x, (double)y[r].y);
                if (phase_partial) pw[r] = y[r].x * y[r].x + y[r].y * y[r].y;
            }
        } else {
#pragma unroll
            for (int r = 0; r < CF_R; ++r) {
                const bool ok = mi + r >= 0 && mi + r < ni;
                if (ok) out[r] = make_double2((double)y[r].x, (double)y[r].y);
                if (phase_partial) pw[r] = ok ? y[r].x * y[r].x + y[r].y * y[r].y : 0.f;
            }
        }
    }
    // ---- per-phase power partial sums for the timing pick (extract_symbols, processor.py:196-210): 16 threads per
    // phase walk the tile's |y|^2 in shared memory (fixed order), a 16-lane shuffle tree finishes the sum.
    if (phase_partial) {
        __syncthreads();
        const int grp = threadIdx.x >> 4, seg = threadIdx.x & 15;
        const int ph = grp * step;
        double t = 0.0;
        if (ph < sps) {
            const int ni = (int)n, t0 = (int)mt0, lo = t0 > 0 ? t0 : 0, hi = t0 + C32_NS * CF_P;
            int d = ph - lo % sps;
            if (d < 0) d += sps;                                             // first sample of this phase in the tile: lo + d
            // symbols of this phase are k < (n - ph) / sps, i.e. samples m = ph + sps k with m + sps <= n
            for (int m = lo + d + seg * sps; m < hi && m + sps <= ni; m += 16 * sps) { const int i = m - t0; t += (double)pws[i + (i >> 5)]; }
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (ph < sps && seg == 0) phase_partial[((long)chunk * tiles_per_chunk + tile) * TB200_MAX_PHASES + ph] = t;
    }
}

// ---- host: modal constants from the Butterworth poles ---------------------------------------
typedef std::complex<double> cd;
static void putd(double dst[2], cd v) { dst[0] = v.real(); dst[1] = v.imag(); }

int channel_filter_poles(double bandwidth, double fs, std::complex<double> poles[4], double *gain);

static int make_chan_params(double bandwidth, double fs, ChanFastParams &K)
{
    cd pz[4];
    double k;
    int rc = channel_filter_poles(bandwidth, fs, pz, &k);
    if (rc) return rc;
    // order: the two poles with positive imaginary part first, then their conjugates
    cd p4[4];
    int np = 0;
    for (int i = 0; i < 4; ++i) if (pz[i].imag() > 0) p4[np++] = pz[i];
    if (np != 2) return fail(TB200_E_ARG, "fast channel filter: expected two complex pole pairs");
    p4[2] = std::conj(p4[0]); p4[3] = std::conj(p4[1]);
    cd r4[4];
    cd prodp(1.0, 0.0);
    for (int i = 0; i < 4; ++i) prodp *= p4[i];
    const double dd = k / prodp.real();
    for (int i = 0; i < 4; ++i) {
        cd num = k * std::pow(1.0 + 1.0 / p4[i], 4), den(1.0, 0.0);
        for (int j = 0; j < 4; ++j) if (j != i) den *= (1.0 - p4[j] / p4[i]);
        r4[i] = num / den;
    }
    cd alpha[2];
    cd sumr(0.0, 0.0);
    for (int i = 0; i < 4; ++i) sumr += r4[i];
    double c0 = (dd * dd + dd * sumr).real();
    double sum_re_alpha = 0.0;
    for (int j = 0; j < 2; ++j) {
        cd acc(dd, 0.0);
        for (int i = 0; i < 4; ++i) acc += r4[i] / (1.0 - p4[i] * p4[j]);
        alpha[j] = r4[j] * acc;
        sum_re_alpha += alpha[j].real();
    }
    memset(&K, 0, sizeof(K));
    K.c0p = c0 - 2.0 * sum_re_alpha;
    for (int j = 0; j < 2; ++j) {
        putd(K.p[j], p4[j]);
        cd pr = std::pow(p4[j], CF_R);
        for (int s = 0; s < 6; ++s) { putd(K.ppow[s][j], pr); pr *= pr; }
        putd(K.a[j], alpha[j]);
        for (int r = 0; r < CF_R; ++r) {
            putd(K.aF[r][j], alpha[j] * std::pow(p4[j], r + 1));
            putd(K.aB[r][j], alpha[j] * std::pow(p4[j], CF_R - r));
        }
        putd(K.inv1mp[j], 1.0 / (1.0 - p4[j]));
        putd(K.pc[j], std::pow(p4[j], 16));
        putd(K.a2[j], 2.0 * alpha[j]);
        for (int r = 0; r < CF_R; ++r) {
            putd(K.aF2[r][j], 2.0 * alpha[j] * std::pow(p4[j], r + 1));
            putd(K.aB2[r][j], 2.0 * alpha[j] * std::pow(p4[j], CF_R - r));
        }
        putd(K.p7[j], std::pow(p4[j], 7));
        putd(K.p15[j], std::pow(p4[j], 15));
        const cd pr8 = std::pow(p4[j], CF_R);
        for (int l = 0; l < 32; ++l) putd(K.lanepow[l][j], std::pow(pr8, l));
        for (int k = 0; k < CF_R; ++k) putd(K.PW[j][k], std::pow(p4[j], k));
        int ns = 0;
        for (int s = 0; s < 5; ++s) if (std::hypot(K.ppow[s][j][0], K.ppow[s][j][1]) >= 1e-17) ns = s + 1;
        K.nstage[j] = ns;
    }
    for (int d = 0; d < CF_R; ++d) {
        double g = 0.0;
        for (int j = 0; j < 2; ++j) g += (2.0 * alpha[j] * std::pow(p4[j], d)).real();
        K.G[d] = d ? g : K.c0p + 2.0 * g;              // both directions contain the d = 0 term
    }
    cd H1(dd, 0.0);
    for (int j = 0; j < 4; ++j) H1 += r4[j] / (1.0 - p4[j]);
    for (int i = 0; i < 2; ++i) {
        cd ka = (H1 - dd) / (1.0 - p4[i]);
        cd kM[4];
        for (int j = 0; j < 4; ++j) {
            ka -= r4[j] * p4[j] / ((1.0 - p4[i] * p4[j]) * (1.0 - p4[j]));
            kM[j] = -r4[j] / (1.0 - p4[i]) + r4[j] * p4[j] / (1.0 - p4[i] * p4[j]);
        }
        putd(K.ka[i], r4[i] * ka);
        for (int j = 0; j < 2; ++j) { putd(K.kM[i][j], r4[i] * kM[j]); putd(K.kMc[i][j], r4[i] * kM[j + 2]); }
    }
    return TB200_OK;
}

static void putf(float dst[2], const double src[2]) { dst[0] = (float)src[0]; dst[1] = (float)src[1]; }

static void make_chan32_params(const ChanFastParams &K, ChanFast32Params &F)
{
    memset(&F, 0, sizeof(F));
    F.c0p = (float)K.c0p;
    for (int d = 0; d < CF_R; ++d) F.G[d] = (float)K.G[d];
    for (int j = 0; j < 2; ++j) {
        putf(F.p[j], K.p[j]);
        const cd pj(K.p[j][0], K.p[j][1]), al(K.a[j][0], K.a[j][1]);
        F.a2[j][0] = (float)(2.0 * al.real()); F.a2[j][1] = (float)(2.0 * al.imag());
        int ns = 0;
        for (int s = 0; s < 6; ++s) {
            putf(F.ppow[s][j], K.ppow[s][j]);
            if (s < 5 && std::hypot(K.ppow[s][j][0], K.ppow[s][j][1]) >= 1e-9) ns = s + 1;
        }
        F.nstage[j] = ns;
        for (int r = 0; r < CF_R; ++r) {
            F.aF2[r][j][0] = (float)(2.0 * K.aF[r][j][0]); F.aF2[r][j][1] = (float)(2.0 * K.aF[r][j][1]);
            F.aB2[r][j][0] = (float)(2.0 * K.aB[r][j][0]); F.aB2[r][j][1] = (float)(2.0 * K.aB[r][j][1]);
        }
        const cd pr = std::pow(pj, CF_R);
        for (int l = 0; l < 32; ++l) { cd v = std::pow(pr, l); F.lanepow[l][j][0] = (float)v.real(); F.lanepow[l][j][1] = (float)v.imag(); }
        putf(F.ka[j], K.ka[j]);
        for (int i = 0; i < 2; ++i) { putf(F.kM[j][i], K.kM[j][i]); putf(F.kMc[j][i], K.kMc[j][i]); }
        putf(F.pc[j], K.pc[j]);
        for (int k = 0; k < CF_R; ++k) { F.PW[j][k][0] = (float)K.PW[j][k][0]; F.PW[j][k][1] = (float)K.PW[j][k][1]; }
        F.inv1mp[j][0] = (float)K.inv1mp[j][0]; F.inv1mp[j][1] = (float)K.inv1mp[j][1];
        cd v7 = std::pow(pj, 7), v15 = std::pow(pj, 15);
        F.p7[j][0] = (float)v7.real(); F.p7[j][1] = (float)v7.imag();
        F.p15[j][0] = (float)v15.real(); F.p15[j][1] = (float)v15.imag();
    }
}

static bool use_packed32(tb200_handle h, int dtype, long n) { return h->chan_f32 && dtype == TB200_DTYPE_C64 && n < (1L << 30); }

// tiles per chunk of the kernel launch_channel_fast will pick for (handle options, input dtype, n)
int chan_fast_tiles(tb200_handle h, int dtype, long n)
{
    (void)h; (void)dtype;
    const int per_tile = CG_NS * CF_P;            // both kernels: 7 storing passes per CTA
    return (int)((n + per_tile - 1) / per_tile);
}

template <typename TI>
static int launch_t(tb200_handle h, const void *x, int n_chunks, long n, long in_stride, const double *freq_off,
                    double fs, const ChanFastParams &K, double *filtered, long out_stride, double *phase_partial,
                    double *phase_tail, int sps, int step, cudaStream_t st)
{
    const int tiles = chan_fast_tiles(h, sizeof(TI) == 4 ? TB200_DTYPE_C64 : TB200_DTYPE_C128, n);
    if (use_packed32(h, sizeof(TI) == 4 ? TB200_DTYPE_C64 : TB200_DTYPE_C128, n)) {
        static std::mutex mu32;
        static ChanFast32Params F32;
        static double key[3] = {0, 0, 0};
        ChanFast32Params F;
        {
            std::lock_guard<std::mutex> lk(mu32);
            if (!(key[0] == K.p[0][0] && key[1] == K.p[0][1] && key[2] == K.c0p)) {
                make_chan32_params(K, F32);
                key[0] = K.p[0][0]; key[1] = K.p[0][1]; key[2] = K.c0p;
            }
            F = F32;
        }
        chan_fast32_kernel<<<n_chunks * tiles, CF_W * 32, 0, st>>>((const float *)x, n, in_stride, tiles, freq_off, fs, F, filtered,
                                                                 out_stride, phase_partial, sps, step);
    } else if (h->chan_f32)
        chan_fast_kernel<TI, float><<<n_chunks * tiles, CF_W * 32, 0, st>>>((const TI *)x, n, in_stride, tiles, freq_off, fs, K,
                                                                          filtered, out_stride, phase_partial, sps, step);
    else if (n < (1L << 30))
        chan_fast64_kernel<TI><<<n_chunks * tiles, 2 * CF_W * 32, 0, st>>>((const TI *)x, n, in_stride, tiles, freq_off, fs, K,
                                                                     filtered, out_stride, phase_partial, sps, step);
    else
        chan_fast_kernel<TI, double><<<n_chunks * tiles, CF_W * 32, 0, st>>>((const TI *)x, n, in_stride, tiles, freq_off, fs, K,
                                                                           filtered, out_stride, phase_partial, sps, step);
    TB_LAUNCH_CHECK(h);
    (void)phase_tail;
    return TB200_OK;
}

// The time-parallel form assumes that a pass forgets its past within one halo pass of CF_P samples.  That holds
// for the rates process() runs the channel filter at (240..480 kHz: |p|^256 < 1e-7), not for a 25 kHz filter
// applied directly at a multi-MHz rate (filter_signal as a stand-alone call): those go to the exact kernel.
bool chan_fast_supported(double bandwidth, double fs)
{
    std::complex<double> pz[4];
    double k;
    if (channel_filter_poles(bandwidth, fs, pz, &k)) return false;
    double r = 0.0;
    for (int i = 0; i < 4; ++i) r = std::max(r, std::abs(pz[i]));
    return std::pow(r, (double)CF_P) <= 1e-7;
}

int launch_channel_fast(tb200_handle h, const void *x, int dtype, int n_chunks, long n, long in_stride,
                        const double *freq_off, double fs, double bandwidth, double *filtered, long out_stride,
                        double *phase_partial, double *phase_tail, int sps, int step, cudaStream_t st)
{
    // constants depend on (bandwidth, fs) only: cached (see launch_decimate_fast)
    struct Cached { double bw, fs; ChanFastParams K; };
    static std::mutex mu;
    static std::vector<Cached> cache;
    ChanFastParams K;
    bool hit = false;
    {
        std::lock_guard<std::mutex> lk(mu);
        for (const Cached &c : cache) if (c.bw == bandwidth && c.fs == fs) { K = c.K; hit = true; break; }
    }
    if (!hit) {
        int rc0 = make_chan_params(bandwidth, fs, K);
        if (rc0) return rc0;
        std::lock_guard<std::mutex> lk(mu);
        if (cache.size() >= 16) cache.erase(cache.begin());
        cache.push_back(Cached{bandwidth, fs, K});
    }
    if (sps > TB200_MAX_PHASES || sps <= 1) phase_partial = phase_tail = nullptr;
    if (dtype == TB200_DTYPE_C64)
        return launch_t<float>(h, x, n_chunks, n, in_stride, freq_off, fs, K, filtered, out_stride, phase_partial, phase_tail, sps, step, st);
    return launch_t<double>(h, x, n_chunks, n, in_stride, freq_off, fs, K, filtered, out_stride, phase_partial, phase_tail, sps, step, st);
}

}  // namespace tb200
