// dsp_fast.cu — time-parallel ("FAST" mode) zero-phase decimator for scipy.signal.decimate(x, q)
// as called by SignalProcessor.process (/root/reference/tetraear/signal/processor.py:254).
//
// Same filter as the reference (the float32-rounded cheby1 biquads, odd extension, steady-state
// initial conditions, forward + backward pass, keep every q-th sample) but evaluated in a form
// that has no long recurrence:
//
//   H(z) H(1/z) over the padded chunk  ==  sum over the 4 conjugate pole pairs p_i of
//        2 Re{ alpha_i ( F_i[n] + p_i B_i[n+1] ) }  -  (right-edge transient),
//   F_i[n] = sum_{k>=0} p_i^k x[n-k]   (causal one-pole),  B_i[n] = sum_{k>=0} p_i^k x[n+k].
//
// Only every q-th output is needed, so per kept output k the work is two q-tap complex FIRs
//   U_i[k] = sum_{j<q} p_i^j x[n_k - j],   V_i[k] = sum_{j<q} p_i^j x[n_k + 1 + j]
// plus the decimated-rate recurrences F[k] = lambda F[k-1] + U[k], Bq[k] = V[k] + lambda Bq[k+1]
// with lambda = p^q (|lambda| <= 0.834 for q = 10).  A pass of 96 outputs forgets its past to
// 0.834^96 = 3e-8, so each warp works on one 96-output pass from a zero state and receives the
// end state of its neighbour warp as carry; warp-level Kogge-Stone scans (shuffles) resolve
// the recurrence inside the pass.  The residues alpha_i are ~1e-2, so float32 is ample: the
// result is within 2e-7 of the exact-arithmetic filter (the reference's own float32 recursion
// is 3e-5 away from it; DESIGN.md "FAST mode accuracy").
//
// Data layout: a CTA stages its tile of the padded chunk in shared memory as separate I and Q
// planes, one row of q samples per output with an odd row pitch (q+1 for even q), so that the
// 32 lanes of a warp — each owning R=3 consecutive rows — read conflict-free.
#include <atomic>
#include <complex>
#include <mutex>
#include <vector>
#include "common.h"
#include "kernels.h"
#include "packed.h"
#include "decim_common.h"

namespace tb200 {

// one direction (causal: row0.., taps reversed; anti-causal: next row, taps ascending) for two of the
// four modes; accumulates Re(alpha * .) into y (I, Q packed) and leaves the pass-end state in `carry`.
template <int Q, int PITCH, bool FWD, int NM>
__device__ __forceinline__ void fd_half(const float2 *__restrict__ tile, int row0, int lane, int m0,
                                        const FastDecimParams &K, float2 y[FD_R], cp *carry)
{
    cp U[FD_R][NM];
#pragma unroll
    for (int r = 0; r < FD_R; ++r) {
#pragma unroll
        for (int mm = 0; mm < NM; ++mm) U[r][mm].re = U[r][mm].im = make_float2(0.f, 0.f);
        const float2 *row = tile + (row0 + r + (FWD ? 0 : 1)) * PITCH;
#pragma unroll
        for (int c = 0; c < Q; ++c) {
            const float2 v = row[c];
#pragma unroll
            for (int mm = 0; mm < NM; ++mm) {
                const int tp = FWD ? Q - 1 - c : c;
                U[r][mm].re = ffma2(splat(K.tap[tp][m0 + mm][0]), v, U[r][mm].re);
                U[r][mm].im = ffma2(splat(K.tap[tp][m0 + mm][1]), v, U[r][mm].im);
            }
        }
    }
#pragma unroll
    for (int mm = 0; mm < NM; ++mm) {
        const int m = m0 + mm;
        const cf lam = ld2(K.lam[m]);
        // zero-state recurrence inside the lane, then an inclusive scan of the lane aggregates
        if (FWD) {
#pragma unroll
            for (int r = 1; r < FD_R; ++r) U[r][mm] = cpfma(lam, U[r - 1][mm], U[r][mm]);
        } else {
#pragma unroll
            for (int r = FD_R - 2; r >= 0; --r) U[r][mm] = cpfma(lam, U[r + 1][mm], U[r][mm]);
        }
        cp a = U[FWD ? FD_R - 1 : 0][mm];
#pragma unroll
        for (int s = 0; s < 5; ++s) {
            if (s >= fd_nstage(m)) continue;             // multiplier (lambda^R)^(2^s) below 1e-9: nothing to add in float32 (compile time)
            cp o;
            o.re = FWD ? shfl2_up(a.re, 1 << s) : shfl2_down(a.re, 1 << s);
            o.im = FWD ? shfl2_up(a.im, 1 << s) : shfl2_down(a.im, 1 << s);
            const bool take = FWD ? (lane >= (1 << s)) : (lane + (1 << s) < 32);
            cf g = ld2(K.lampow[s][m]);
            if (!take) g.x = g.y = 0.f;                  // select on the multiplier, not on the 4-register result
            a = cpfma(g, o, a);
        }
        cp cin;
        cin.re = FWD ? shfl2_up(a.re, 1) : shfl2_down(a.re, 1);
        cin.im = FWD ? shfl2_up(a.im, 1) : shfl2_down(a.im, 1);
        if (lane == (FWD ? 0 : 31)) cin.re = cin.im = make_float2(0.f, 0.f);
        if (lane == (FWD ? 31 : 0)) carry[m] = a;
        const cf a0 = FWD ? ld2(K.a0[m]) : ld2(K.ap[m]);
#pragma unroll
        for (int r = 0; r < FD_R; ++r) {
            const cf af = FWD ? ld2(K.aF[r][m]) : ld2(K.aB[r][m]);
            y[r] = re_acc(af, cin, re_acc(a0, U[r][mm], y[r]));
        }
    }
}


// kappa_i (packed over I, Q) of one chunk from the staged tile; s_end = tile index of ext[M-1].  One warp.
__device__ __forceinline__ void fd_kappa(const float2 *__restrict__ tile, int s_end, int lane, const FastDecimParams &K, cp *kappa)
{
    cp S[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) S[j].re = S[j].im = make_float2(0.f, 0.f);
    const int s0 = s_end - 32 * FD_EDGE_S + 1 + lane * FD_EDGE_S;
    for (int i = 0; i < FD_EDGE_S; ++i) {
        const int si = s0 + i;
        const float2 e = si >= 0 ? tile[si] : tile[0];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            cp nx = cpmul(ld2(K.p[j]), S[j]);
            nx.re.x += e.x; nx.re.y += e.y;
            S[j] = nx;
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int st = 0; st < 5; ++st) {
            cp o;
            o.re = shfl2_up(S[j].re, 1 << st); o.im = shfl2_up(S[j].im, 1 << st);
            if (lane >= (1 << st)) S[j] = cpfma(ld2(K.pS[st][j]), o, S[j]);
        }
    }
    // lane 31 holds F_j[M-1]; kappa_i = ka_i x_last + sum_j kM_ij F_j + kMc_ij conj(F_j)
    if (lane == 31) {
        const float2 xl = tile[s_end];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            cp k;
            k.re = make_float2(K.ka[i][0] * xl.x, K.ka[i][0] * xl.y);
            k.im = make_float2(K.ka[i][1] * xl.x, K.ka[i][1] * xl.y);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                k = cpfma(ld2(K.kM[i][j]), S[j], k);
                cp cj; cj.re = S[j].re; cj.im = make_float2(-S[j].im.x, -S[j].im.y);
                k = cpfma(ld2(K.kMc[i][j]), cj, k);
            }
            kappa[i] = k;
        }
    }
}

template <int Q>
__global__ void __launch_bounds__(FD_W * 32, 3)
decim_fast_kernel(const float2 *__restrict__ x, long n, long chunk_stride, int tiles_per_chunk, long n_dec,
                  FastDecimParams K, float2 *__restrict__ dec_out)
{
    // The tile is the CONTIGUOUS run of ROWS*Q input samples (interleaved I,Q) behind the CTA's outputs,
    // laid out densely (row pitch Q) so that one TMA bulk copy stages it; lanes are R rows = R*Q float2
    // apart (30 for q = 10: a 2-way bank conflict on the 8-byte loads, cheaper than a padded layout that
    // TMA cannot write).  One extra leading sample keeps the global source 16-byte aligned.
    constexpr int PITCH = Q;
    constexpr int ROWS = FD_NP * FD_P + 1;
    constexpr int TILE = ROWS * Q;
    constexpr unsigned TMA_BYTES = (unsigned)(((TILE + 1) * 8 + 15) & ~15);
    extern __shared__ __align__(128) float2 fd_raw[];
    __shared__ __align__(8) uint64_t tile_bar;
    float2 *fd_tile = fd_raw + 1;
    __shared__ cp carryF[FD_NP][4];
    __shared__ cp carryB[FD_NP][4];
    __shared__ cf lanepow[32][4];
    __shared__ cp kappa[4];
    if (threadIdx.x < 128) lanepow[threadIdx.x >> 2][threadIdx.x & 3] = ld2(K.lanepow[threadIdx.x >> 2][threadIdx.x & 3]);

    const int chunk = blockIdx.x / tiles_per_chunk;
    const int tile = blockIdx.x - chunk * tiles_per_chunk;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float2 *xc = x + (long)chunk * chunk_stride;
    // first output of pass 0 (the left halo) and the padded-signal position of tile row 0, col 0.
    // Tiles are aligned to the END of the chunk (the first one may be partial): the last FD_NS*FD_P outputs and
    // the whole right extension then live in ONE tile, which also computes and applies the right-edge term.
    const bool last_tile = tile == tiles_per_chunk - 1;
    const long k_tile0 = n_dec - (long)(tiles_per_chunk - tile) * FD_NS * FD_P - FD_P;
    const long t0 = 27 + (long)Q * k_tile0 - (Q - 1);

    // ---- stage the tile with TMA bulk copies issued by thread 0 (everybody waits on the mbarrier).  Interior
    // tiles: ONE copy.  Tiles touching a chunk end: the in-range, 16-byte aligned middle part by TMA, the odd
    // extension / constant continuation sample by sample.  fd_raw[j] is padded-signal sample t0 - 1 + j, i.e.
    // chunk sample g0 + j with g0 = t0 - 28 (one sample early: 16-byte aligned source for even t0).
    const long g0 = t0 - 28;
    const bool aligned = ((reinterpret_cast<uintptr_t>(xc) & 15) == 0) && ((g0 & 1) == 0);
    const bool interior = aligned && g0 >= 0 && (g0 + (long)(TMA_BYTES / 8) <= n);
    if (interior) {
        if (threadIdx.x == 0) {
            mbar_init(&tile_bar, 1);
            tma_load_1d_stream(fd_raw, xc + g0, TMA_BYTES, &tile_bar);
        }
        __syncthreads();                                       // barrier initialised before anybody polls it
        mbar_wait(&tile_bar, 0);
    } else {
        long ja = g0 < 0 ? -g0 : 0, jb = (n - g0 < (long)(TILE + 1)) ? n - g0 : (long)(TILE + 1);
        ja = (ja + 1) & ~1L;
        jb = ja + ((jb - ja) & ~1L);
        const bool bulk = aligned && jb - ja >= 64;
        if (!bulk) ja = jb = 0;
        if (bulk && threadIdx.x == 0) {
            mbar_init(&tile_bar, 1);
            tma_load_1d_stream(fd_raw + ja, xc + g0 + ja, (unsigned)((jb - ja) * 8), &tile_bar);
        }
        for (int s = threadIdx.x; s < TILE + 1; s += blockDim.x)
            if (s < ja || s >= jb) fd_raw[s] = padded_sample(xc, n, t0 - 1 + s);
        __syncthreads();
        if (bulk) mbar_wait(&tile_bar, 0);
    }

    // warp 0 runs the two halo half-passes (causal over pass 0, anti-causal over pass FD_NP-1), warps 1..FD_NS one
    // storing pass each (both directions): every warp does the same amount of work.
    const bool store = w > 0;
    const int passB = store ? w : FD_NP - 1;
    const int row0 = w * FD_P + lane * FD_R, row0B = passB * FD_P + lane * FD_R;
    float2 y[FD_R];
#pragma unroll
    for (int r = 0; r < FD_R; ++r) y[r] = make_float2(0.f, 0.f);
    fd_half<Q, PITCH, true, FD_NM>(fd_tile, row0, lane, 0, K, y, carryF[w]);
    if (FD_NM == 2) fd_half<Q, PITCH, true, FD_NM>(fd_tile, row0, lane, 2, K, y, carryF[w]);
    fd_half<Q, PITCH, false, FD_NM>(fd_tile, row0B, lane, 0, K, y, carryB[passB]);
    if (FD_NM == 2) fd_half<Q, PITCH, false, FD_NM>(fd_tile, row0B, lane, 2, K, y, carryB[passB]);
    // Right-edge transient (see the header): warp 0 of the chunk's last tile derives kappa_i from
    // F_i[M-1] = sum_k p_i^k ext[M-1-k] over the tile in smem.
    if (last_tile && w == 0) fd_kappa(fd_tile, (int)((n + 53) - t0), lane, K, kappa);
    __syncthreads();
    if (!store) return;
    // ---- carries from the neighbour passes: lambda^(lane R) cF on the left, lambda^((31-lane) R) cB on the right
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const cp gF = cpmul(lanepow[lane][m], carryF[w - 1][m]);
        const cp gB = cpmul(lanepow[31 - lane][m], carryB[w + 1][m]);
#pragma unroll
        for (int r = 0; r < FD_R; ++r) y[r] = re_acc(ld2(K.aB[r][m]), gB, re_acc(ld2(K.aF[r][m]), gF, y[r]));
    }
    float2 *out = dec_out + (long)chunk * n_dec;
    const long k0 = k_tile0 + row0;
    if (last_tile && n_dec - 1 - k0 < FD_NFIX + FD_R) {
        // y[k] -= Re sum_i kappa_i pc_i lambda_i^mm,  mm = n_dec-1-k  (only the last FD_NFIX outputs matter)
#pragma unroll
        for (int r = 0; r < FD_R; ++r) {
            const long mm = n_dec - 1 - (k0 + r);
            if (mm < 0 || mm >= FD_NFIX) continue;
            const int a = (int)(mm / FD_R), b = (int)(mm - (long)a * FD_R);
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                cf g = ld2(K.pc[m]);
#pragma unroll
                for (int sft = 0; sft < 6; ++sft) if (a & (1 << sft)) g = cmulf(g, ld2(K.lampow[sft][m]));
                const cf lam = ld2(K.lam[m]);
                for (int i = 0; i < b; ++i) g = cmulf(g, lam);
                y[r] = re_acc(cf{-g.x, -g.y}, kappa[m], y[r]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < FD_R; ++r) {
        long k = k0 + r;
        if (k >= 0 && k < n_dec) out[k] = make_float2(2.0f * y[r].x, 2.0f * y[r].y);
    }
}

// ---- host: modal design from the float32-rounded biquads (double precision) -----------------
typedef std::complex<double> cd;

static void put(float dst[2], cd v) { dst[0] = (float)v.real(); dst[1] = (float)v.imag(); }

int make_fast_params(const DecimDesign &d, int q, long n, long n_dec, FastDecimParams &K)
{
    cd p8[8], r8[8];
    const double g = (double)d.sos32[0];
    for (int s = 0; s < 4; ++s) {
        double a1 = (double)d.sos32[6 * s + 4], a2 = (double)d.sos32[6 * s + 5];
        double disc = a1 * a1 - 4 * a2;
        if (!(disc < 0)) return fail(TB200_E_ARG, "fast decimator: real pole pair in section %d", s);
        p8[s] = cd(-a1 / 2, std::sqrt(-disc) / 2);
        p8[s + 4] = std::conj(p8[s]);
    }
    cd prodp(1.0, 0.0);
    for (int i = 0; i < 8; ++i) prodp *= p8[i];
    const double dd = g / prodp.real();
    for (int i = 0; i < 8; ++i) {
        cd num = g * std::pow(1.0 + 1.0 / p8[i], 8), den(1.0, 0.0);
        for (int j = 0; j < 8; ++j) if (j != i) den *= (1.0 - p8[j] / p8[i]);
        r8[i] = num / den;
    }
    cd alpha[4], lam[4];
    for (int j = 0; j < 4; ++j) {
        cd acc(dd, 0.0);
        for (int i = 0; i < 8; ++i) acc += r8[i] / (1.0 - p8[i] * p8[j]);
        alpha[j] = r8[j] * acc;
        lam[j] = std::pow(p8[j], q);
    }
    memset(&K, 0, sizeof(K));
    for (int j = 0; j < 4; ++j) {
        cd pw(1.0, 0.0);
        for (int t = 0; t < q; ++t) { put(K.tap[t][j], pw); pw *= p8[j]; }
        put(K.lam[j], lam[j]);
        cd lr = std::pow(lam[j], FD_R);
        K.nstage[j] = 0;
        for (int s = 0; s < 6; ++s) {
            put(K.lampow[s][j], lr);
            if (s < 5 && std::abs(lr) >= 1e-9) K.nstage[j] = s + 1;
            lr *= lr;
        }
        if (K.nstage[j] > fd_nstage(j))
            return fail(TB200_E_ARG, "fast decimator: mode %d of q=%d needs %d scan stages, kernel has %d", j, q, K.nstage[j], fd_nstage(j));
        put(K.a0[j], alpha[j]);
        put(K.ap[j], alpha[j] * p8[j]);
        for (int r = 0; r < FD_R; ++r) {
            put(K.aF[r][j], alpha[j] * std::pow(lam[j], r + 1));
            put(K.aB[r][j], alpha[j] * p8[j] * std::pow(lam[j], FD_R - r));
        }
        put(K.p[j], p8[j]);
        put(K.inv1mp[j], 1.0 / (1.0 - p8[j]));
        for (int l = 0; l < 32; ++l) put(K.lanepow[l][j], std::pow(lam[j], FD_R * l));
        cd ps = std::pow(p8[j], FD_EDGE_S);
        for (int st = 0; st < 5; ++st) { put(K.pS[st][j], ps); ps *= ps; }
    }
    // kappa_i = r_i E_i,  E_i = (G - wlast)/(1-p_i) + sum_j r_j delta_j p_j / (1 - p_i p_j),
    // G = xc H(1), wlast = d xc + sum_j r_j F_j, delta_j = F_j - xc/(1-p_j)   (j over all 8 poles)
    cd H1(dd, 0.0);
    for (int j = 0; j < 8; ++j) H1 += r8[j] / (1.0 - p8[j]);
    const long M = n + 54;
    const long n_last = 27 + (long)q * (n_dec - 1);
    for (int i = 0; i < 4; ++i) {
        cd ka = (H1 - dd) / (1.0 - p8[i]);
        cd kM[8];
        for (int j = 0; j < 8; ++j) {
            ka -= r8[j] * p8[j] / ((1.0 - p8[i] * p8[j]) * (1.0 - p8[j]));
            kM[j] = -r8[j] / (1.0 - p8[i]) + r8[j] * p8[j] / (1.0 - p8[i] * p8[j]);
        }
        put(K.ka[i], r8[i] * ka);
        for (int j = 0; j < 4; ++j) { put(K.kM[i][j], r8[i] * kM[j]); put(K.kMc[i][j], r8[i] * kM[j + 4]); }
        put(K.pc[i], std::pow(p8[i], (double)(M - n_last)));
    }
    return TB200_OK;
}

int fast_decimator_supported(int q) { return q == 10 || q == 7 || q == 8; }

template <int Q>
static int launch_q(tb200_handle h, const void *x, int n_chunks, long n, long chunk_stride, long n_dec,
                    const FastDecimParams &K, void *dec_out, cudaStream_t st)
{
    constexpr int ROWS = FD_NP * FD_P + 1;
    const size_t smem = ((size_t)(ROWS * Q + 1) * sizeof(float2) + 15 + 16) & ~(size_t)15;
    static std::atomic<bool> attr_set[64];       // per device: the attribute belongs to the device's copy of the kernel (setting it twice is harmless)
    const int dv = h->cfg.device & 63;
    if (!attr_set[dv]) {
        cudaError_t e = cudaFuncSetAttribute(decim_fast_kernel<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(TB200_E_CUDA, "cudaFuncSetAttribute(decim_fast): %s", cudaGetErrorString(e));
        attr_set[dv] = true;
    }
    const int per_tile = FD_NS * FD_P;
    const int tiles = (int)((n_dec + per_tile - 1) / per_tile);
    decim_fast_kernel<Q><<<n_chunks * tiles, FD_W * 32, smem, st>>>((const float2 *)x, n, chunk_stride, tiles, n_dec, K,
                                                                  (float2 *)dec_out);
    TB_LAUNCH_CHECK(h);
    return TB200_OK;
}

int launch_decimate_fast(tb200_handle h, const void *x, int n_chunks, long n, long chunk_stride, int q,
                         const DecimDesign &d, void *dec_out, long n_dec, cudaStream_t st)
{
    if (!fast_decimator_supported(q)) return fail(TB200_E_ARG, "fast decimator: q=%d not instantiated", q);
    // the modal constants depend on (design, q, n, n_dec) only: keep the last few sets instead of re-deriving them
    // (dozens of complex pow) on every launch — the host must stay ahead of a ~50 us kernel
    struct Cached { int q; long n, n_dec; float sos[24]; FastDecimParams K; };
    static std::mutex mu;
    static std::vector<Cached> cache;
    FastDecimParams K;
    bool hit = false;
    {
        std::lock_guard<std::mutex> lk(mu);
        for (const Cached &c : cache)
            if (c.q == q && c.n == n && c.n_dec == n_dec && !memcmp(c.sos, d.sos32, sizeof(c.sos))) { K = c.K; hit = true; break; }
    }
    if (!hit) {
        int rc0 = make_fast_params(d, q, n, n_dec, K);
        if (rc0) return rc0;
        std::lock_guard<std::mutex> lk(mu);
        if (cache.size() >= 16) cache.erase(cache.begin());
        Cached c; c.q = q; c.n = n; c.n_dec = n_dec; memcpy(c.sos, d.sos32, sizeof(c.sos)); c.K = K;
        cache.push_back(c);
    }
    int rc = TB200_OK;
    if (q == 10) rc = launch_q<10>(h, x, n_chunks, n, chunk_stride, n_dec, K, dec_out, st);
    else if (q == 7) rc = launch_q<7>(h, x, n_chunks, n, chunk_stride, n_dec, K, dec_out, st);
    else rc = launch_q<8>(h, x, n_chunks, n, chunk_stride, n_dec, K, dec_out, st);
    return rc;
}

}  // namespace tb200
