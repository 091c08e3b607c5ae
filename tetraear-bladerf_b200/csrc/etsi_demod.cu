// etsi_demod.cu — `etsi` mode physical layer per 25 kHz channel: RRC (alpha 0.35) matched filter,
// closed-loop symbol timing at 18 ksym/s (Gardner detector + PI loop, fractional-delay polyphase
// matched-filter bank), pi/4-DQPSK differential detection with SOFT bit output, and the
// normal-burst training-sequence search on the soft bits.
//
// None of this exists in the reference (it uses a Butterworth low-pass, an
// open-loop 13-phase power pick at 18 461.5 sym/s, a hard slicer with mis-centred decision
// regions and a hard-bit Hamming scan) — parity unpinned; validated end to end by
// modulator -> demod -> TS search -> channel decoder round trips (tests/test_etsi_phy.py).
//
// Eight lanes per stream (two in the time-parallel form): the feedback state (timing phase, loop integrator, power estimate) is replicated in the
// group's registers; each lane owns 4 of the 32 taps of the interpolating matched filter (reduced with 3 shuffle
// stages) for the on-time and the half-symbol sample of every symbol.
#include <cmath>
#include "common.h"
#include "kernels.h"
#include "packed.h"

namespace tb200 {

constexpr int ED_TAPS = 32;
constexpr int ED_PH = 32;

// Mapping: EIGHT lanes per stream, four streams per warp.  The loop is a chain of dependent steps per symbol
// (timing phase -> two matched-filter outputs -> Gardner error -> loop filter -> next phase), so what counts is the
// latency of that chain, not arithmetic: each lane owns 4 of the 32 taps (4 multiply-adds per output in registers,
// then a 3-stage shuffle reduction instead of 5), the stream's samples sit in a shared-memory ring the group refills
// 64 samples at a time well ahead of their use, and the tap table is read as one float4 per output.
constexpr int ED_RING = 128;              // ring capacity per stream (samples): window 33 + two refills + a symbol
constexpr int ED_FILL = 32;               // refill granularity

// Time-parallel form (SEG = true).  The loop of a stream is sequential, but it forgets: after ~100 symbols its state
// (timing phase, integrator, power estimate) no longer depends on where it started.  A stream is therefore cut at the
// nominal symbol instants T_j = t_first + j * seg_syms * sps into segments that run concurrently, each in its own
// 8-lane group: segment j > 0 starts ED_WARM symbols before T_j from the nominal state, keeps quiet until its timing
// instant passes T_j, then writes its symbols (soft pair, instant) to a staging row, and runs 2 symbols past T_(j+1).
// etsi_stitch_kernel joins the rows: segment j keeps the symbols that lie more than half a symbol before the first
// symbol of segment j+1, so every symbol is taken exactly once even if two loops disagree about the instant by a
// fraction of a symbol.  Work grows by ED_WARM / seg_syms; parallelism by the number of segments.
constexpr int ED_WARM = 160;              // warm-up symbols of a segment (loop time constant ~15 symbols)
constexpr int ED_SEG_MIN = 640;           // nominal symbols per segment (at least)
constexpr int ED_SEG_MAX_N = 128;         // segments per stream (at most)
constexpr int ED_SLACK = 16;              // staging slots beyond seg_syms

struct EtsiSegMeta { int count; };

// L lanes per group: the sequential form wants the shortest dependent chain per symbol (L = 8: 4 taps per lane, three
// shuffle stages); the time-parallel form has groups to spare and wants the fewest instructions per symbol (L = 2: the
// loop arithmetic, replicated in every lane of a group, is the larger part of the work at L = 8).
template <bool SEG, int L>
__global__ void __launch_bounds__(128)
etsi_demod_kernel(const float2 *__restrict__ iq, int n_streams, long n, long stride, float sps, float kp,
                  float ki, const float *__restrict__ taps_g, float2 *__restrict__ sym_out,
                  float *__restrict__ soft_out, long soft_stride, int32_t *__restrict__ n_soft,
                  int n_seg, int seg_syms, float4 *__restrict__ stage, float2 *__restrict__ sym_stage, int32_t *__restrict__ seg_count)
{
    __shared__ __align__(16) float taps[ED_PH * ED_TAPS];
    constexpr int TPL = ED_TAPS / L, GPC = 128 / L;            // taps per lane, groups per CTA
    extern __shared__ float2 ed_ring[];                        // [GPC][ED_RING]
    // lane q owns taps q, q + L, q + 2L, ...: the lanes of a group read consecutive samples (one 64-byte run per tap
    // index instead of 32-byte strides); the table is stored lane-major so that a lane's taps are one float4 run
    for (int i = threadIdx.x; i < ED_PH * ED_TAPS; i += blockDim.x) {
        const int ph = i / ED_TAPS, t = i % ED_TAPS;
        taps[ph * ED_TAPS + (t % L) * TPL + t / L] = taps_g[i];
    }
    __syncthreads();
    const int q = threadIdx.x % L;
    const int grp = threadIdx.x / L;                           // group slot in the CTA
    const int gid = blockIdx.x * GPC + grp;
    const int s = SEG ? gid / n_seg : gid;
    const int j = SEG ? gid - s * n_seg : 0;
    const bool live = s < n_streams;
    const float2 *x = iq + (long)(live ? s : 0) * stride;
    float2 *rb = ed_ring + grp * ED_RING;
    const float t_first = 16.0f + sps;                         // first symbol instant with a full filter window
    const int ni = (int)n;                                     // n < 2^30 (launch check): 32-bit index arithmetic in the loop
    const float nf_end = (float)n - 17.0f;
    const float half = 0.5f * sps;
    float tau = t_first, emit_from = -1.0f, run_until = nf_end;
    int slots = (int)(soft_stride / 2);
    if (SEG) {
        const float Tj = t_first + (float)j * (float)seg_syms * sps;
        if (j > 0) { tau = Tj - (float)ED_WARM * sps; emit_from = Tj; }
        if (j < n_seg - 1) run_until = fminf(nf_end, Tj + ((float)seg_syms + 2.0f) * sps);
        slots = seg_syms + ED_SLACK;
    }
    float integ = 0.f, pavg = 0.f;
    float2 prev = make_float2(0.f, 0.f);
    int k = 0, it = 0;                                         // symbols written, loop iterations
    float2 *so = nullptr;
    float *sb = soft_out + (long)(live ? s : 0) * soft_stride;
    float4 *stg = nullptr;
    if (SEG) {
        const long row = ((long)(live ? s : 0) * n_seg + j) * slots;
        stg = stage + row;
        if (sym_stage && live) so = sym_stage + row;
    } else if (sym_out && live) so = sym_out + (long)s * slots;
    // the ring holds samples [filled - ED_RING, filled); a segment starts it just below its first filter window
    int filled = 0;
    if (SEG && j > 0) { const int lo = (int)floorf(tau - half) - 16; filled = lo > 0 ? (lo & ~(ED_FILL - 1)) : 0; }
    float2 pre[ED_FILL / L];                                    // this lane's share of the next refill
#pragma unroll
    for (int jj = 0; jj < ED_FILL / L; ++jj) {
        const int idx = filled + q * (ED_FILL / L) + jj;
        pre[jj] = (live && idx < ni) ? x[idx] : make_float2(0.f, 0.f);
    }
    bool active = live && (tau < run_until) && (k < slots);
    while (__any_sync(0xffffffffu, active)) {
        const float tmid = tau - half;
        const float fl_on = floorf(tau), fl_mid = floorf(tmid);
        const int i_on = (int)fl_on, i_mid = (int)fl_mid;
        // keep [i_mid - 15, i_on + 16] inside the ring: refill ahead, ED_FILL samples at a time.  The samples of the next
        // refill are already in registers (loaded when the previous one was stored): no load latency in the loop.
#pragma unroll 1
        while (active && filled < i_on + 17 + ED_FILL) {
            const int g = filled + q * (ED_FILL / L);
#pragma unroll
            for (int jj = 0; jj < ED_FILL / L; ++jj) rb[(g + jj) & (ED_RING - 1)] = pre[jj];
            filled += ED_FILL;
#pragma unroll
            for (int jj = 0; jj < ED_FILL / L; ++jj) {
                const int idx = g + ED_FILL + jj;
                pre[jj] = (idx < ni) ? x[idx] : make_float2(0.f, 0.f);
            }
        }
        __syncwarp();
        int ph_on = (int)((tau - fl_on) * ED_PH), ph_mid = (int)((tmid - fl_mid) * ED_PH);
        ph_on = min(max(ph_on, 0), ED_PH - 1);
        ph_mid = min(max(ph_mid, 0), ED_PH - 1);
        const int b_on = i_on - 15 + q, b_mid = i_mid - 15 + q;
        float2 on = make_float2(0.f, 0.f), mid = make_float2(0.f, 0.f);
#pragma unroll
        for (int t4 = 0; t4 < TPL / 4; ++t4) {
            const float4 g_on = *reinterpret_cast<const float4 *>(&taps[ph_on * ED_TAPS + TPL * q + 4 * t4]);
            const float4 g_mid = *reinterpret_cast<const float4 *>(&taps[ph_mid * ED_TAPS + TPL * q + 4 * t4]);
            const float gw[4] = {g_on.x, g_on.y, g_on.z, g_on.w}, hw[4] = {g_mid.x, g_mid.y, g_mid.z, g_mid.w};
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int a = b_on + L * (4 * t4 + jj), c = b_mid + L * (4 * t4 + jj);
                // a, c >= the ring's first sample: tau only moves forward from t_first (a >= 2) or from the segment start
                const float2 va = rb[a & (ED_RING - 1)];
                const float2 vc = rb[c & (ED_RING - 1)];
                on.x = fmaf(va.x, gw[jj], on.x); on.y = fmaf(va.y, gw[jj], on.y);
                mid.x = fmaf(vc.x, hw[jj], mid.x); mid.y = fmaf(vc.y, hw[jj], mid.y);
            }
        }
#pragma unroll
        for (int o = 1; o < L; o <<= 1) {
            on.x += __shfl_xor_sync(0xffffffffu, on.x, o); on.y += __shfl_xor_sync(0xffffffffu, on.y, o);
            mid.x += __shfl_xor_sync(0xffffffffu, mid.x, o); mid.y += __shfl_xor_sync(0xffffffffu, mid.y, o);
        }
        if (active) {
            const float p = on.x * on.x + on.y * on.y;
            pavg = (it == 0) ? p : 0.98f * pavg + 0.02f * p;
            const float inv = __fdividef(1.0f, pavg + 1e-20f);
            // Gardner: e = Re{(y[k] - y[k-1]) conj(y[k-1/2])}; e > 0 <=> sampling late
            float e = ((on.x - prev.x) * mid.x + (on.y - prev.y) * mid.y) * inv;
            e = fminf(fmaxf(e, -1.0f), 1.0f);
            // differential detection: d = y[k] conj(y[k-1]); b1 = [Im d < 0], b2 = [Re d < 0]; soft = -(.)/P
            const float dre = on.x * prev.x + on.y * prev.y, dim = on.y * prev.x - on.x * prev.y;
            if (!SEG || tau >= emit_from) {
                if (q == 0) {
                    if (so) so[k] = on;
                    if (SEG) stg[k] = make_float4(-dim * inv, -dre * inv, tau, 0.f);
                    else if (k > 0) *reinterpret_cast<float2 *>(sb + 2 * (k - 1)) = make_float2(-dim * inv, -dre * inv);
                }
                ++k;
            }
            prev = on;
            if (it > 0) { integ += ki * e; tau += sps - kp * e - integ; }
            else tau += sps;
            ++it;
            active = (tau < run_until) && (k < slots);
        }
        __syncwarp();                                           // ring reads of this symbol precede the next refill
    }
    if (live && q == 0) {
        if (SEG) seg_count[(long)s * n_seg + j] = k;
        else n_soft[s] = (int32_t)(k > 0 ? 2 * (k - 1) : 0);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Two-stage time-parallel form (default of tb200_etsi_demod_segmented).  The eight-lane kernel above evaluates the
// 32-tap matched filter INSIDE the sequential loop, twice per symbol at the loop's current timing phase: ~210 warp
// instructions per symbol and the kernel is issue-bound (757 M instructions per 800 x 10 000 samples).  But the matched
// filter does not depend on the loop: stage 1 runs it ONCE per input sample for all streams in parallel
// (etsi_mf_kernel, y[i] = sum_t h0[t] x[i - 15 + t]: the phase-0 row of the same tap table), and the loop of stage 2
// only INTERPOLATES y at its timing instants with an 8-tap fractional-delay kernel (Kaiser-windowed sinc, beta = 6,
// 32 phases: 1e-3 rms of full scale on a raised-cosine signal at 2 ... 4 samples per symbol, below the 1/32-sample
// phase quantisation both forms share) — 16 multiply-adds per sample instead of 64, by ONE THREAD per segment, so a warp
// advances 32 segments per instruction and the segments can be short (160 symbols + 160 of warm-up).  Same loop
// equations, same staging rows, same stitch kernel.
constexpr int EI_TAPS = 8;
constexpr int ETS_SEG_MIN = 160;

__global__ void __launch_bounds__(256)
etsi_mf_kernel(const float2 *__restrict__ iq, long n, long stride, const float *__restrict__ taps_g, float2 *__restrict__ y, long ypitch)
{
    // 1024 outputs per CTA, FOUR CONSECUTIVE outputs per thread from a 35-sample register window: 8.75 shared loads per
    // output instead of 32 (the first form, one load per tap and output, was shared-memory bound: 311 us per 32 M outputs).
    // The tile is stored [e & 3][e >> 2] so that lane L's element 4 L + c is word L + (c >> 2) of row c & 3: conflict-free.
    constexpr int OUT = 1024, COLS = OUT / 4 + 9;
    __shared__ float h0[ED_TAPS];
    __shared__ float2 xs[4][COLS];
    const int s = blockIdx.y;
    const long i0 = (long)blockIdx.x * OUT;
    const float2 *x = iq + (long)s * stride;
    if (threadIdx.x < ED_TAPS) h0[threadIdx.x] = taps_g[threadIdx.x];          // phase 0 row of the table
    for (int e = threadIdx.x; e < OUT + ED_TAPS + 3; e += 256) {
        const long t = i0 - 15 + e;
        xs[e & 3][e >> 2] = (t >= 0 && t < n) ? x[t] : make_float2(0.f, 0.f);
    }
    __syncthreads();
    float2 w[ED_TAPS + 3];
#pragma unroll
    for (int c = 0; c < ED_TAPS + 3; ++c) w[c] = xs[c & 3][threadIdx.x + (c >> 2)];
    float2 acc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        acc[u] = make_float2(0.f, 0.f);
#pragma unroll
        for (int t = 0; t < ED_TAPS; ++t) acc[u] = ffma2(splat(h0[t]), w[u + t], acc[u]);
    }
    const long o = i0 + 4 * threadIdx.x;
    float2 *ys = y + (long)s * ypitch + o;                                       // ypitch even, o multiple of 4: 16-byte aligned
    if (o + 3 < n) {
        *reinterpret_cast<float4 *>(ys) = make_float4(acc[0].x, acc[0].y, acc[1].x, acc[1].y);
        *reinterpret_cast<float4 *>(ys + 2) = make_float4(acc[2].x, acc[2].y, acc[3].x, acc[3].y);
    } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) if (o + u < n) ys[u] = acc[u];
    }
}

// both interpolations of a symbol — Y(tau - half) and Y(tau) — from ONE aligned 12-sample window: six 16-byte loads per
// symbol instead of sixteen 8-byte ones.  (The threads of a warp walk 32 different segments, so every load instruction
// costs 32 L1 wavefronts whatever its width: the first form of this loop was bound by exactly that, 639 us per 0.8 s.)
__device__ __forceinline__ void etsi_interp2(const float2 *__restrict__ ys, const float *__restrict__ itab, float tau, float half,
                                             float2 &on, float2 &mid)
{
    const float tm = tau - half;
    const float fl_on = floorf(tau), fl_mid = floorf(tm);
    const int i_on = (int)fl_on, i_mid = (int)fl_mid;
    int ph_on = (int)((tau - fl_on) * ED_PH), ph_mid = (int)((tm - fl_mid) * ED_PH);
    ph_on = min(max(ph_on, 0), ED_PH - 1);
    ph_mid = min(max(ph_mid, 0), ED_PH - 1);
    const int lo = i_mid - (EI_TAPS / 2 - 1), base = lo & ~1;
    int D = i_on - i_mid;                                     // 1, 2 (or 3 at 4 samples per symbol): how far the on-time window is ahead
    D = min(max(D, 1), 3);
    float2 w[12];
    const float4 *p = reinterpret_cast<const float4 *>(ys + base);
#pragma unroll
    for (int q = 0; q < 6; ++q) { const float4 v = __ldg(p + q); w[2 * q] = make_float2(v.x, v.y); w[2 * q + 1] = make_float2(v.z, v.w); }
    if (lo & 1) {
#pragma unroll
        for (int i = 0; i < 11; ++i) w[i] = w[i + 1];
    }
    const float *cm = itab + ph_mid * EI_TAPS, *co = itab + ph_on * EI_TAPS;
    float2 a0 = make_float2(0.f, 0.f), a1 = a0, b0 = a0, b1 = a0;
#pragma unroll
    for (int m = 0; m < EI_TAPS; m += 2) {
        a0 = ffma2(splat(cm[m]), w[m], a0);
        a1 = ffma2(splat(cm[m + 1]), w[m + 1], a1);
        const float2 v0 = D == 1 ? w[m + 1] : (D == 2 ? w[m + 2] : w[m + 3]);
        const float2 v1 = D == 1 ? w[m + 2] : (D == 2 ? w[m + 3] : w[m + 4]);
        b0 = ffma2(splat(co[m]), v0, b0);
        b1 = ffma2(splat(co[m + 1]), v1, b1);
    }
    mid = make_float2(a0.x + a1.x, a0.y + a1.y);
    on = make_float2(b0.x + b1.x, b0.y + b1.y);
}

__global__ void __launch_bounds__(128)
etsi_loop_kernel(const float2 *__restrict__ y, long ypitch, int n_streams, long n, float sps, float kp, float ki,
                 const float *__restrict__ itab_g, int n_seg, int seg_syms, float4 *__restrict__ stage,
                 float2 *__restrict__ sym_stage, int32_t *__restrict__ seg_count)
{
    __shared__ float itab[ED_PH * EI_TAPS];
    for (int i = threadIdx.x; i < ED_PH * EI_TAPS; i += blockDim.x) itab[i] = itab_g[i];
    __syncthreads();
    const long gid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int s = (int)(gid / n_seg), j = (int)(gid - (long)s * n_seg);
    if (s >= n_streams) return;
    const float2 *ys = y + (long)s * ypitch;
    const float t_first = 16.0f + sps, nf_end = (float)n - 17.0f, half = 0.5f * sps;
    const float Tj = t_first + (float)j * (float)seg_syms * sps;
    float tau = t_first, emit_from = -1.0f, run_until = nf_end;
    if (j > 0) { tau = Tj - (float)ED_WARM * sps; emit_from = Tj; }
    if (j < n_seg - 1) run_until = fminf(nf_end, Tj + ((float)seg_syms + 2.0f) * sps);
    const int slots = seg_syms + ED_SLACK;
    const long row = ((long)s * n_seg + j) * slots;
    float4 *stg = stage + row;
    float2 *so = sym_stage ? sym_stage + row : nullptr;
    float integ = 0.f, pavg = 0.f;
    float2 prev = make_float2(0.f, 0.f);
    int k = 0, it = 0;
    while (tau < run_until && k < slots) {
        float2 on, mid;
        etsi_interp2(ys, itab, tau, half, on, mid);
        const float p = on.x * on.x + on.y * on.y;
        pavg = (it == 0) ? p : 0.98f * pavg + 0.02f * p;
        const float inv = __fdividef(1.0f, pavg + 1e-20f);
        float e = ((on.x - prev.x) * mid.x + (on.y - prev.y) * mid.y) * inv;       // Gardner, as in etsi_demod_kernel
        e = fminf(fmaxf(e, -1.0f), 1.0f);
        const float dre = on.x * prev.x + on.y * prev.y, dim = on.y * prev.x - on.x * prev.y;
        if (tau >= emit_from) {
            if (so) so[k] = on;
            stg[k] = make_float4(-dim * inv, -dre * inv, tau, 0.f);
            ++k;
        }
        prev = on;
        if (it > 0) { integ += ki * e; tau += sps - kp * e - integ; }
        else tau += sps;
        ++it;
    }
    seg_count[(long)s * n_seg + j] = k;
}

// 8-tap fractional-delay kernels: c[ph][m] weights y[floor(t) - 3 + m] for t = floor(t) + ph / 32
static void make_interp_table(float *tab)
{
    auto bessel_i0 = [](double x) { double s = 1.0, term = 1.0; for (int k = 1; k < 40; ++k) { term *= (x / (2.0 * k)) * (x / (2.0 * k)); s += term; } return s; };
    const double beta = 6.0, hw = EI_TAPS / 2;
    for (int ph = 0; ph < ED_PH; ++ph) {
        double c[EI_TAPS], sum = 0.0;
        for (int m = 0; m < EI_TAPS; ++m) {
            const double u = (double)(m - (EI_TAPS / 2 - 1)) - (double)ph / ED_PH;
            const double r = 1.0 - (u / hw) * (u / hw);
            const double w = r > 0.0 ? bessel_i0(beta * std::sqrt(r)) / bessel_i0(beta) : 0.0;
            const double sc = std::fabs(u) < 1e-12 ? 1.0 : std::sin(M_PI * u) / (M_PI * u);
            c[m] = sc * w; sum += c[m];
        }
        for (int m = 0; m < EI_TAPS; ++m) tab[ph * EI_TAPS + m] = (float)(c[m] / sum);
    }
}

// One CTA per stream: which symbols of every segment are kept, where they go, then the copy into the stream's
// contiguous soft bits (pair k-1 belongs to symbol k >= 1, as in the sequential form) and symbols.
__global__ void __launch_bounds__(256)
etsi_stitch_kernel(int n_seg, int seg_syms, float sps, const float4 *__restrict__ stage, const float2 *__restrict__ sym_stage,
                   const int32_t *__restrict__ seg_count, float2 *__restrict__ sym_out, float *__restrict__ soft_out,
                   long soft_stride, int32_t *__restrict__ n_soft)
{
    __shared__ int keep[ED_SEG_MAX_N], base[ED_SEG_MAX_N + 1];
    const int s = blockIdx.x, slots = seg_syms + ED_SLACK, max_sym = (int)(soft_stride / 2);
    const float4 *st = stage + (long)s * n_seg * slots;
    for (int j = threadIdx.x; j < n_seg; j += blockDim.x) {
        int c = seg_count[(long)s * n_seg + j];
        // the next segment that produced anything decides where this one ends
        float thresh = 3.0e38f;
        for (int jn = j + 1; jn < n_seg; ++jn)
            if (seg_count[(long)s * n_seg + jn] > 0) { thresh = st[(long)jn * slots].z - 0.5f * sps; break; }
        while (c > 0 && st[(long)j * slots + c - 1].z >= thresh) --c;
        keep[j] = c;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int j = 0; j < n_seg; ++j) { base[j] = acc; acc += keep[j]; }
        base[n_seg] = acc;
        const int total = acc < max_sym ? acc : max_sym;
        n_soft[s] = total > 0 ? 2 * (total - 1) : 0;
    }
    __syncthreads();
    float *sb = soft_out + (long)s * soft_stride;
    float2 *so = sym_out ? sym_out + (long)s * max_sym : nullptr;
    // one flat loop over the output symbols (segment of symbol k by binary search in base[]): every iteration is
    // independent, so the loads of many symbols are in flight (a loop over the segments with an inner loop over their
    // symbols serialised ~2 dependent round trips per segment: 122 us per 800 streams x 90 segments)
    const int total = base[n_seg] < max_sym ? base[n_seg] : max_sym;
    for (int k = threadIdx.x; k < total; k += blockDim.x) {
        int lo = 0, hi = n_seg - 1;
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (base[mid] <= k) lo = mid; else hi = mid - 1; }
        const int e = k - base[lo];
        const float4 v = st[(long)lo * slots + e];
        if (k > 0) *reinterpret_cast<float2 *>(sb + 2 * (k - 1)) = make_float2(v.x, v.y);
        if (so) so[k] = sym_stage[((long)s * n_seg + lo) * slots + e];
    }
}

// training sequences (EN 300 392-2 cl. 9.4.4.3): normal n, p, and the extended/tail q
__constant__ int8_t c_ts_pm[3][22] = {
    {1, 1, -1, 1, -1, -1, -1, -1, 1, 1, 1, -1, 1, -1, -1, 1, 1, 1, -1, 1, -1, -1},
    {-1, 1, 1, 1, 1, -1, 1, -1, -1, 1, -1, -1, -1, -1, 1, 1, -1, 1, 1, 1, 1, -1},
    {1, -1, 1, 1, -1, 1, 1, 1, -1, -1, -1, -1, -1, 1, 1, -1, 1, -1, 1, 1, -1, 1}};

// normalised soft correlation with the two normal training sequences at every (even) bit position
__global__ void etsi_ts_corr_kernel(const float *__restrict__ soft, long soft_stride, const int32_t *__restrict__ n_soft,
                                    float *__restrict__ metric)
{
    int s = blockIdx.y;
    long nb = n_soft[s];
    long p = 2 * (blockIdx.x * (long)blockDim.x + threadIdx.x);      // bursts start on symbol boundaries
    if (p + 22 > nb) return;
    const float *b = soft + (long)s * soft_stride + p;
    float c0 = 0.f, c1 = 0.f, en = 0.f;
#pragma unroll
    // soft bits are normalised to ~0.7: clip so that a few large noise products cannot fake a match
    for (int i = 0; i < 22; ++i) { float v = fminf(fmaxf(b[i], -1.0f), 1.0f); c0 += v * c_ts_pm[0][i]; c1 += v * c_ts_pm[1][i]; en += fabsf(v); }
    float best = fmaxf(c0, c1) / (en + 1e-20f);
    metric[(long)s * (soft_stride / 2) + p / 2] = (c1 > c0) ? -best : best;     // sign carries the sequence id
}

// one warp per stream: first position over threshold, refine to the local maximum of the next 4
// symbol positions, emit the burst (TS at bit 244), skip ahead one burst.
__global__ void etsi_find_bursts_kernel(const float *__restrict__ soft, long soft_stride, const int32_t *__restrict__ n_soft,
                                        const float *__restrict__ metric, int n_streams, float thr,
                                        int8_t *__restrict__ blocks, int32_t *__restrict__ burst_pos,
                                        int32_t *__restrict__ n_bursts, int max_bursts, float quant)
{
    const int lane = threadIdx.x & 31;
    const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (s >= n_streams) return;
    const long nb = n_soft[s];
    const long npos = (nb >= 22) ? (nb - 22) / 2 + 1 : 0;
    const float *m = metric + (long)s * (soft_stride / 2);
    const float *b = soft + (long)s * soft_stride;
    int found = 0;
    long i = 0;
    while (i < npos && found < max_bursts) {
        // 512 positions per step (two burst periods): sixteen independent loads in flight instead of one
        float v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) { const long p = i + lane + 32 * u; v[u] = (p < npos) ? fabsf(m[p]) : 0.f; }
        long first = -1;
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const unsigned ball = __ballot_sync(0xffffffffu, v[u] >= thr);
            if (ball && first < 0) first = i + 32 * u + (__ffs(ball) - 1);
        }
        if (first < 0) { i += 512; continue; }
        // local maximum over first .. first+3
        long best = first;
        float bv = fabsf(m[first]);
        for (int d = 1; d < 4; ++d) if (first + d < npos && fabsf(m[first + d]) > bv) { bv = fabsf(m[first + d]); best = first + d; }
        long ts_bit = 2 * best;
        long start = ts_bit - 244;
        if (start >= 0 && start + 510 <= nb) {
            // the two blocks are cut out by etsi_extract_kernel afterwards, in parallel over (stream, burst, bit): doing it
            // here put 14 dependent load -> store rounds of ONE warp between two hits (the search is a serial chain)
            if (lane == 0) burst_pos[(long)s * max_bursts + found] = (int32_t)start;
            ++found;
        }
        i = best + 200;          // next training sequence is 255 symbols later
    }
    if (lane == 0) n_bursts[s] = found;
}


// ---- synchronisation continuous downlink burst (EN 300 392-2 cl. 9.4.4.2.6): q(12) PA(2) frequency correction(80)
// SB1(120) sync training sequence y(38) BB(30) BKN2(216) PB(2) q(10): y starts at bit 214, SB1 at 94, BKN2 at 282.
__constant__ int8_t c_ts_y[38] = {1, 1, -1, -1, -1, -1, -1, 1, 1, -1, -1, 1, 1, 1, -1, -1, 1, 1, 1, -1, 1, -1, -1, 1, 1, 1, -1, -1, -1,
                                  -1, -1, 1, 1, -1, -1, 1, 1, 1};

__global__ void etsi_sync_corr_kernel(const float *__restrict__ soft, long soft_stride, const int32_t *__restrict__ n_soft,
                                      float *__restrict__ metric)
{
    int s = blockIdx.y;
    long nb = n_soft[s];
    long p = 2 * (blockIdx.x * (long)blockDim.x + threadIdx.x);
    if (p + 38 > nb) return;
    const float *b = soft + (long)s * soft_stride + p;
    float c = 0.f, en = 0.f;
#pragma unroll
    for (int i = 0; i < 38; ++i) { float v = fminf(fmaxf(b[i], -1.0f), 1.0f); c += v * c_ts_y[i]; en += fabsf(v); }
    metric[(long)s * (soft_stride / 2) + p / 2] = c / (en + 1e-20f);
}

// one warp per stream: first position over threshold, local maximum of the next 4 symbol positions, emit SB1
// (120 soft bits, input of the BSCH decoder) and BKN2 (216), skip ahead one burst.
__global__ void etsi_find_sync_kernel(const float *__restrict__ soft, long soft_stride, const int32_t *__restrict__ n_soft,
                                      const float *__restrict__ metric, int n_streams, float thr, int8_t *__restrict__ sb1,
                                      int8_t *__restrict__ bkn2, int32_t *__restrict__ burst_pos, int32_t *__restrict__ n_bursts,
                                      int max_bursts, float quant)
{
    const int lane = threadIdx.x & 31;
    const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (s >= n_streams) return;
    const long nb = n_soft[s];
    const long npos = (nb >= 38) ? (nb - 38) / 2 + 1 : 0;
    const float *m = metric + (long)s * (soft_stride / 2);
    const float *b = soft + (long)s * soft_stride;
    int found = 0;
    long i = 0;
    while (i < npos && found < max_bursts) {
        float v[8];                                   // 256 positions per step, as in etsi_find_bursts_kernel
#pragma unroll
        for (int u = 0; u < 8; ++u) { const long p = i + lane + 32 * u; v[u] = (p < npos) ? m[p] : 0.f; }
        long first = -1;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const unsigned ball = __ballot_sync(0xffffffffu, v[u] >= thr);
            if (ball && first < 0) first = i + 32 * u + (__ffs(ball) - 1);
        }
        if (first < 0) { i += 256; continue; }
        long best = first;
        float bv = m[first];
        for (int d = 1; d < 4; ++d) if (first + d < npos && m[first + d] > bv) { bv = m[first + d]; best = first + d; }
        long start = 2 * best - 214;
        if (start >= 0 && start + 510 <= nb) {
            int8_t *o1 = sb1 + ((long)s * max_bursts + found) * 120;
            int8_t *o2 = bkn2 ? bkn2 + ((long)s * max_bursts + found) * 216 : nullptr;
            for (int j = lane; j < 120; j += 32) o1[j] = (int8_t)fminf(fmaxf(rintf(b[start + 94 + j] * quant), -127.f), 127.f);
            if (o2) for (int j = lane; j < 216; j += 32) o2[j] = (int8_t)fminf(fmaxf(rintf(b[start + 282 + j] * quant), -127.f), 127.f);
            if (lane == 0) burst_pos[(long)s * max_bursts + found] = (int32_t)start;
            ++found;
        }
        i = best + 200;
    }
    if (lane == 0) n_bursts[s] = found;
}


// ---- generic training-sequence search (EN 300 392-2 cl. 9.4.4.3): the clipped, normalised soft correlation of the
// kernels above for ANY of the standard's sequences — normal n and p (22 bits), the tail sequence q (22 bits: a
// continuous downlink ends every burst with q(1..10) and opens the next with q(11..22), so q marks the slot boundary),
// the extended sequence x (30 bits, uplink control bursts) and the synchronisation sequence y (38) — returning hit
// positions instead of extracted blocks.
__constant__ int8_t c_seq_len[5] = {22, 22, 22, 30, 38};
__constant__ int8_t c_seq_pm[5][38] = {
    {1, 1, -1, 1, -1, -1, -1, -1, 1, 1, 1, -1, 1, -1, -1, 1, 1, 1, -1, 1, -1, -1},
    {-1, 1, 1, 1, 1, -1, 1, -1, -1, 1, -1, -1, -1, -1, 1, 1, -1, 1, 1, 1, 1, -1},
    {1, -1, 1, 1, -1, 1, 1, 1, -1, -1, -1, -1, -1, 1, 1, -1, 1, -1, 1, 1, -1, 1},
    {1, -1, -1, 1, 1, 1, -1, 1, -1, -1, -1, -1, 1, 1, 1, -1, 1, -1, -1, 1, 1, 1, -1, 1, -1, -1, -1, -1, 1, 1},
    {1, 1, -1, -1, -1, -1, -1, 1, 1, -1, -1, 1, 1, 1, -1, -1, 1, 1, 1, -1, 1, -1, -1, 1, 1, 1, -1, -1, -1,
     -1, -1, 1, 1, -1, -1, 1, 1, 1}};

__global__ void etsi_seq_corr_kernel(const float *__restrict__ soft, long soft_stride, const int32_t *__restrict__ n_soft,
                                     int seq, float *__restrict__ metric)
{
    const int s = blockIdx.y;
    const long nb = n_soft[s];
    const int L = c_seq_len[seq];
    const long p = 2 * (blockIdx.x * (long)blockDim.x + threadIdx.x);
    if (p + L > nb) return;
    const float *b = soft + (long)s * soft_stride + p;
    float c = 0.f, en = 0.f;
    for (int i = 0; i < L; ++i) { const float v = fminf(fmaxf(b[i], -1.0f), 1.0f); c += v * c_seq_pm[seq][i]; en += fabsf(v); }
    metric[(long)s * (soft_stride / 2) + p / 2] = c / (en + 1e-20f);
}

// one warp per stream: first position at or over the threshold, refined to the maximum of the next 4 symbol positions
// (first wins ties), then `gap` symbols are skipped.  Positions are bit offsets into the stream's soft bits.
__global__ void etsi_find_seq_kernel(const int32_t *__restrict__ n_soft, const float *__restrict__ metric, long metric_stride,
                                     int n_streams, int L, float thr, int gap, int32_t *__restrict__ pos, float *__restrict__ corr,
                                     int32_t *__restrict__ n_hits, int max_hits)
{
    const int lane = threadIdx.x & 31;
    const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (s >= n_streams) return;
    const long nb = n_soft[s];
    const long npos = (nb >= L) ? (nb - L) / 2 + 1 : 0;
    const float *m = metric + (long)s * metric_stride;
    int found = 0;
    long i = 0;
    while (i < npos && found < max_hits) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const long p = i + lane + 32 * u; v[u] = (p < npos) ? m[p] : -2.f; }
        long first = -1;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const unsigned ball = __ballot_sync(0xffffffffu, v[u] >= thr);
            if (ball && first < 0) first = i + 32 * u + (__ffs(ball) - 1);
        }
        if (first < 0) { i += 256; continue; }
        long best = first;
        float bv = m[first];
        for (int d = 1; d < 4; ++d) if (first + d < npos && m[first + d] > bv) { bv = m[first + d]; best = first + d; }
        if (lane == 0) { pos[(long)s * max_hits + found] = (int32_t)(2 * best); corr[(long)s * max_hits + found] = bv; }
        ++found;
        i = best + (gap > 0 ? gap : 1);
    }
    if (lane == 0) n_hits[s] = found;
}


// ---- extraction of two bit ranges of every found burst (e.g. the broadcast block BB = bits 230..243 + 266..281 of a
// normal downlink burst, the 30 AACH bits; bits 252..281 of a synchronisation burst), quantised like the blocks
__global__ void etsi_extract_kernel(const float *__restrict__ soft, long soft_stride, const int32_t *__restrict__ n_soft,
                                    const int32_t *__restrict__ burst_pos, const int32_t *__restrict__ n_bursts, int max_bursts,
                                    int off_a, int len_a, int off_b, int len_b, int8_t *__restrict__ out, float quant)
{
    const int s = blockIdx.y;
    const int L = len_a + len_b;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;        // (burst, bit)
    const int bi = idx / L, j = idx - bi * L;
    if (bi >= max_bursts) return;
    int8_t v = 0;
    if (bi < n_bursts[s]) {
        const long src = (long)burst_pos[(long)s * max_bursts + bi] + (j < len_a ? off_a + j : off_b + (j - len_a));
        if (src >= 0 && src < n_soft[s]) v = (int8_t)fminf(fmaxf(rintf(soft[(long)s * soft_stride + src] * quant), -127.f), 127.f);
    }
    out[((long)s * max_bursts + bi) * L + j] = v;
}

// ---- feed-forward carrier-offset correction on the differential symbols.  The soft pair of symbol k is the
// normalised d_k = y_k conj(y_(k-1)):  Re d = -soft[2k+1], Im d = -soft[2k].  For pi/4-DQPSK arg d is one of
// +-pi/4, +-3pi/4 plus theta = 2 pi df T, so d^4 has phase pi + 4 theta whatever the data: theta = (arg sum d^4 - pi) / 4,
// unambiguous for |df| < symbol rate / 8 = 2250 Hz.  One CTA per stream: sum, then rotate every pair by -theta in place.
__global__ void __launch_bounds__(256)
etsi_afc_kernel(float *__restrict__ soft, long soft_stride, const int32_t *__restrict__ n_soft, float *__restrict__ theta_out)
{
    __shared__ double red[2][8];
    __shared__ float rot[2];
    const int s = blockIdx.x;
    float *b = soft + (long)s * soft_stride;
    const int nsym = n_soft[s] / 2;
    double sr = 0.0, si = 0.0;
    for (int k = threadIdx.x; k < nsym; k += blockDim.x) {
        const float2 p = *reinterpret_cast<const float2 *>(b + 2 * k);
        float re = -p.y, im = -p.x;
        const float m = fmaxf(sqrtf(re * re + im * im), 1e-12f);
        re /= m; im /= m;                                   // unit modulus: every symbol one vote
        const float r2 = re * re - im * im, i2 = 2.f * re * im;
        sr += (double)(r2 * r2 - i2 * i2); si += (double)(2.f * r2 * i2);
    }
    for (int o = 16; o > 0; o >>= 1) { sr += __shfl_xor_sync(0xffffffffu, sr, o); si += __shfl_xor_sync(0xffffffffu, si, o); }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sr; red[1][threadIdx.x >> 5] = si; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double tr = 0.0, ti = 0.0;
        for (int w = 0; w < 8; ++w) { tr += red[0][w]; ti += red[1][w]; }
        // phase of -S4 is 4 theta
        const float theta = (nsym > 0) ? 0.25f * atan2f((float)-ti, (float)-tr) : 0.f;
        rot[0] = cosf(theta); rot[1] = sinf(theta);
        if (theta_out) theta_out[s] = theta;
    }
    __syncthreads();
    const float c = rot[0], sn = rot[1];
    for (int k = threadIdx.x; k < nsym; k += blockDim.x) {
        float2 p = *reinterpret_cast<float2 *>(b + 2 * k);
        const float re = -p.y, im = -p.x;
        const float r = re * c + im * sn, i = im * c - re * sn;     // d * exp(-j theta)
        p.x = -i; p.y = -r;
        *reinterpret_cast<float2 *>(b + 2 * k) = p;
    }
}

void make_mf_taps(double sps, float *taps)
{
    const double alpha = 0.35;
    for (int ph = 0; ph < ED_PH; ++ph) {
        for (int j = 0; j < ED_TAPS; ++j) {
            double t = ((double)ph / ED_PH - (double)(j - 15)) / sps;      // in symbols
            double v;
            double at = std::fabs(t);
            if (at < 1e-9) v = 1 - alpha + 4 * alpha / M_PI;
            else if (std::fabs(at - 1 / (4 * alpha)) < 1e-9)
                v = alpha / std::sqrt(2.0) * ((1 + 2 / M_PI) * std::sin(M_PI / (4 * alpha)) + (1 - 2 / M_PI) * std::cos(M_PI / (4 * alpha)));
            else
                v = (std::sin(M_PI * t * (1 - alpha)) + 4 * alpha * t * std::cos(M_PI * t * (1 + alpha))) /
                    (M_PI * t * (1 - (4 * alpha * t) * (4 * alpha * t)));
            taps[ph * ED_TAPS + j] = (float)(v / sps);
        }
    }
}

// segments per stream and symbols per segment the time-parallel forms use for n samples (n_seg < 2: sequential form);
// seg_min = ED_SEG_MIN for the eight-lane kernel, ETS_SEG_MIN for the two-stage form
static void etsi_seg_plan(long n, double sps, int seg_min, int *n_seg, int *seg_syms)
{
    const long n_sym = (long)(((double)n - 33.0) / sps);
    long seg = seg_min;
    if (n_sym > seg * ED_SEG_MAX_N) seg = (n_sym + ED_SEG_MAX_N - 1) / ED_SEG_MAX_N;
    long ns = n_sym > 0 ? (n_sym + seg - 1) / seg : 0;
    *n_seg = (int)ns;
    *seg_syms = (int)seg;
}

static size_t etsi_stage_bytes(int n_streams, int n_seg, int seg_syms, int want_symbols)
{
    const size_t rows = (size_t)n_streams * n_seg;
    return rows * (seg_syms + ED_SLACK) * (sizeof(float4) + (want_symbols ? sizeof(float2) : 0)) + rows * sizeof(int32_t) + 256;
}

// enough for either form: the eight-lane kernel's staging rows, or the two-stage form's rows + the matched-filter output
// y [n_streams][n] complex64 + its interpolation table
int64_t etsi_demod_scratch_bytes(int n_streams, long n, double fs, int want_symbols)
{
    int n_seg, seg_syms, n_seg2, seg_syms2;
    etsi_seg_plan(n, fs / 18000.0, ED_SEG_MIN, &n_seg, &seg_syms);
    etsi_seg_plan(n, fs / 18000.0, ETS_SEG_MIN, &n_seg2, &seg_syms2);
    if (n_seg < 2) return 0;
    const size_t a = etsi_stage_bytes(n_streams, n_seg, seg_syms, want_symbols);
    const size_t b = ((etsi_stage_bytes(n_streams, n_seg2, seg_syms2, want_symbols) + 255) & ~(size_t)255) +
                     (size_t)n_streams * (size_t)((n + 1) & ~1L) * sizeof(float2) + 256;
    return (int64_t)(a > b ? a : b);
}

int launch_etsi_demod(tb200_handle h, const float2 *iq, int n_streams, long n, long stride, double fs, float *taps_dev,
                      float2 *sym, float *soft, long soft_stride, int32_t *n_soft, void *scratch, size_t scratch_bytes, cudaStream_t st)
{
    if (n_streams <= 0) return TB200_OK;
    double sps = fs / 18000.0;
    if (sps < 2.0 || sps > 4.01) return fail(TB200_E_ARG, "tb200_etsi_demod_batch: %.3f samples/symbol outside [2, 4]", sps);
    if (n >= (1L << 30) || (soft_stride & 1) || (reinterpret_cast<uintptr_t>(soft) & 7)) return fail(TB200_E_ARG, "tb200_etsi_demod_batch: n < 2^30, even soft_stride and 8-byte aligned d_soft required");
    if (h->etsi_taps_fs != fs) {                 // the tap table depends on fs only: upload once per handle
        float taps[ED_PH * ED_TAPS], itab[ED_PH * EI_TAPS];
        make_mf_taps(sps, taps);
        make_interp_table(itab);
        TB_CUDA(cudaMemcpyAsync(taps_dev, taps, sizeof(taps), cudaMemcpyHostToDevice, st));
        TB_CUDA(cudaMemcpyAsync(h->etsi_itab, itab, sizeof(itab), cudaMemcpyHostToDevice, st));
        TB_CUDA(cudaStreamSynchronize(st));      // stack buffers
        h->etsi_taps_fs = fs;
    }
    const double loop = h->etsi_open_loop ? 0.0 : 1.0;
    const float kp = (float)(loop * 0.05 * sps), ki = (float)(loop * 0.0005 * sps);
    int n_seg = 0, seg_syms = 0;
    etsi_seg_plan(n, sps, ED_SEG_MIN, &n_seg, &seg_syms);
    const int64_t need = etsi_demod_scratch_bytes(n_streams, n, fs, sym != nullptr);
    const bool seg_ok = scratch && n_seg >= 2 && need > 0 && (size_t)need <= scratch_bytes && (reinterpret_cast<uintptr_t>(scratch) & 15) == 0;
    if (seg_ok && h->etsi_seg_lanes != 8) {
        // two-stage form: matched filter once per sample, then one thread per (stream, segment) interpolating it
        etsi_seg_plan(n, sps, ETS_SEG_MIN, &n_seg, &seg_syms);
        const size_t rows = (size_t)n_streams * n_seg, slots = (size_t)seg_syms + ED_SLACK;
        float4 *stage = reinterpret_cast<float4 *>(scratch);
        float2 *sym_stage = sym ? reinterpret_cast<float2 *>(stage + rows * slots) : nullptr;
        int32_t *seg_count = reinterpret_cast<int32_t *>(reinterpret_cast<char *>(scratch) + rows * slots * (sizeof(float4) + (sym ? sizeof(float2) : 0)));
        char *after = reinterpret_cast<char *>(scratch) + ((etsi_stage_bytes(n_streams, n_seg, seg_syms, sym != nullptr) + 255) & ~(size_t)255);
        float2 *y = reinterpret_cast<float2 *>(after);
        const float *itab = h->etsi_itab;
        dim3 gmf((unsigned)((n + 1023) / 1024), (unsigned)n_streams);
        const long ypitch = (n + 1) & ~1L;                     // even: every row of y starts 16-byte aligned
        etsi_mf_kernel<<<gmf, 256, 0, st>>>(iq, n, stride, taps_dev, y, ypitch);
        TB_LAUNCH_CHECK(h);
        etsi_loop_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, st>>>(y, ypitch, n_streams, n, (float)sps, kp, ki, itab, n_seg, seg_syms,
                                                                         stage, sym_stage, seg_count);
        TB_LAUNCH_CHECK(h);
        etsi_stitch_kernel<<<n_streams, 256, 0, st>>>(n_seg, seg_syms, (float)sps, stage, sym_stage, seg_count, sym, soft, soft_stride, n_soft);
        TB_LAUNCH_CHECK(h);
        return TB200_OK;
    }
    if (seg_ok) {
        // time-parallel, eight lanes per segment (TB200_OPT_ETSI_SEG_LANES = 8): the matched filter inside the loop
        const size_t rows = (size_t)n_streams * n_seg, slots = (size_t)seg_syms + ED_SLACK;
        float4 *stage = reinterpret_cast<float4 *>(scratch);
        float2 *sym_stage = sym ? reinterpret_cast<float2 *>(stage + rows * slots) : nullptr;
        int32_t *seg_count = reinterpret_cast<int32_t *>(reinterpret_cast<char *>(scratch) + rows * slots * (sizeof(float4) + (sym ? sizeof(float2) : 0)));
        const long groups = (long)rows;
        constexpr int LS = 8, GPC = 128 / LS;
        constexpr int ring_bytes = GPC * ED_RING * (int)sizeof(float2);
        cudaError_t e = cudaFuncSetAttribute(etsi_demod_kernel<true, LS>, cudaFuncAttributeMaxDynamicSharedMemorySize, ring_bytes);
        if (e != cudaSuccess) return fail(TB200_E_CUDA, "cudaFuncSetAttribute(etsi_demod): %s", cudaGetErrorString(e));
        etsi_demod_kernel<true, LS><<<(unsigned)((groups + GPC - 1) / GPC), 128, ring_bytes, st>>>(iq, n_streams, n, stride, (float)sps, kp, ki, taps_dev, sym, soft,
                                                                            soft_stride, n_soft, n_seg, seg_syms, stage, sym_stage, seg_count);
        TB_LAUNCH_CHECK(h);
        etsi_stitch_kernel<<<n_streams, 256, 0, st>>>(n_seg, seg_syms, (float)sps, stage, sym_stage, seg_count, sym, soft, soft_stride, n_soft);
        TB_LAUNCH_CHECK(h);
        return TB200_OK;
    }
    int threads = 128, blocks = (n_streams + 15) / 16;              // 8 lanes per stream, 16 streams per CTA
    etsi_demod_kernel<false, 8><<<blocks, threads, 16 * ED_RING * sizeof(float2), st>>>(iq, n_streams, n, stride, (float)sps, kp, ki, taps_dev, sym, soft, soft_stride, n_soft,
                                                         1, 0, nullptr, nullptr, nullptr);
    TB_LAUNCH_CHECK(h);
    return TB200_OK;
}

int launch_etsi_find_bursts(tb200_handle h, const float *soft, int n_streams, long soft_stride, const int32_t *n_soft,
                            float thr, float *metric, int8_t *blocks, int32_t *burst_pos, int32_t *n_bursts, int max_bursts,
                            cudaStream_t st)
{
    if (n_streams <= 0) return TB200_OK;
    dim3 grid((unsigned)((soft_stride / 2 + 255) / 256), (unsigned)n_streams);
    etsi_ts_corr_kernel<<<grid, 256, 0, st>>>(soft, soft_stride, n_soft, metric);
    TB_LAUNCH_CHECK(h);
    int threads = 128, blocks_n = (n_streams * 32 + threads - 1) / threads;
    etsi_find_bursts_kernel<<<blocks_n, threads, 0, st>>>(soft, soft_stride, n_soft, metric, n_streams, thr, blocks, burst_pos,
                                                         n_bursts, max_bursts, 45.0f);
    TB_LAUNCH_CHECK(h);
    return launch_etsi_extract(h, soft, n_streams, soft_stride, n_soft, burst_pos, n_bursts, max_bursts, 14, 216, 282, 216, blocks, st);
}

int launch_etsi_find_sequence(tb200_handle h, const float *soft, int n_streams, long soft_stride, const int32_t *n_soft, int seq,
                              float thr, int gap, float *metric, int32_t *pos, float *corr, int32_t *n_hits, int max_hits,
                              cudaStream_t st)
{
    static const int len[5] = {22, 22, 22, 30, 38};
    if (n_streams <= 0) return TB200_OK;
    dim3 grid((unsigned)((soft_stride / 2 + 255) / 256), (unsigned)n_streams);
    etsi_seq_corr_kernel<<<grid, 256, 0, st>>>(soft, soft_stride, n_soft, seq, metric);
    TB_LAUNCH_CHECK(h);
    int threads = 128, blocks_n = (n_streams * 32 + threads - 1) / threads;
    etsi_find_seq_kernel<<<blocks_n, threads, 0, st>>>(n_soft, metric, soft_stride / 2, n_streams, len[seq], thr, gap, pos, corr,
                                                      n_hits, max_hits);
    TB_LAUNCH_CHECK(h);
    return TB200_OK;
}

int launch_etsi_extract(tb200_handle h, const float *soft, int n_streams, long soft_stride, const int32_t *n_soft,
                        const int32_t *burst_pos, const int32_t *n_bursts, int max_bursts, int off_a, int len_a, int off_b, int len_b,
                        int8_t *out, cudaStream_t st)
{
    if (n_streams <= 0) return TB200_OK;
    const int tot = max_bursts * (len_a + len_b);
    dim3 grid((unsigned)((tot + 127) / 128), (unsigned)n_streams);
    etsi_extract_kernel<<<grid, 128, 0, st>>>(soft, soft_stride, n_soft, burst_pos, n_bursts, max_bursts, off_a, len_a, off_b, len_b,
                                             out, 45.0f);
    TB_LAUNCH_CHECK(h);
    return TB200_OK;
}

int launch_etsi_afc(tb200_handle h, float *soft, int n_streams, long soft_stride, const int32_t *n_soft, float *theta, cudaStream_t st)
{
    if (n_streams <= 0) return TB200_OK;
    etsi_afc_kernel<<<n_streams, 256, 0, st>>>(soft, soft_stride, n_soft, theta);
    TB_LAUNCH_CHECK(h);
    return TB200_OK;
}

int launch_etsi_find_sync(tb200_handle h, const float *soft, int n_streams, long soft_stride, const int32_t *n_soft, float thr,
                          float *metric, int8_t *sb1, int8_t *bkn2, int32_t *burst_pos, int32_t *n_bursts, int max_bursts,
                          cudaStream_t st)
{
    if (n_streams <= 0) return TB200_OK;
    dim3 grid((unsigned)((soft_stride / 2 + 255) / 256), (unsigned)n_streams);
    etsi_sync_corr_kernel<<<grid, 256, 0, st>>>(soft, soft_stride, n_soft, metric);
    TB_LAUNCH_CHECK(h);
    int threads = 128, blocks_n = (n_streams * 32 + threads - 1) / threads;
    etsi_find_sync_kernel<<<blocks_n, threads, 0, st>>>(soft, soft_stride, n_soft, metric, n_streams, thr, sb1, bkn2, burst_pos,
                                                       n_bursts, max_bursts, 45.0f);
    TB_LAUNCH_CHECK(h);
    return TB200_OK;
}

}  // namespace tb200
