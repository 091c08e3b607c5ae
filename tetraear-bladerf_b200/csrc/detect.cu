// detect.cu — the scanner's per-dwell signal heuristics, batched over dwells:
// TetraSignalDetector.calculate_power / detect_tetra_modulation / detect_sync_pattern / check_power_stability
// (/root/reference/tetraear/signal/scanner.py:42-147, :204-233).
//
// Per dwell the device produces the raw integers and sums; the host shim finishes the few scalar
// operations (10 log10, std, ratios) in the reference's dtype.  Arithmetic follows numpy's:
//   complex64 input  -> float32 angles / steps / wrap  (np.angle, np.diff, (d + pi) % 2pi - pi in float32),
//   complex128 input -> float64,
//   x / (max|x| + 1e-10) is numpy's complex / real-scalar quotient = a product by the rounded reciprocal,
//   atan2 and |x| are evaluated in double and rounded once (what a correctly-rounded libm gives).
// Kernels: (1) detect_sums: per-window and total sum |x|^2 and max |x|^2, fixed-order partials;
// (2) detect_steps: modulation matches (integer count) + the sync bit stream as 32-bit ballot words;
// (3) detect_sync: best 31-bit sliding match (funnel shift + popc).  All HBM-bound: 8 B/sample read twice.
#include "common.h"
#include "kernels.h"

namespace tb200 {

constexpr int DT_NB = 32;          // blocks per dwell in the two streaming kernels
constexpr int DT_NW = 8;           // max windows of check_power_stability
constexpr int DT_PART = DT_NW + 2; // per-block partial: windows, total, max

template <typename T> struct wrapc;
template <> struct wrapc<float> {
    static __device__ __forceinline__ float pi() { return 3.14159274101257324f; }          // float32(np.pi)
    static __device__ __forceinline__ float twopi() { return 6.28318548202514648f; }       // float32(2 np.pi)
    static __device__ __forceinline__ float qpi() { return 0.785398185253143311f; }        // float32(np.pi / 4)
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float mod(float a, float b) { return fmodf(a, b); }
    static __device__ __forceinline__ float rint_(float a) { return rintf(a); }
};
template <> struct wrapc<double> {
    static __device__ __forceinline__ double pi() { return 3.141592653589793; }
    static __device__ __forceinline__ double twopi() { return 6.283185307179586; }
    static __device__ __forceinline__ double qpi() { return 0.7853981633974483; }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double mod(double a, double b) { return fmod(a, b); }
    static __device__ __forceinline__ double rint_(double a) { return rint(a); }
};

// np.angle in the input's precision
template <typename T> __device__ __forceinline__ T angle_of(T re, T im) { return (T)atan2((double)im, (double)re); }

// (d + pi) % (2 pi) - pi with numpy's floor-mod
template <typename T> __device__ __forceinline__ T wrap_step(T d)
{
    typedef wrapc<T> W;
    const T b = W::twopi();
    T m = W::mod(W::add(d, W::pi()), b);
    if (m != T(0)) { if (m < T(0)) m = W::add(m, b); } else m = T(0);
    return W::sub(m, W::pi());
}

template <typename T>
__global__ void __launch_bounds__(256)
detect_sums_kernel(const T *__restrict__ x, long n, long stride, int num_windows, double *__restrict__ partial)
{
    const int dwell = blockIdx.y, b = blockIdx.x;
    const T *xc = x + (long)dwell * stride * 2;
    const long lo = n * b / DT_NB, hi = n * (b + 1) / DT_NB;
    const long ws = num_windows > 0 ? n / num_windows : 0;
    double acc[DT_NW], tot = 0.0, mx = 0.0;
#pragma unroll
    for (int k = 0; k < DT_NW; ++k) acc[k] = 0.0;
    // a thread's samples come in increasing order, so its window index never decreases: one running sum for the current
    // window, handed to acc[] when the window changes (same additions in the same order as adding every sample to acc[w])
    int cur_w = -1;
    long w_end = lo;                                   // first sample index beyond the current window
    double accw = 0.0;
    auto flush = [&]() {
#pragma unroll
        for (int k = 0; k < DT_NW; ++k) if (k == cur_w && k < num_windows) acc[k] = accw;
    };
    for (long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const double re = (double)xc[2 * i], im = (double)xc[2 * i + 1];
        double p;
        if (sizeof(T) == 4) { float a = (float)sqrt(re * re + im * im); p = (double)__fmul_rn(a, a); }   // np.abs(c64)**2 in float32
        else { double a = hypot(re, im); p = a * a; }
        tot += p;
        mx = fmax(mx, re * re + im * im);
        if (ws > 0) {
            if (i >= w_end) {                          // new window (a block's range rarely spans more than two)
                flush();
                cur_w = (int)(i / ws);
                w_end = (long)(cur_w + 1) * ws;
                accw = 0.0;
            }
            accw += p;
        }
    }
    flush();
    __shared__ double red[8][DT_PART];
    const int wp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double v[DT_PART];
#pragma unroll
    for (int k = 0; k < DT_NW; ++k) v[k] = acc[k];
    v[DT_NW] = tot;
#pragma unroll
    for (int k = 0; k <= DT_NW; ++k)
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    v[DT_NW + 1] = mx;
    if (lane == 0)
        for (int k = 0; k < DT_PART; ++k) red[wp][k] = v[k];
    __syncthreads();
    if (threadIdx.x < DT_PART) {
        double t = red[0][threadIdx.x];
        for (int ww = 1; ww < 8; ++ww) t = (threadIdx.x == DT_NW + 1) ? fmax(t, red[ww][threadIdx.x]) : t + red[ww][threadIdx.x];
        partial[((long)dwell * DT_NB + b) * DT_PART + threadIdx.x] = t;
    }
}

// sums of the block partials in block order; also clears the integer counters of the dwell
__global__ void detect_finish_sums_kernel(const double *__restrict__ partial, long n, int num_windows, tb200_detect_stats *__restrict__ out)
{
    const int dwell = blockIdx.x;
    if (threadIdx.x < DT_PART) {
        double t = 0.0;
        for (int b = 0; b < DT_NB; ++b) {
            double p = partial[((long)dwell * DT_NB + b) * DT_PART + threadIdx.x];
            t = (threadIdx.x == DT_NW + 1) ? fmax(t, p) : t + p;
        }
        if (threadIdx.x < DT_NW) out[dwell].window_sum[threadIdx.x] = t;
        else if (threadIdx.x == DT_NW) out[dwell].power_sum = t;
        else out[dwell].max_abs = sqrt(t);
    }
    if (threadIdx.x == 0) {
        out[dwell].n = n;
        out[dwell].window_size = (int32_t)(num_windows > 0 ? n / num_windows : 0);
        out[dwell].n_steps = 0; out[dwell].matches = 0; out[dwell].n_sync_bits = 0; out[dwell].sync_best = 0;
        out[dwell].reserved = 0;
    }
}


// ---- decisions.  Exact form (any T): the reference's own sequence of operations.
template <typename T>
__device__ __noinline__ bool mod_match_exact(T re0, T im0, T re1, T im1)
{
    typedef wrapc<T> W;
    const T a0 = angle_of<T>(re0, im0), a1 = angle_of<T>(re1, im1);
    const double d = (double)wrap_step<T>(W::sub(a1, a0));
    double best = 1e9;
#pragma unroll
    for (int k = -4; k <= 3; ++k) {
        // np.array([-np.pi, -3*np.pi/4, -np.pi/2, -np.pi/4, 0, np.pi/4, np.pi/2, 3*np.pi/4])
        const double e = (k == -4) ? -3.141592653589793 : (k == -3) ? -2.356194490192345 : (k == -2) ? -1.5707963267948966
                       : (k == -1) ? -0.7853981633974483 : (k == 0) ? 0.0 : (k == 1) ? 0.7853981633974483
                       : (k == 2) ? 1.5707963267948966 : 2.356194490192345;
        best = fmin(best, fabs(e - d));
    }
    return best < 0.39269908169872414;
}

template <typename T>
__device__ __noinline__ bool sync_bit_exact(T re0, T im0, T re1, T im1)
{
    typedef wrapc<T> W;
    const T a0 = angle_of<T>(re0, im0), a1 = angle_of<T>(re1, im1);
    const T d = wrap_step<T>(W::sub(a1, a0));
    return W::rint_(W::div(d, W::qpi())) == T(0);
}

// float32 fast form: atan2f (<= 2 ulp) gives the phase step to ~2e-6; the decision is taken from it unless the
// step sits within 2e-5 of a decision boundary (or of the +-pi wrap), where the exact form is evaluated instead —
// so the result always equals the exact form's.
__device__ __forceinline__ bool mod_match_fast(float a0, float a1, float re0, float im0, float re1, float im1)
{
    const float d = wrap_step<float>(__fsub_rn(a1, a0));
    const float k = fminf(3.f, fmaxf(-4.f, rintf(d * 1.27323954473516f)));
    const float best = fabsf(fmaf(-k, 0.785398163397448f, d));
    if (fabsf(best - 0.392699081698724f) < 2e-5f || fabsf(d) > 3.14157f) return mod_match_exact<float>(re0, im0, re1, im1);
    return best < 0.392699081698724f;
}

__device__ __forceinline__ bool sync_bit_fast(float re0, float im0, float re1, float im1)
{
    const float d = wrap_step<float>(__fsub_rn(atan2f(im1, re1), atan2f(im0, re0)));
    const float q = d * 1.27323954473516f;
    if (fabsf(fabsf(q) - 0.5f) < 2e-5f) return sync_bit_exact<float>(re0, im0, re1, im1);
    return rintf(q) == 0.f;
}

template <typename T>
__global__ void __launch_bounds__(256)
detect_steps_kernel(const T *__restrict__ x, long n, long stride, int sync_step, int do_mod,
                    tb200_detect_stats *__restrict__ out, uint32_t *__restrict__ words, long words_stride)
{
    typedef wrapc<T> W;
    const int dwell = blockIdx.y;
    const T *xc = x + (long)dwell * stride * 2;
    const long gtid = (long)blockIdx.x * blockDim.x + threadIdx.x, gsz = (long)gridDim.x * blockDim.x;
    // ---- detect_tetra_modulation (scanner.py:66-96): steps of the normalised signal within pi/8 of k pi/4, k = -4..3
    if (do_mod) {
        const T s = W::add((T)out[dwell].max_abs, (T)1e-10);
        const T inv = W::div(T(1), s);
        constexpr int S = 8;
        int cnt = 0;
        for (long base = gtid * S; base < n - 1; base += gsz * S) {
            T pre = W::mul(xc[2 * base], inv), pim = W::mul(xc[2 * base + 1], inv);
            float pa = 0.f;
            if (sizeof(T) == 4) pa = atan2f((float)pim, (float)pre);
#pragma unroll
            for (int j = 1; j <= S; ++j) {
                const long i = base + j;
                if (i >= n) break;
                const T cre = W::mul(xc[2 * i], inv), cim = W::mul(xc[2 * i + 1], inv);
                bool hit;
                if (sizeof(T) == 4) {
                    const float ca = atan2f((float)cim, (float)cre);
                    hit = mod_match_fast(pa, ca, (float)pre, (float)pim, (float)cre, (float)cim);
                    pa = ca;
                } else {
                    hit = mod_match_exact<T>(pre, pim, cre, cim);
                }
                pre = cre; pim = cim;
                cnt += hit ? 1 : 0;
            }
        }
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&out[dwell].matches, cnt);
        if (gtid == 0) out[dwell].n_steps = (int32_t)(n - 1);
    }
    // ---- detect_sync_pattern bits (scanner.py:108-128): one bit per step between samples j*step and (j+1)*step
    const long nsym = (n + sync_step - 1) / sync_step;
    if (nsym >= 100) {
        const long nbits = nsym - 1;
        const long nb32 = (nbits + 31) & ~31L;
        for (long j = gtid; j < nb32; j += gsz) {
            bool bit = false;
            if (j < nbits) {
                const long i0 = j * sync_step, i1 = i0 + sync_step;
                if (sizeof(T) == 4) bit = sync_bit_fast((float)xc[2 * i0], (float)xc[2 * i0 + 1], (float)xc[2 * i1], (float)xc[2 * i1 + 1]);
                else bit = sync_bit_exact<T>(xc[2 * i0], xc[2 * i0 + 1], xc[2 * i1], xc[2 * i1 + 1]);
            }
            const unsigned wv = __ballot_sync(0xffffffffu, bit);
            if ((threadIdx.x & 31) == 0) words[(long)dwell * words_stride + (j >> 5)] = wv;
        }
        if (gtid == 0) out[dwell].n_sync_bits = (int32_t)nbits;
    }
}

__global__ void __launch_bounds__(256)
detect_sync_kernel(const uint32_t *__restrict__ words, long words_stride, long n, int sync_step, tb200_detect_stats *__restrict__ out)
{
    const int dwell = blockIdx.y;
    const long nsym = (n + sync_step - 1) / sync_step;
    if (nsym < 100) return;
    const long nbits = nsym - 1, nwin = nbits - 31;          // range(len(bits) - 31): the last window is not tried
    if (nwin <= 0) return;
    const uint32_t *wd = words + (long)dwell * words_stride;
    const long nwords = (nbits + 31) >> 5;
    // pattern bit t = SYNC31[t] (scanner.py:133-134), t = 0 the first bit
    const uint32_t PAT = 0x1 << 1 | 0x1 << 3 | 0x1 << 4 | 0x1 << 7 | 0x1 << 8 | 0x1 << 9 | 0x1 << 13 | 0x1 << 16 | 0x1 << 18 |
                         0x1 << 19 | 0x1 << 22 | 0x1 << 23 | 0x1 << 24 | 0x1 << 28;
    int best = 0;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nwin; i += (long)gridDim.x * blockDim.x) {
        const long wi = i >> 5;
        const uint32_t lo = wd[wi], hi = (wi + 1 < nwords) ? wd[wi + 1] : 0u;
        const uint32_t v = __funnelshift_r(lo, hi, (unsigned)(i & 31));
        best = max(best, 31 - __popc((v ^ PAT) & 0x7fffffffu));
    }
    best = __reduce_max_sync(0xffffffffu, best);
    if ((threadIdx.x & 31) == 0 && best) atomicMax(&out[dwell].sync_best, best);
}

size_t detect_scratch_bytes(int n_dwells, long n)
{
    const long words_stride = (n / 32 + 4);
    return (size_t)n_dwells * ((size_t)DT_NB * DT_PART * sizeof(double) + (size_t)words_stride * 4);
}

int launch_detect(tb200_handle h, const void *x, int dtype, int n_dwells, long n, long stride, double fs, int num_windows,
                  tb200_detect_stats *out, void *scratch, size_t scratch_bytes, cudaStream_t st)
{
    if (n_dwells == 0) return TB200_OK;
    if (num_windows < 0 || num_windows > DT_NW) return fail(TB200_E_ARG, "tb200_detect_batch: num_windows must be 0..%d", DT_NW);
    if (detect_scratch_bytes(n_dwells, n) > scratch_bytes) return fail(TB200_E_ARG, "tb200_detect_batch: batch exceeds the handle scratch");
    double *partial = (double *)scratch;
    uint32_t *words = (uint32_t *)(partial + (size_t)n_dwells * DT_NB * DT_PART);
    const long words_stride = n / 32 + 4;
    int sync_step = (int)(fs / 18000 / 10);                    // max(1, int(sample_rate / symbol_rate / 10)), scanner.py:111
    if (sync_step < 1) sync_step = 1;
    const dim3 grid(DT_NB, n_dwells);
    if (n > 0) {
        if (dtype == TB200_DTYPE_C64) detect_sums_kernel<float><<<grid, 256, 0, st>>>((const float *)x, n, stride, num_windows, partial);
        else detect_sums_kernel<double><<<grid, 256, 0, st>>>((const double *)x, n, stride, num_windows, partial);
        TB_LAUNCH_CHECK(h);
    } else {
        TB_CUDA(cudaMemsetAsync(partial, 0, (size_t)n_dwells * DT_NB * DT_PART * sizeof(double), st));
    }
    detect_finish_sums_kernel<<<n_dwells, 32, 0, st>>>(partial, n, num_windows, out);
    TB_LAUNCH_CHECK(h);
    if (n < 2) return TB200_OK;
    const int do_mod = n >= 1000 ? 1 : 0;                       // scanner.py:66-67
    const int nbx = (int)((n / 8 + 255) / 256 < 1 ? 1 : ((n / 8 + 255) / 256 > 296 ? 296 : (n / 8 + 255) / 256));
    const dim3 g2(nbx, n_dwells);
    if (dtype == TB200_DTYPE_C64) detect_steps_kernel<float><<<g2, 256, 0, st>>>((const float *)x, n, stride, sync_step, do_mod, out, words, words_stride);
    else detect_steps_kernel<double><<<g2, 256, 0, st>>>((const double *)x, n, stride, sync_step, do_mod, out, words, words_stride);
    TB_LAUNCH_CHECK(h);
    const dim3 g3(8, n_dwells);
    detect_sync_kernel<<<g3, 256, 0, st>>>(words, words_stride, n, sync_step, out);
    TB_LAUNCH_CHECK(h);
    return TB200_OK;
}

}  // namespace tb200
