// symbols.cu — open-loop timing pick + differential slicer
// (SignalProcessor.extract_symbols processor.py:168-219, demodulate_dqpsk :102-166).
//
// One CTA per chunk.  The per-chunk feedback quantities (13-phase power argmax, max |x|)
// are block reductions through warp shuffles; sums and comparisons are float64 like numpy.
#include "common.h"
#include "kernels.h"

namespace tb200 {

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}

template <bool IS_MAX>
__device__ double block_reduce(double v, double *sh)
{
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = IS_MAX ? warp_max(v) : warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    if (w == 0) {
        v = (lane < nw) ? sh[lane] : (IS_MAX ? -1.0 : 0.0);
        v = IS_MAX ? warp_max(v) : warp_sum(v);
        if (lane == 0) sh[0] = v;
    }
    __syncthreads();
    return sh[0];
}

// slicer thresholds of processor.py:152-161, float64 like np.pi expressions
__device__ __forceinline__ uint8_t slice_dqpsk(double ar, double ai, double pr, double pi_)
{
    const double PI = 3.141592653589793;
    // diff = s[k] * conj(s[k-1])  (numpy: ar*br - ai*bi, ar*bi + ai*br with b = conj(prev))
    double br = pr, bi = -pi_;
    double re = __dsub_rn(__dmul_rn(ar, br), __dmul_rn(ai, bi));
    double im = __dadd_rn(__dmul_rn(ar, bi), __dmul_rn(ai, br));
    double ph = atan2(im, re);
    if (ph < -5 * PI / 8) return 3;
    if (ph < -3 * PI / 8) return 2;
    if (ph < 3 * PI / 8) return 0;
    if (ph < 5 * PI / 8) return 1;
    return 3;
}

__global__ void symbols_kernel(const double *__restrict__ filtered, long n, long in_stride, int sps, int step,
                               double *__restrict__ symbols, uint8_t *__restrict__ dibits, long out_stride,
                               int32_t *__restrict__ n_dibits, int32_t *__restrict__ n_symbols,
                               int32_t *__restrict__ phase_out, const double *__restrict__ phase_partial,
                               const double *__restrict__ phase_tail, int n_tiles)
{
    __shared__ double sh[32];
    int chunk = blockIdx.x;
    const double *x = filtered + (long)chunk * in_stride * 2;
    int best = 0;
    long num = n;
    long stride_s = 1;
    __shared__ double s_pw[TB200_MAX_PHASES];
    __shared__ int s_best;
    if (n > 0 && sps > 1 && phase_partial) {
        // phase powers already summed per (chunk, tile) by the channel-filter kernel: thread p adds its phase
        if (threadIdx.x < sps) {
            int p = threadIdx.x;
            long cnt = (n - p >= 0) ? (n - p) / sps : -1;
            double tot = -2.0;                                   // phases that do not count never win
            if ((p % step) == 0 && cnt > 0) {
                tot = 0.0;
                for (int t = 0; t < n_tiles; ++t) tot += phase_partial[((long)chunk * n_tiles + t) * TB200_MAX_PHASES + p];
                if (phase_tail) tot += phase_tail[(long)chunk * TB200_MAX_PHASES + p];
                tot = tot / (double)cnt;
            }
            s_pw[p] = tot;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double max_power = -1.0;
            int bb = 0;
            for (int p = 0; p < sps; p += step) if (s_pw[p] > max_power) { max_power = s_pw[p]; bb = p; }   // first strict max
            s_best = bb;
        }
        __syncthreads();
        best = s_best;
        num = (n - best) / sps;
        stride_s = sps;
    } else if (n > 0 && sps > 1) {
        double max_power = -1.0;
        for (int p = 0; p < sps; p += step) {
            long cnt = (n - p >= 0) ? (n - p) / sps : -1;
            if (cnt <= 0) continue;
            double tot;
            if (phase_partial) {
                // sums produced by the channel-filter kernel: tiles in index order, then the tail
                tot = 0.0;
                for (int t = 0; t < n_tiles; ++t) tot += phase_partial[((long)chunk * n_tiles + t) * TB200_MAX_PHASES + p];
                tot += phase_tail[(long)chunk * TB200_MAX_PHASES + p];
            } else {
                double acc = 0.0;
                for (long k = threadIdx.x; k < cnt; k += blockDim.x) {
                    double a = hypot(x[2 * (p + k * sps)], x[2 * (p + k * sps) + 1]);
                    acc += a * a;
                }
                tot = block_reduce<false>(acc, sh);
            }
            double power = tot / (double)cnt;
            if (power > max_power) { max_power = power; best = p; }
        }
        num = (n - best) / sps;
        stride_s = sps;
    }
    // gather symbols, max |x|
    double mx = 0.0;
    double *so = symbols ? symbols + (long)chunk * out_stride * 2 : nullptr;
    for (long k = threadIdx.x; k < num; k += blockDim.x) {
        double re = x[2 * (best + k * stride_s)], im = x[2 * (best + k * stride_s) + 1];
        if (so) { so[2 * k] = re; so[2 * k + 1] = im; }
        mx = fmax(mx, hypot(re, im));
    }
    mx = block_reduce<true>(mx, sh);
    double scl = (mx > 0.0) ? 1.0 / mx : 1.0;
    if (dibits && num >= 2) {
        uint8_t *dout = dibits + (long)chunk * out_stride;
        for (long k = 1 + threadIdx.x; k < num; k += blockDim.x) {
            long i1 = best + k * stride_s, i0 = i1 - stride_s;
            double ar = x[2 * i1], ai = x[2 * i1 + 1], pr = x[2 * i0], pi_ = x[2 * i0 + 1];
            if (mx > 0.0) { ar = __dmul_rn(ar, scl); ai = __dmul_rn(ai, scl); pr = __dmul_rn(pr, scl); pi_ = __dmul_rn(pi_, scl); }
            dout[k - 1] = slice_dqpsk(ar, ai, pr, pi_);
        }
    }
    if (threadIdx.x == 0) {
        if (n_dibits) n_dibits[chunk] = (int32_t)(num >= 2 ? num - 1 : 0);
        if (n_symbols) n_symbols[chunk] = (int32_t)num;
        if (phase_out) phase_out[chunk] = best;
    }
}

// standalone demodulate_dqpsk on one array of n complex128 symbols
__global__ void slicer_kernel(const double *__restrict__ s, long n, uint8_t *__restrict__ dibits)
{
    __shared__ double sh[32];
    double mx = 0.0;
    for (long k = threadIdx.x; k < n; k += blockDim.x) mx = fmax(mx, hypot(s[2 * k], s[2 * k + 1]));
    mx = block_reduce<true>(mx, sh);
    double scl = (mx > 0.0) ? 1.0 / mx : 1.0;
    for (long k = 1 + threadIdx.x; k < n; k += blockDim.x) {
        double ar = s[2 * k], ai = s[2 * k + 1], pr = s[2 * k - 2], pi_ = s[2 * k - 1];
        if (mx > 0.0) { ar = __dmul_rn(ar, scl); ai = __dmul_rn(ai, scl); pr = __dmul_rn(pr, scl); pi_ = __dmul_rn(pi_, scl); }
        dibits[k - 1] = slice_dqpsk(ar, ai, pr, pi_);
    }
}

int launch_symbols(tb200_handle h, const double *filtered, int n_chunks, long n, long in_stride, int sps,
                   int phase_step, double *symbols, uint8_t *dibits, long out_stride, int32_t *n_dibits,
                   int32_t *n_symbols, int32_t *phase, cudaStream_t st, const double *phase_partial,
                   const double *phase_tail, int n_tiles)
{
    if (n_chunks <= 0) return TB200_OK;
    symbols_kernel<<<n_chunks, 256, 0, st>>>(filtered, n, in_stride, sps, phase_step, symbols, dibits, out_stride,
                                             n_dibits, n_symbols, phase, phase_partial, phase_tail, n_tiles);
    TB_LAUNCH_CHECK(h);
    return TB200_OK;
}

int launch_slicer(tb200_handle h, const double *symbols, long n, uint8_t *dibits, cudaStream_t st)
{
    if (n < 2) return TB200_OK;
    slicer_kernel<<<1, 1024, 0, st>>>(symbols, n, dibits);
    TB_LAUNCH_CHECK(h);
    return TB200_OK;
}

// SC16Q11 -> complex64: complex64(I, Q) / 32768 (capture.py:255-269; the quotient by 2^15 is exact)
__global__ void sc16_to_c64_kernel(const short2 *__restrict__ in, long n, float2 *__restrict__ out)
{
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    long stride = (long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) { short2 v = in[i]; out[i] = make_float2((float)v.x * (1.0f / 32768.0f), (float)v.y * (1.0f / 32768.0f)); }
}

int launch_sc16_to_c64(tb200_handle h, const void *in, long n, void *out, cudaStream_t st)
{
    if (n <= 0) return TB200_OK;
    long blocks = (n + 255) / 256;
    if (blocks > 148L * 16) blocks = 148L * 16;
    sc16_to_c64_kernel<<<(unsigned)blocks, 256, 0, st>>>((const short2 *)in, n, (float2 *)out);
    TB_LAUNCH_CHECK(h);
    return TB200_OK;
}

// frequency_shift as a stand-alone stage (processor.py:85-100)
template <typename TI>
__global__ void freq_shift_kernel(const TI *__restrict__ x, long n, double W, double fs, double *__restrict__ out)
{
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double xr = (double)x[2 * i], xi = (double)x[2 * i + 1];
    double t = (double)i / fs;
    double s, c;
    sincos(__dmul_rn(W, t), &s, &c);
    out[2 * i] = __dsub_rn(__dmul_rn(xr, c), __dmul_rn(xi, s));
    out[2 * i + 1] = __dadd_rn(__dmul_rn(xr, s), __dmul_rn(xi, c));
}

int launch_freq_shift(tb200_handle h, const void *x, int dtype, long n, double f, double fs, double *out, cudaStream_t st)
{
    if (n <= 0) return TB200_OK;
    double W = -6.283185307179586 * f;
    unsigned blocks = (unsigned)((n + 255) / 256);
    if (dtype == TB200_DTYPE_C64) freq_shift_kernel<float><<<blocks, 256, 0, st>>>((const float *)x, n, W, fs, out);
    else freq_shift_kernel<double><<<blocks, 256, 0, st>>>((const double *)x, n, W, fs, out);
    TB_LAUNCH_CHECK(h);
    return TB200_OK;
}

}  // namespace tb200
