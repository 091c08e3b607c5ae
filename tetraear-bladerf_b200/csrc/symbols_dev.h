// symbols_dev.h — device code of the timing pick + slicer (symbols.cu), shared with the fused tail kernel of decode.cu
#pragma once
#include "common.h"

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
__device__ __forceinline__ double block_reduce(double v, double *sh)
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

// One CTA: timing pick, symbols and dibits of chunk `chunk` (processor.py:168-219, :102-166)
__device__ __forceinline__ void symbols_body(int chunk, const double *__restrict__ filtered, long n, long in_stride, int sps, int step,
                                             double *__restrict__ symbols, uint8_t *__restrict__ dibits, long out_stride,
                                             int32_t *__restrict__ n_dibits, int32_t *__restrict__ n_symbols,
                                             int32_t *__restrict__ phase_out, const double *__restrict__ phase_partial,
                                             const double *__restrict__ phase_tail, int n_tiles)
{
    __shared__ double sh[32];
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

}  // namespace tb200
