// symbols.cu — open-loop timing pick + differential slicer
// (SignalProcessor.extract_symbols processor.py:168-219, demodulate_dqpsk :102-166).
//
// One CTA per chunk.  The per-chunk feedback quantities (13-phase power argmax, max |x|)
// are block reductions through warp shuffles; sums and comparisons are float64 like numpy.
#include "common.h"
#include "kernels.h"
#include "symbols_dev.h"

namespace tb200 {

__global__ void symbols_kernel(const double *__restrict__ filtered, long n, long in_stride, int sps, int step,
                               double *__restrict__ symbols, uint8_t *__restrict__ dibits, long out_stride,
                               int32_t *__restrict__ n_dibits, int32_t *__restrict__ n_symbols,
                               int32_t *__restrict__ phase_out, const double *__restrict__ phase_partial,
                               const double *__restrict__ phase_tail, int n_tiles)
{
    symbols_body(blockIdx.x, filtered, n, in_stride, sps, step, symbols, dibits, out_stride, n_dibits, n_symbols, phase_out,
                 phase_partial, phase_tail, n_tiles);
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
