// dsp_exact.cu — bit-faithful ("EXACT" mode) kernels of SignalProcessor.process
// (/root/reference/tetraear/signal/processor.py:221-273).
//
// The reference's decimator is scipy.signal.decimate = 8th-order Chebyshev-I as 4 biquads,
// direct-form II transposed, run forward and backward in FLOAT32 over each chunk
// (sosfiltfilt).  Float32 recursion is not associative, so the only way to reproduce its
// output bit for bit is to run the same recurrence in the same operation order.  That is
// what these kernels do: one thread per (chunk, I/Q) lane, the four sections software-
// pipelined (section s works on sample t-s) so the loop-carried chain is one section deep.
// Parallelism = chunks x 2; this mode is the parity anchor, the FAST mode (dsp_fast.cu) is
// the throughput path.
//
// Layout: forward output is staged in a [time][lane] scratch so that a warp's 32 lanes
// write/read 128 contiguous bytes per time step (coalesced) in both passes.
#include "common.h"
#include "kernels.h"

namespace tb200 {

// ---- rounding-exact scalar arithmetic (never contracted into FMA) --------------------
__device__ __forceinline__ float xmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float xadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float xsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double xmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double xadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double xsub(double a, double b) { return __dsub_rn(a, b); }

template <typename T> struct SosCoef { T c[24]; T zi[8]; };

// one biquad, scipy _sosfilt order:  y = b0 x + z0 ; z0 = (b1 x - a1 y) + z1 ; z1 = b2 x - a2 y
template <typename T>
__device__ __forceinline__ T biquad(const T *c, T &z0, T &z1, T x)
{
    T y = xadd(xmul(c[0], x), z0);
    z0 = xadd(xsub(xmul(c[1], x), xmul(c[4], y)), z1);
    z1 = xsub(xmul(c[2], x), xmul(c[5], y));
    return y;
}

// Skewed 4-section cascade: feed one input per call, get the cascade output of the input
// fed three calls earlier.  `t` is the call index; section s is live for t in [s, m+s).
template <typename T>
struct Cascade4 {
    T z[8];
    T p0, p1, p2;
    __device__ __forceinline__ void init(const T *zi, T first)
    {
#pragma unroll
        for (int i = 0; i < 8; ++i) z[i] = xmul(zi[i], first);
        p0 = p1 = p2 = T(0);
    }
    template <bool EDGE>
    __device__ __forceinline__ T step(const T *c, T in, long t, long m)
    {
        T o0 = p0, o1 = p1, o2 = p2, o3 = T(0);
        if (!EDGE || (t >= 3 && t < m + 3)) o3 = biquad<T>(c + 18, z[6], z[7], p2);
        if (!EDGE || (t >= 2 && t < m + 2)) o2 = biquad<T>(c + 12, z[4], z[5], p1);
        if (!EDGE || (t >= 1 && t < m + 1)) o1 = biquad<T>(c + 6, z[2], z[3], p0);
        if (!EDGE || (t < m)) o0 = biquad<T>(c, z[0], z[1], in);
        p0 = o0; p1 = o1; p2 = o2;
        return o3;
    }
};

// odd extension sample t of the chunk (scipy odd_ext, edge = 27), in the arithmetic of T
template <typename T>
__device__ __forceinline__ T ext_sample(const T *x, long n, long edge, long t)
{
    if (t < edge) return xsub(xmul(T(2), x[0]), x[2 * (edge - t)]);
    if (t < edge + n) return x[2 * (t - edge)];
    return xsub(xmul(T(2), x[2 * (n - 1)]), x[2 * (n - 2 - (t - edge - n))]);
}

// forward pass: x (interleaved complex, component picked by lane parity) -> scratch[t][lane]
template <typename T>
__global__ void decim_exact_fwd(const T *__restrict__ x, long n, long chunk_stride, int n_lanes,
                                SosCoef<T> K, T *__restrict__ scratch)
{
    int lane = blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= n_lanes) return;
    const long edge = 27;
    const long m = n + 2 * edge;
    const T *xp = x + (long)(lane >> 1) * chunk_stride * 2 + (lane & 1);
    Cascade4<T> f;
    f.init(K.zi, ext_sample<T>(xp, n, edge, 0));
    T *out = scratch + lane;
    long t = 0;
    for (; t < 3; ++t) f.template step<true>(K.c, ext_sample<T>(xp, n, edge, t), t, m);
    for (; t < edge; ++t) out[(t - 3) * n_lanes] = f.template step<false>(K.c, ext_sample<T>(xp, n, edge, t), t, m);
    // body: plain samples, 4x unrolled so the loads are issued ahead of the recurrence
    long tb = edge + n;
    for (; t + 4 <= tb; t += 4) {
        T a0 = xp[2 * (t - edge)], a1 = xp[2 * (t + 1 - edge)], a2 = xp[2 * (t + 2 - edge)], a3 = xp[2 * (t + 3 - edge)];
        out[(t - 3) * n_lanes] = f.template step<false>(K.c, a0, t, m);
        out[(t - 2) * n_lanes] = f.template step<false>(K.c, a1, t + 1, m);
        out[(t - 1) * n_lanes] = f.template step<false>(K.c, a2, t + 2, m);
        out[(t) * n_lanes] = f.template step<false>(K.c, a3, t + 3, m);
    }
    for (; t < m; ++t) out[(t - 3) * n_lanes] = f.template step<false>(K.c, ext_sample<T>(xp, n, edge, t), t, m);
    for (; t < m + 3; ++t) out[(t - 3) * n_lanes] = f.template step<true>(K.c, T(0), t, m);
}

// backward pass over scratch (time reversed), keep every q-th sample of the un-padded range
template <typename T>
__global__ void decim_exact_bwd(const T *__restrict__ scratch, long n, int q, int n_lanes, long n_dec,
                                SosCoef<T> K, T *__restrict__ dec_out)
{
    int lane = blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= n_lanes) return;
    const long edge = 27;
    const long m = n + 2 * edge;
    const T *in = scratch + lane;
    T *out = dec_out + (long)(lane >> 1) * n_dec * 2 + (lane & 1);
    Cascade4<T> f;
    f.init(K.zi, in[(m - 1) * n_lanes]);
    // j-th fed sample is ext position m-1-j; the output of step j belongs to position m-1-(j-3)
    long last_kept = ((n - 1) / q) * (long)q;          // largest kept index in [0, n)
    long next_keep = edge + last_kept;                  // ext position of the next sample to keep
    auto emit = [&](long j, T o) {
        long pos = m - 1 - (j - 3);
        if (j >= 3 && pos == next_keep && pos >= edge) {
            out[2 * ((pos - edge) / q)] = o;
            next_keep -= q;
        }
    };
    long j = 0;
    for (; j < 3; ++j) emit(j, f.template step<true>(K.c, in[(m - 1 - j) * n_lanes], j, m));
    // body: 8 loads in flight ahead of the recurrence (the scratch rows are 4*n_lanes bytes apart)
    for (; j + 8 <= m; j += 8) {
        T v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = in[(m - 1 - (j + u)) * n_lanes];
#pragma unroll
        for (int u = 0; u < 8; ++u) emit(j + u, f.template step<false>(K.c, v[u], j + u, m));
    }
    for (; j < m; ++j) emit(j, f.template step<false>(K.c, in[(m - 1 - j) * n_lanes], j, m));
    for (; j < m + 3; ++j) emit(j, f.template step<true>(K.c, T(0), j, m));
}

// ---- channel filter: NCO (processor.py:97-100) + butter(4) filtfilt in float64 (:74-79) --
struct ChanCoef { double b[5], a[5], zi[4]; };

__device__ __forceinline__ double df2t4(const ChanCoef &K, double *Z, double x)
{
    // scipy lfilter: y = Z0 + b0 x ; Zk = (Zk+1 + bk+1 x) - ak+1 y ; Z3 = b4 x - a4 y
    double y = xadd(Z[0], xmul(K.b[0], x));
    Z[0] = xsub(xadd(Z[1], xmul(K.b[1], x)), xmul(K.a[1], y));
    Z[1] = xsub(xadd(Z[2], xmul(K.b[2], x)), xmul(K.a[2], y));
    Z[2] = xsub(xadd(Z[3], xmul(K.b[3], x)), xmul(K.a[3], y));
    Z[3] = xsub(xmul(K.b[4], x), xmul(K.a[4], y));
    return y;
}

// sample i of the (optionally frequency-shifted) decimated chunk, component comp, as double
template <typename TI>
__device__ __forceinline__ double shifted_sample(const TI *x, long i, int comp, double W, double fs)
{
    double xr = (double)x[2 * i], xi = (double)x[2 * i + 1];
    if (W == 0.0) return comp ? xi : xr;
    double t = (double)i / fs;
    double s, c;
    sincos(xmul(W, t), &s, &c);
    return comp ? xadd(xmul(xr, s), xmul(xi, c)) : xsub(xmul(xr, c), xmul(xi, s));
}

template <typename TI>
__device__ __forceinline__ double chan_ext(const TI *x, long n, long edge, long t, int comp, double W, double fs)
{
    if (t >= edge && t < edge + n) return shifted_sample<TI>(x, t - edge, comp, W, fs);
    bool left = t < edge;
    long ie = left ? 0 : n - 1;
    long ii = left ? (edge - t) : (n - 2 - (t - edge - n));
    if (W == 0.0 && sizeof(TI) == 4) {
        // freq_offset == 0: complex64 reaches filtfilt un-promoted, odd_ext rounds in float32
        float e = (float)x[2 * ie + comp], v = (float)x[2 * ii + comp];
        return (double)__fsub_rn(__fmul_rn(2.0f, e), v);
    }
    return xsub(xmul(2.0, shifted_sample<TI>(x, ie, comp, W, fs)), shifted_sample<TI>(x, ii, comp, W, fs));
}

template <typename TI>
__global__ void chan_exact_fwd(const TI *__restrict__ x, long n, long in_stride, int n_lanes,
                               const double *__restrict__ freq_off, double fs, ChanCoef K,
                               double *__restrict__ scratch)
{
    int lane = blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= n_lanes) return;
    int chunk = lane >> 1, comp = lane & 1;
    const long edge = 15, m = n + 2 * edge;
    const TI *xp = x + (long)chunk * in_stride * 2;
    double f = freq_off ? freq_off[chunk] : 0.0;
    // -1j * 2 * np.pi * f  ->  imaginary part  -(2 pi) * f
    double W = (f != 0.0) ? xmul(-6.283185307179586, f) : 0.0;
    double Z[4];
    double first = chan_ext<TI>(xp, n, edge, 0, comp, W, fs);
#pragma unroll
    for (int k = 0; k < 4; ++k) Z[k] = xmul(K.zi[k], first);
    for (long t = 0; t < m; ++t) scratch[t * n_lanes + lane] = df2t4(K, Z, chan_ext<TI>(xp, n, edge, t, comp, W, fs));
}

__global__ void chan_exact_bwd(const double *__restrict__ scratch, long n, int n_lanes, ChanCoef K,
                               double *__restrict__ filtered, long out_stride)
{
    int lane = blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= n_lanes) return;
    const long edge = 15, m = n + 2 * edge;
    double Z[4];
    double first = scratch[(m - 1) * n_lanes + lane];
#pragma unroll
    for (int k = 0; k < 4; ++k) Z[k] = xmul(K.zi[k], first);
    double *out = filtered + (long)(lane >> 1) * out_stride * 2 + (lane & 1);
    long t = m - 1;
    for (; t >= 7; t -= 8) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = scratch[(t - u) * n_lanes + lane];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            double y = df2t4(K, Z, v[u]);
            long tt = t - u;
            if (tt >= edge && tt < edge + n) out[2 * (tt - edge)] = y;
        }
    }
    for (; t >= 0; --t) {
        double y = df2t4(K, Z, scratch[t * n_lanes + lane]);
        if (t >= edge && t < edge + n) out[2 * (t - edge)] = y;
    }
}

// un-filtered path (len <= 15: scipy raises, processor.py:81-83 returns the input): NCO only
template <typename TI>
__global__ void chan_passthrough(const TI *__restrict__ x, long n, long in_stride, int n_chunks,
                                 const double *__restrict__ freq_off, double fs,
                                 double *__restrict__ filtered, long out_stride)
{
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    int chunk = blockIdx.y;
    if (i >= n || chunk >= n_chunks) return;
    double f = freq_off ? freq_off[chunk] : 0.0;
    double W = (f != 0.0) ? xmul(-6.283185307179586, f) : 0.0;
    const TI *xp = x + (long)chunk * in_stride * 2;
    double *o = filtered + (long)chunk * out_stride * 2;
    o[2 * i] = shifted_sample<TI>(xp, i, 0, W, fs);
    o[2 * i + 1] = shifted_sample<TI>(xp, i, 1, W, fs);
}

// ---- host launchers ---------------------------------------------------------------------
template <typename T>
static SosCoef<T> make_sos(const DecimDesign &d);
template <> SosCoef<float> make_sos<float>(const DecimDesign &d)
{
    SosCoef<float> k;
    for (int i = 0; i < 24; ++i) k.c[i] = d.sos32[i];
    for (int i = 0; i < 8; ++i) k.zi[i] = d.zi32[i];
    return k;
}
template <> SosCoef<double> make_sos<double>(const DecimDesign &d)
{
    SosCoef<double> k;
    for (int i = 0; i < 24; ++i) k.c[i] = d.sos64[i];
    for (int i = 0; i < 8; ++i) k.zi[i] = d.zi64[i];
    return k;
}

int launch_decimate_exact(tb200_handle h, const void *x, int dtype, int n_chunks, long n, long chunk_stride,
                          int q, const DecimDesign &d, void *scratch, void *dec_out, long n_dec, cudaStream_t st)
{
    int lanes = 2 * n_chunks;
    int threads = 32;
    int blocks = (lanes + threads - 1) / threads;
    if (dtype == TB200_DTYPE_C64) {
        SosCoef<float> K = make_sos<float>(d);
        decim_exact_fwd<float><<<blocks, threads, 0, st>>>((const float *)x, n, chunk_stride, lanes, K, (float *)scratch);
        TB_LAUNCH_CHECK(h);
        decim_exact_bwd<float><<<blocks, threads, 0, st>>>((const float *)scratch, n, q, lanes, n_dec, K, (float *)dec_out);
        TB_LAUNCH_CHECK(h);
    } else {
        SosCoef<double> K = make_sos<double>(d);
        decim_exact_fwd<double><<<blocks, threads, 0, st>>>((const double *)x, n, chunk_stride, lanes, K, (double *)scratch);
        TB_LAUNCH_CHECK(h);
        decim_exact_bwd<double><<<blocks, threads, 0, st>>>((const double *)scratch, n, q, lanes, n_dec, K, (double *)dec_out);
        TB_LAUNCH_CHECK(h);
    }
    return TB200_OK;
}

int launch_channel_exact(tb200_handle h, const void *x, int dtype, int n_chunks, long n, long in_stride,
                         const double *freq_off, double fs, const ChanDesign &cd, int filtered_flag,
                         double *scratch, double *filtered, long out_stride, cudaStream_t st)
{
    ChanCoef K;
    for (int i = 0; i < 5; ++i) { K.b[i] = cd.b[i]; K.a[i] = cd.a[i]; }
    for (int i = 0; i < 4; ++i) K.zi[i] = cd.zi[i];
    if (!filtered_flag) {
        dim3 grid((unsigned)((n + 255) / 256), (unsigned)n_chunks);
        if (n > 0) {
            if (dtype == TB200_DTYPE_C64)
                chan_passthrough<float><<<grid, 256, 0, st>>>((const float *)x, n, in_stride, n_chunks, freq_off, fs, filtered, out_stride);
            else
                chan_passthrough<double><<<grid, 256, 0, st>>>((const double *)x, n, in_stride, n_chunks, freq_off, fs, filtered, out_stride);
            TB_LAUNCH_CHECK(h);
        }
        return TB200_OK;
    }
    int lanes = 2 * n_chunks, threads = 32, blocks = (lanes + threads - 1) / threads;
    if (dtype == TB200_DTYPE_C64)
        chan_exact_fwd<float><<<blocks, threads, 0, st>>>((const float *)x, n, in_stride, lanes, freq_off, fs, K, scratch);
    else
        chan_exact_fwd<double><<<blocks, threads, 0, st>>>((const double *)x, n, in_stride, lanes, freq_off, fs, K, scratch);
    TB_LAUNCH_CHECK(h);
    chan_exact_bwd<<<blocks, threads, 0, st>>>(scratch, n, lanes, K, filtered, out_stride);
    TB_LAUNCH_CHECK(h);
    return TB200_OK;
}

}  // namespace tb200
