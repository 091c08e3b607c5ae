// psd.cu — waterfall PSD, dB-domain EMA "denoise" and the AFC / signal-gate statistics
// (inline block of CaptureThread.run, /root/reference/tetraear/ui/modern.py:1926-2029, and
// WaterfallWidget.update_spectrum :923-947).
//
// One CTA per frame: the Hann-windowed frame is staged in shared memory (complex128 like
// numpy, which promotes complex64 * float64), transformed by an in-place shared-memory
// radix-4 (+ one radix-2 stage when log2 N is odd) decimation-in-time FFT with a twiddle
// table, then turned into dBFS, fft-shifted, written as float32, and reduced to the 8
// band statistics the AFC loop reads — all without leaving the SM.
#include <cmath>
#include <atomic>
#include "common.h"
#include "kernels.h"

namespace tb200 {

__global__ void psd_tables_kernel(int n_fft, double *__restrict__ window, double2 *__restrict__ twiddle)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_fft) {
        // np.hanning(M): 0.5 + 0.5 cos(pi n / (M-1)), n = 1-M, 3-M, ...
        window[i] = 0.5 + 0.5 * cospi((double)(2 * i + 1 - n_fft) / (double)(n_fft - 1));
    }
    if (i < n_fft / 2 + n_fft / 4 + 1) {
        double s, c;
        sincospi(-2.0 * (double)i / (double)n_fft, &s, &c);   // exp(-2 pi j i / N)
        twiddle[i] = make_double2(c, s);
    }
}

template <typename T> struct V2;
template <> struct V2<double> { typedef double2 type; };
template <> struct V2<float> { typedef float2 type; };
__device__ __forceinline__ double2 mk2(double a, double b) { return make_double2(a, b); }
__device__ __forceinline__ float2 mk2(float a, float b) { return make_float2(a, b); }
template <typename C> __device__ __forceinline__ C cmul(C a, C b) { return mk2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double db20(double mag) { return 20.0 * log10(mag + 1e-20); }
__device__ __forceinline__ float db20(float mag) { return 20.0f * log10f(mag + 1e-20f); }
__device__ __forceinline__ double mag2(double2 v) { return hypot(v.x, v.y); }
__device__ __forceinline__ float mag2(float2 v) { return sqrtf(v.x * v.x + v.y * v.y); }

struct BandParams {
    int start, end;          // in-channel bins  [start, end)
    int n_lo, n_hi0;         // noise bins [0, n_lo) and [n_hi0, N)
    double freq_step;        // 1 / (N * (1/fs))
};

extern __shared__ __align__(16) unsigned char psd_smem_raw[];

// T = double: numpy's arithmetic (complex128 FFT, float64 dB) — used for the reference's 2048-point frames.
// T = float : the dense "config" mode (tolerance: <= 1e-4 of full scale in amplitude, north_star).
template <typename T>
__global__ void psd_kernel(const float2 *__restrict__ iq, long frame_stride, int n_fft, int log2n,
                           const double *__restrict__ window, const double2 *__restrict__ tw, BandParams bp,
                           float *__restrict__ power, double *__restrict__ stats, int avg)
{
    typedef typename V2<T>::type C;
    C *x = reinterpret_cast<C *>(psd_smem_raw);
    __shared__ double red[3][32];
    __shared__ int redi[32];
    const long frame = blockIdx.x;                  // OUTPUT frame: the linear (Welch) average of `avg` input frames
    const int N = n_fft;
    T acc[8];                                       // sum over the averaged frames of (|X| / N)^2 for this thread's bins
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = (T)0;
    for (int ka = 0; ka < avg; ++ka) {
    const float2 *in = iq + (frame * avg + ka) * frame_stride;
    if (ka) __syncthreads();                        // the previous frame's bins have been read
    // load, window, bit-reverse (radix-2 ordering)
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        float2 v = in[i];
        T w = (T)window[i];
        int r = __brev((unsigned)i) >> (32 - log2n);
        x[r] = mk2((T)v.x * w, (T)v.y * w);
    }
    __syncthreads();
    // first stage radix-2 when log2n is odd, then radix-4 stages (two radix-2 levels fused)
    int len = 1;                 // current sub-transform length
    if (log2n & 1) {
        for (int i = threadIdx.x; i < N / 2; i += blockDim.x) {
            C a = x[2 * i], b = x[2 * i + 1];
            x[2 * i] = mk2(a.x + b.x, a.y + b.y);
            x[2 * i + 1] = mk2(a.x - b.x, a.y - b.y);
        }
        len = 2;
        __syncthreads();
    }
    while (len < N) {
        // combine 4 sub-transforms of length len into one of 4*len (bit-reversed input order:
        // blocks are  E0=[even-even], E1=[even-odd], O0, O1  -> radix-2 twice)
        const int L4 = 4 * len;
        const int tstep1 = N / (2 * len);      // twiddle stride for the length-2len level
        const int tstep2 = N / L4;             // for the length-4len level
        for (int i = threadIdx.x; i < N / 4; i += blockDim.x) {
            int blk = i / len, k = i - blk * len;
            int base = blk * L4 + k;
            C a = x[base], b = x[base + len], c = x[base + 2 * len], d = x[base + 3 * len];
            double2 t1 = tw[k * tstep1], t2a = tw[k * tstep2], t2b = tw[(k + len) * tstep2];
            C w1 = mk2((T)t1.x, (T)t1.y), w2a = mk2((T)t2a.x, (T)t2a.y), w2b = mk2((T)t2b.x, (T)t2b.y);
            C bt = cmul(b, w1), dt = cmul(d, w1);
            C e0 = mk2(a.x + bt.x, a.y + bt.y), e1 = mk2(a.x - bt.x, a.y - bt.y);
            C o0 = mk2(c.x + dt.x, c.y + dt.y), o1 = mk2(c.x - dt.x, c.y - dt.y);
            C p0 = cmul(o0, w2a), p1 = cmul(o1, w2b);
            x[base] = mk2(e0.x + p0.x, e0.y + p0.y);
            x[base + 2 * len] = mk2(e0.x - p0.x, e0.y - p0.y);
            x[base + len] = mk2(e1.x + p1.x, e1.y + p1.y);
            x[base + 3 * len] = mk2(e1.x - p1.x, e1.y - p1.y);
        }
        len = L4;
        __syncthreads();
    }
    if (avg > 1) {
        const T invNa = (T)1 / (T)N;
        int i = 0;
        for (int j = threadIdx.x; j < N; j += blockDim.x, ++i) {
            const T mg = mag2(x[(j + N / 2) & (N - 1)]) * invNa;
            acc[i & 7] += mg * mg;
        }
    }
    }   // averaged frames
    // dBFS, fftshift, band statistics
    const T invN = (T)1 / (T)N;   // |X| / n_fft : power-of-two N, the product is exact
    double sig = 0.0, noise = 0.0, pk = -1e300;
    int pki = 0x7fffffff;
    int ia = 0;
    for (int j = threadIdx.x; j < N; j += blockDim.x, ++ia) {
        C v = x[(j + N / 2) & (N - 1)];
        // avg == 1: the reference's formula 20 log10(|X|/N + 1e-20) (ui/modern.py:1941); avg > 1: the same with |X|/N
        // replaced by its RMS over the averaged frames (linear power average, then dB)
        double p = avg > 1 ? (double)db20((T)sqrt(acc[ia & 7] / (T)avg)) : (double)db20(mag2(v) * invN);
        if (power) power[frame * N + j] = (float)p;
        if (j >= bp.start && j < bp.end) {
            sig += p;
            if (p > pk) { pk = p; pki = j; }
        }
        if (j < bp.n_lo || j >= bp.n_hi0) noise += p;
    }
    if (!stats) return;
    // block reductions (sum, sum, max + first argmax)
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int o = 16; o > 0; o >>= 1) {
        sig += __shfl_down_sync(0xffffffffu, sig, o);
        noise += __shfl_down_sync(0xffffffffu, noise, o);
        double opk = __shfl_down_sync(0xffffffffu, pk, o);
        int oi = __shfl_down_sync(0xffffffffu, pki, o);
        if (opk > pk || (opk == pk && oi < pki)) { pk = opk; pki = oi; }
    }
    if (lane == 0) { red[0][w] = sig; red[1][w] = noise; red[2][w] = pk; redi[w] = pki; }
    __syncthreads();
    if (w == 0) {
        sig = lane < nw ? red[0][lane] : 0.0;
        noise = lane < nw ? red[1][lane] : 0.0;
        pk = lane < nw ? red[2][lane] : -1e300;
        pki = lane < nw ? redi[lane] : 0x7fffffff;
        for (int o = 16; o > 0; o >>= 1) {
            sig += __shfl_down_sync(0xffffffffu, sig, o);
            noise += __shfl_down_sync(0xffffffffu, noise, o);
            double opk = __shfl_down_sync(0xffffffffu, pk, o);
            int oi = __shfl_down_sync(0xffffffffu, pki, o);
            if (opk > pk || (opk == pk && oi < pki)) { pk = opk; pki = oi; }
        }
        if (lane == 0) {
            double *s = stats + frame * 8;
            int nband = bp.end - bp.start;
            int nnoise = bp.n_lo + (N - bp.n_hi0);
            if (nband <= 0) { for (int k = 0; k < 8; ++k) s[k] = 0.0; return; }
            double signal_power = sig / (double)nband;
            double noise_floor = nnoise > 0 ? noise / (double)nnoise : -100.0;     // modern.py:1989-1992
            double snr = signal_power - noise_floor;
            double off = (double)(pki - N / 2) * bp.freq_step;                      // freqs[peak_idx]
            bool present = (snr > 15.0) && (pk > -70.0) && ((pk - signal_power) > 3.0);   // :2006
            s[0] = signal_power; s[1] = pk; s[2] = (double)pki; s[3] = off; s[4] = noise_floor; s[5] = snr;
            s[6] = present ? 1.0 : 0.0;
            s[7] = (present && pk > -70.0) ? off : 0.0;                             // :2028
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Dense config mode, N = 4096, float32: three radix-16 passes with the 16 points of each small DFT in registers
// (4096 = 16 x 16 x 16), two shared-memory exchanges, persistent CTAs that keep the twiddle tables in shared
// memory and each thread's 16 window coefficients in registers.  With n = 256 a + b and k = c + 16 d:
//   X[c + 16 d] = sum_b W256^(b d) [ W4096^(b c) sum_a W16^(a c) x[256 a + b] ]            (pass 1, thread b)
// and the 256-point transforms over b = 16 e + f, d = g + 16 h split the same way:
//   Z[c][g][f] = W256^(f g) sum_e W16^(e g) Y[c][16 e + f]   (pass 2, thread (c, f)),
//   X[c + 16 g + 256 h] = sum_f W16^(f h) Z[c][g][f]         (pass 3, thread t = 16 g + c: k = t + 256 h, coalesced).
constexpr int P4_PITCH = 17;                                  // Z rows of 16 padded to 17: conflict-free pass-3 reads
constexpr int P4_BUF = 256 * P4_PITCH;

__device__ __forceinline__ float2 f2add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 f2sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 f2mul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// 4-point DFT in place: (a, b, c, d) <- (X0, X1, X2, X3)
__device__ __forceinline__ void dft4(float2 &a, float2 &b, float2 &c, float2 &d)
{
    const float2 s0 = f2add(a, c), d0 = f2sub(a, c), s1 = f2add(b, d), d1 = f2sub(b, d);
    const float2 mj = make_float2(d1.y, -d1.x);               // -j d1
    a = f2add(s0, s1); c = f2sub(s0, s1);
    b = f2add(d0, mj); d = f2sub(d0, mj);
}
// 16-point DFT of v[0..15] (natural order in); X[k] is left at position 4 (k & 3) + (k >> 2)
__device__ __forceinline__ void dft16(float2 (&v)[16])
{
#pragma unroll
    for (int j = 0; j < 4; ++j) dft4(v[j], v[j + 4], v[j + 8], v[j + 12]);
    const float c1 = 0.923879532511287f, s1 = 0.382683432365090f, r2 = 0.707106781186548f;
    const float2 W1 = make_float2(c1, -s1), W2 = make_float2(r2, -r2), W3 = make_float2(s1, -c1), W6 = make_float2(-r2, -r2),
                 W9 = make_float2(-c1, s1);
    v[5] = f2mul(v[5], W1); v[6] = f2mul(v[6], W2); v[7] = f2mul(v[7], W3);
    v[9] = f2mul(v[9], W2); v[10] = make_float2(v[10].y, -v[10].x); v[11] = f2mul(v[11], W6);
    v[13] = f2mul(v[13], W3); v[14] = f2mul(v[14], W6); v[15] = f2mul(v[15], W9);
#pragma unroll
    for (int q = 0; q < 4; ++q) dft4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}
__device__ __forceinline__ int dft16_index(int p) { return (p >> 2) + 4 * (p & 3); }

template <bool AVG>                                        // AVG = false: no accumulators (avg == 1)
__global__ void __launch_bounds__(256, 3)
psd4096_f32_kernel(const float2 *__restrict__ iq, long frame_stride, long n_frames, const double *__restrict__ window,
                   const double2 *__restrict__ tw, BandParams bp, float *__restrict__ power, double *__restrict__ stats, int avg)
{
    constexpr int N = 4096;
    float2 *buf = reinterpret_cast<float2 *>(psd_smem_raw);   // pass exchange: Y[c][b] (pitch 256), then Z[t][f] (pitch 17)
    float2 *T1 = buf + P4_BUF;                                // W4096^(b c), [c][b]
    float2 *T2 = T1 + N;                                      // W256^(f g),  [g][f]
    __shared__ double red[3][8];
    __shared__ int redi[8];
    const int t = threadIdx.x;
    // the double table holds exp(-2 pi j i / N) for i <= 3N/4 only: use W^(i + N/2) = -W^i
    for (int i = t; i < N; i += 256) {
        const int e = ((i & 255) * (i >> 8)) & (N - 1);
        const double2 w = tw[e & (N / 2 - 1)];
        const float sg = e >= N / 2 ? -1.f : 1.f;
        T1[i] = make_float2(sg * (float)w.x, sg * (float)w.y);
    }
    {
        const int e = (16 * (t & 15) * (t >> 4)) & (N - 1);
        const double2 w = tw[e & (N / 2 - 1)];
        const float sg = e >= N / 2 ? -1.f : 1.f;
        T2[t] = make_float2(sg * (float)w.x, sg * (float)w.y);
    }
    float win[16];
#pragma unroll
    for (int a = 0; a < 16; ++a) win[a] = (float)window[256 * a + t];
    __syncthreads();
    const float invN = 1.0f / (float)N;
    const int lane = t & 31, wp = t >> 5;
    for (long frame = blockIdx.x; frame < n_frames; frame += gridDim.x) {      // OUTPUT frames (averages of `avg` input frames)
        float acc[AVG ? 16 : 1];
#pragma unroll
        for (int p = 0; p < (AVG ? 16 : 1); ++p) acc[p] = 0.f;
        float2 v[16];
        for (int ka = 0; ka < (AVG ? avg : 1); ++ka) {
        const float2 *in = iq + (AVG ? frame * avg + ka : frame) * frame_stride;
        if (ka) __syncthreads();                               // buf is reused
#pragma unroll
        for (int a = 0; a < 16; ++a) { const float2 x = in[256 * a + t]; v[a] = make_float2(x.x * win[a], x.y * win[a]); }
        dft16(v);
#pragma unroll
        for (int p = 0; p < 16; ++p) {
            const int c = dft16_index(p);
            buf[c * 256 + t] = c ? f2mul(v[p], T1[c * 256 + t]) : v[p];
        }
        __syncthreads();
        {
            const int c2 = t >> 4, f = t & 15;
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = buf[c2 * 256 + 16 * e + f];
            dft16(v);
            __syncthreads();
#pragma unroll
            for (int p = 0; p < 16; ++p) {
                const int g = dft16_index(p);
                buf[(g * 16 + c2) * P4_PITCH + f] = g ? f2mul(v[p], T2[g * 16 + f]) : v[p];
            }
        }
        __syncthreads();
#pragma unroll
        for (int f = 0; f < 16; ++f) v[f] = buf[t * P4_PITCH + f];
        dft16(v);
        if (AVG) {
#pragma unroll
            for (int p = 0; p < 16; ++p) { const float mg = mag2(v[p]) * invN; acc[AVG ? p : 0] = fmaf(mg, mg, acc[AVG ? p : 0]); }
        }
        }   // averaged frames
        // dBFS, fftshift, band statistics
        double sig = 0.0, noise = 0.0, pk = -1e300;
        int pki = 0x7fffffff;
#pragma unroll
        for (int p = 0; p < 16; ++p) {
            const int k = t + 256 * dft16_index(p);
            const int j = (k + N / 2) & (N - 1);
            const double pdb = AVG ? (double)db20(sqrtf(acc[AVG ? p : 0] / (float)avg)) : (double)db20(mag2(v[p]) * invN);
            if (power) power[frame * N + j] = (float)pdb;
            if (j >= bp.start && j < bp.end) {
                sig += pdb;
                if (pdb > pk || (pdb == pk && j < pki)) { pk = pdb; pki = j; }
            }
            if (j < bp.n_lo || j >= bp.n_hi0) noise += pdb;
        }
        if (stats) {
            for (int o = 16; o > 0; o >>= 1) {
                sig += __shfl_down_sync(0xffffffffu, sig, o);
                noise += __shfl_down_sync(0xffffffffu, noise, o);
                const double opk = __shfl_down_sync(0xffffffffu, pk, o);
                const int oi = __shfl_down_sync(0xffffffffu, pki, o);
                if (opk > pk || (opk == pk && oi < pki)) { pk = opk; pki = oi; }
            }
            if (lane == 0) { red[0][wp] = sig; red[1][wp] = noise; red[2][wp] = pk; redi[wp] = pki; }
            __syncthreads();
            if (t == 0) {
                sig = 0.0; noise = 0.0; pk = -1e300; pki = 0x7fffffff;
                for (int ww = 0; ww < 8; ++ww) {
                    sig += red[0][ww]; noise += red[1][ww];
                    if (red[2][ww] > pk || (red[2][ww] == pk && redi[ww] < pki)) { pk = red[2][ww]; pki = redi[ww]; }
                }
                double *s = stats + frame * 8;
                const int nband = bp.end - bp.start, nnoise = bp.n_lo + (N - bp.n_hi0);
                if (nband <= 0) { for (int kk = 0; kk < 8; ++kk) s[kk] = 0.0; }
                else {
                    const double signal_power = sig / (double)nband;
                    const double noise_floor = nnoise > 0 ? noise / (double)nnoise : -100.0;     // modern.py:1989-1992
                    const double snr = signal_power - noise_floor;
                    const double off = (double)(pki - N / 2) * bp.freq_step;                      // freqs[peak_idx]
                    const bool present = (snr > 15.0) && (pk > -70.0) && ((pk - signal_power) > 3.0);   // :2006
                    s[0] = signal_power; s[1] = pk; s[2] = (double)pki; s[3] = off; s[4] = noise_floor; s[5] = snr;
                    s[6] = present ? 1.0 : 0.0;
                    s[7] = (present && pk > -70.0) ? off : 0.0;                                   // :2028
                }
            }
        }
        __syncthreads();                                       // buf (and red) are reused by the next frame
    }
}

// EMA across frames, dB domain: S = S (1 - a) + P a ; first frame S = P  (modern.py:929-935).
// The recurrence forgets its past as (1-a)^k, so the frame axis is cut into groups of EMA_G frames and each
// group re-runs the EMA_W frames before it as warm-up (exact when that reaches frame 0, otherwise the
// truncated history is below (1-a)^EMA_W of a dB difference: 9e-10 for a = 0.15).  Parallelism:
// groups x bins threads instead of one thread per bin walking every frame.
// Warm-up length = frames until (1-a)^W < 1e-9 (128 for a = 0.15); groups of 4 W frames read the input 1.25 times.
// A smoothing factor so small that the warm-up would exceed 4096 frames runs as ONE group (exact, bins-parallel only).

// One launch covers the output frames [f_begin, f_end) (a batch, see launch_psd); warm-up frames may lie before f_begin.
__global__ void psd_ema_kernel(const float *__restrict__ power, long n_frames, int n_fft, double alpha, int first,
                               const double *__restrict__ state_in, double *__restrict__ state_out,
                               float *__restrict__ smoothed, long EMA_G, long EMA_W, long f_begin, long f_end)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_fft) return;
    const long g0 = f_begin + (long)(gridDim.y - 1 - blockIdx.y) * EMA_G;      // last group first: see launch_psd
    const long g1 = g0 + EMA_G < f_end ? g0 + EMA_G : f_end;
    long f = g0 - EMA_W;
    double S;
    if (f <= 0) {
        f = 0;
        if (first) { S = (double)power[j]; if (g0 == 0) smoothed[j] = (float)S; f = 1; }
        else S = state_in[j];
    } else {
        S = (double)power[f * n_fft + j];      // any start value: its weight at g0 is (1-a)^EMA_W
        ++f;
    }
    const double om = 1.0 - alpha;
    for (; f < g1; ++f) {
        double p = (double)power[f * n_fft + j];
        S = __dadd_rn(__dmul_rn(S, om), __dmul_rn(p, alpha));
        if (f >= g0) smoothed[f * n_fft + j] = (float)S;
    }
    if (g1 == n_frames) state_out[j] = S;
}

int launch_psd(tb200_handle h, const float *iq, int64_t n_frames, int64_t frame_stride, int n_fft, double fs,
               double alpha, int first, float *power, float *smoothed, double *ema_state, double *stats,
               double *tables, cudaStream_t st, int use_f32, int avg)
{
    // linear (Welch) averaging: every output frame is the power average of `avg` consecutive input frames
    if (avg < 1) avg = 1;
    n_frames /= avg;
    if (n_frames <= 0) return TB200_OK;
    int log2n = 0;
    while ((1 << log2n) < n_fft) ++log2n;
    double *window = tables;
    double2 *tw = reinterpret_cast<double2 *>(tables + n_fft);
    if (h->psd_tables_n != n_fft) {
        psd_tables_kernel<<<(n_fft + 255) / 256, 256, 0, st>>>(n_fft, window, tw);
        TB_LAUNCH_CHECK(h);
        h->psd_tables_n = n_fft;
    }
    BandParams bp;
    int center = n_fft / 2;
    double freq_resolution = fs / n_fft;
    int bins = (int)(25000 / freq_resolution);
    bp.start = center - bins / 2 > 0 ? center - bins / 2 : 0;
    bp.end = center + bins / 2 < n_fft ? center + bins / 2 : n_fft;
    bp.n_lo = bp.start - 10 > 0 ? bp.start - 10 : 0;
    bp.n_hi0 = bp.end + 10 < n_fft ? bp.end + 10 : n_fft;
    bp.freq_step = 1.0 / ((double)n_fft * (1.0 / fs));
    // EMA grouping (psd_ema_kernel): state_in and state_out may alias only when a single group runs; otherwise read a snapshot
    double *state_in = ema_state;
    long warm = 1, group = n_frames;
    if (smoothed) {
        if (alpha < 1.0 && alpha > 0.0) warm = (long)std::ceil(-20.72 / std::log1p(-alpha));     // (1-a)^warm < 1e-9
        group = 4 * warm;
        // short runs (e.g. averaged spectra): smaller groups, down to one warm-up length, until the launch has ~1000 CTAs
        while (group > warm && warm > 0 && ((n_frames + group - 1) / group) * ((n_fft + 127) / 128) < 1024) group /= 2;
        if (group < warm) group = warm;
        if (!(alpha > 0.0) || warm > 4096 || group >= n_frames) { group = n_frames; warm = 0; }
        const long groups = (n_frames + group - 1) / group;
        if (groups > 1 && !first) {
            state_in = tables + n_fft + 2 * (n_fft / 2 + n_fft / 4 + 1);       // spare n_fft doubles behind the tables
            TB_CUDA(cudaMemcpyAsync(state_in, ema_state, sizeof(double) * n_fft, cudaMemcpyDeviceToDevice, st));
        }
    }
    // One batch: cutting the frame axis into L2-sized batches (FFT then smoothing per ~32 MB of spectra) was measured and
    // lost — each launch then has too few CTAs (17 x 2 launches: 2.9 ms against 1.0 ms).  Instead the smoothing pass
    // walks its groups from the LAST frame back, so that it starts on the spectra the FFT pass wrote most recently and
    // that are still in the 126 MB L2.
    const long batch = n_frames;
    for (long f0 = 0; f0 < n_frames; f0 += batch) {
    const long nb = f0 + batch < n_frames ? batch : n_frames - f0;
    const float2 *iq_b = reinterpret_cast<const float2 *>(iq) + f0 * avg * frame_stride;
    float *power_b = power + f0 * n_fft;
    double *stats_b = stats ? stats + f0 * 8 : nullptr;
    if (use_f32 && n_fft == 4096) {
        const size_t sm4 = sizeof(float2) * (size_t)(P4_BUF + 4096 + 256);
        static std::atomic<bool> attr4_dev[64];  // per device
        std::atomic<bool> &attr4 = attr4_dev[h->cfg.device & 63];
        if (!attr4) {
            cudaError_t e = cudaFuncSetAttribute(psd4096_f32_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm4);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(psd4096_f32_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm4);
            if (e != cudaSuccess) return fail(TB200_E_CUDA, "cudaFuncSetAttribute(psd4096_f32_kernel): %s", cudaGetErrorString(e));
            attr4 = true;
        }
        const long ctas = nb < 3 * 148 ? nb : 3 * 148;            // persistent: 3 CTAs per SM
        if (avg > 1) psd4096_f32_kernel<true><<<(unsigned)ctas, 256, sm4, st>>>(iq_b, frame_stride, nb, window, tw, bp, power_b, stats_b, avg);
        else psd4096_f32_kernel<false><<<(unsigned)ctas, 256, sm4, st>>>(iq_b, frame_stride, nb, window, tw, bp, power_b, stats_b, avg);
        TB_LAUNCH_CHECK(h);
    } else {
    size_t smem = (use_f32 ? sizeof(float2) : sizeof(double2)) * (size_t)n_fft;
    if (smem > 48 * 1024) {
        cudaError_t e = use_f32 ? cudaFuncSetAttribute(psd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                                : cudaFuncSetAttribute(psd_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(TB200_E_CUDA, "cudaFuncSetAttribute(psd_kernel): %s", cudaGetErrorString(e));
    }
    int threads = n_fft >= 2048 ? 512 : (n_fft >= 512 ? 128 : 64);
    if (use_f32)
        psd_kernel<float><<<(unsigned)nb, threads, smem, st>>>(iq_b, frame_stride, n_fft, log2n, window, tw, bp, power_b, stats_b, avg);
    else
        psd_kernel<double><<<(unsigned)nb, threads, smem, st>>>(iq_b, frame_stride, n_fft, log2n, window, tw, bp, power_b, stats_b, avg);
    TB_LAUNCH_CHECK(h);
    }
    if (smoothed) {
        const long nb_groups = (nb + group - 1) / group;
        dim3 grid((n_fft + 127) / 128, (unsigned)nb_groups);
        psd_ema_kernel<<<grid, 128, 0, st>>>(power, n_frames, n_fft, alpha, first, state_in, ema_state, smoothed, group, warm, f0, f0 + nb);
        TB_LAUNCH_CHECK(h);
    }
    }   // batches
    return TB200_OK;
}

}  // namespace tb200
