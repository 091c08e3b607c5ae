// channelizer.cu — `etsi` mode polyphase FIR channelizer: one wideband complex64 stream at
// fs = M x 25 kHz  ->  M channels (centre k x 25 kHz) at 50 kHz each (2x oversampled, D = M/2).
//
// The reference has no channelizer (process() decimates ONE channel with an IIR;
// other channels need frequency_shift + process per channel, i.e. M passes over the input).  Parity
// unpinned: checked against the defining sum  y_k[m] = sum_n x[n] h[mD - n] exp(-j 2 pi k n / M)
// evaluated in numpy (tests/test_etsi_channelizer.py), and end to end through the etsi receiver.
//
// Per output frame m:  v[r] = sum_{p<P} h[pM + r] x[mD - pM - r]   (polyphase FIR, P taps per branch)
//                      Y[k] = sum_r v[r] exp(+j 2 pi k r / M)      (size-M inverse DFT)
//                      y_k[m] = (-1)^(k m) Y[k]
// The DFT is a two-stage Cooley-Tukey M = N1 x N2 with direct small DFTs from twiddle tables in
// shared memory (M(N1+N2) complex MACs per frame).  A CTA handles F consecutive frames so that each
// channel row receives F consecutive outputs (one 64-byte segment for F = 8).
#include <cmath>
#include <vector>
#include "common.h"
#include "kernels.h"

namespace tb200 {

constexpr int CH_P = 6;          // prototype taps per polyphase branch

// N2 = 2^LOG2_N2 (the power-of-two factor of M), N1 = M / N2.  `a` rows are padded to N2+1 so that lanes
// reading consecutive k1 rows in stage B hit different banks.
__global__ void __launch_bounds__(256)
channelizer_kernel(const float2 *__restrict__ x, long n, int M, int N1, int log2_n2, int F, long n_frames,
                   const float *__restrict__ proto, const float2 *__restrict__ tw, float2 *__restrict__ out,
                   long out_stride)
{
    extern __shared__ float2 ch_smem[];
    const int N2 = 1 << log2_n2;
    const int APITCH = N2 + 1;
    float2 *v = ch_smem;                      // [F][M]
    float2 *a = v + (long)F * M;              // [F][N1][APITCH]
    float2 *w1 = a + (long)F * N1 * APITCH;   // [N1] exp(+j 2 pi i / N1)
    float2 *w2 = w1 + N1;                     // [N2]
    float2 *wm = w2 + N2;                     // [M]  exp(+j 2 pi i / M)
    const int D = M / 2;
    const long m0 = (long)blockIdx.x * F;
    for (int i = threadIdx.x; i < N1 + N2 + M; i += blockDim.x) w1[i] = tw[i];
    // ---- polyphase FIR: v[f][r] = sum_p h[pM + r] x[(m0+f) D - pM - r]
    for (int f = 0; f < F; ++f) {
        const long m = m0 + f;
        for (int r = threadIdx.x; r < M; r += blockDim.x) {
            float2 acc = make_float2(0.f, 0.f);
            if (m < n_frames) {
#pragma unroll
                for (int p = 0; p < CH_P; ++p) {
                    long idx = m * D - (long)p * M - r;
                    if (idx >= 0 && idx < n) {
                        float2 s = x[idx];
                        float h = proto[p * M + r];
                        acc.x = fmaf(h, s.x, acc.x); acc.y = fmaf(h, s.y, acc.y);
                    }
                }
            }
            v[f * M + r] = acc;
        }
    }
    __syncthreads();
    // ---- stage A: for every r2, a size-N1 DFT over r1 (r = N2 r1 + r2), times exp(+j 2 pi k1 r2 / M)
    for (int j = threadIdx.x; j < M; j += blockDim.x) {
        const int k1 = j >> log2_n2, r2 = j & (N2 - 1);
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
    // ---- stage B: size-N2 DFT over r2; k = k1 + N1 k2; sign (-1)^(k m); store [k][m0 .. m0+F)
    for (int k = threadIdx.x; k < M; k += blockDim.x) {
        const int k2 = k / N1, k1 = k - k2 * N1;
        float2 *o = out + (long)k * out_stride + m0;
        for (int f = 0; f < F; ++f) {
            const long m = m0 + f;
            if (m >= n_frames) break;
            const float2 *af = a + ((long)f * N1 + k1) * APITCH;
            float2 acc = make_float2(0.f, 0.f);
            int e = 0;                                        // (k2 r2) mod N2
            for (int r2 = 0; r2 < N2; ++r2) {
                float2 s = af[r2], w = w2[e];
                acc.x = fmaf(s.x, w.x, fmaf(-s.y, w.y, acc.x)); acc.y = fmaf(s.x, w.y, fmaf(s.y, w.x, acc.y));
                e = (e + k2) & (N2 - 1);
            }
            if ((k & 1) && (m & 1)) { acc.x = -acc.x; acc.y = -acc.y; }
            o[f] = acc;
        }
    }
}

static double bessel_i0(double x)
{
    double s = 1.0, t = 1.0;
    for (int k = 1; k < 60; ++k) { t *= (x / (2.0 * k)) * (x / (2.0 * k)); s += t; if (t < 1e-18 * s) break; }
    return s;
}

// Kaiser-windowed sinc low-pass: cutoff 25 kHz of fs = M x 25 kHz (i.e. 1/M of fs), length P*M, unit DC gain
void channelizer_prototype(int M, std::vector<float> &h)
{
    const int L = CH_P * M;
    const double beta = 8.0, fc = 1.0 / M;               // cycles per sample
    std::vector<double> d(L);
    double sum = 0.0;
    for (int i = 0; i < L; ++i) {
        double t = i - (L - 1) / 2.0;
        double sinc = (std::fabs(t) < 1e-12) ? 2 * fc : std::sin(2 * M_PI * fc * t) / (M_PI * t);
        double r = 2.0 * i / (L - 1) - 1.0;
        double w = bessel_i0(beta * std::sqrt(std::max(0.0, 1 - r * r))) / bessel_i0(beta);
        d[i] = sinc * w;
        sum += d[i];
    }
    h.resize(L);
    for (int i = 0; i < L; ++i) h[i] = (float)(d[i] / sum);
}

// M = N1 * N2 with N2 the largest power of two dividing M (capped at 64)
int channelizer_factor(int M, int *N1, int *N2)
{
    int n2 = 1;
    while (n2 < 64 && M % (2 * n2) == 0) n2 *= 2;
    *N2 = n2; *N1 = M / n2;
    return (*N1 <= 256) ? TB200_OK : fail(TB200_E_ARG, "channelizer: M=%d = %d x %d, odd factor too large", M, *N1, n2);
}

int launch_channelizer(tb200_handle h, const float2 *x, long n, int M, float *proto_dev, float2 *tw_dev, float2 *out,
                       long out_stride, long n_frames, cudaStream_t st)
{
    int N1, N2;
    int rc = channelizer_factor(M, &N1, &N2);
    if (rc) return rc;
    if (h->chan_M != M) {
        std::vector<float> hp;
        channelizer_prototype(M, hp);
        std::vector<float2> tw(N1 + N2 + M);
        for (int i = 0; i < N1; ++i) tw[i] = make_float2((float)std::cos(2 * M_PI * i / N1), (float)std::sin(2 * M_PI * i / N1));
        for (int i = 0; i < N2; ++i) tw[N1 + i] = make_float2((float)std::cos(2 * M_PI * i / N2), (float)std::sin(2 * M_PI * i / N2));
        for (int i = 0; i < M; ++i) tw[N1 + N2 + i] = make_float2((float)std::cos(2 * M_PI * i / M), (float)std::sin(2 * M_PI * i / M));
        TB_CUDA(cudaMemcpyAsync(proto_dev, hp.data(), hp.size() * sizeof(float), cudaMemcpyHostToDevice, st));
        TB_CUDA(cudaMemcpyAsync(tw_dev, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice, st));
        TB_CUDA(cudaStreamSynchronize(st));            // host vectors go out of scope
        h->chan_M = M;
    }
    int F = 8;
    int log2_n2 = 0;
    while ((1 << log2_n2) < N2) ++log2_n2;
    auto need = [&](int f) { return sizeof(float2) * ((size_t)f * M + (size_t)f * N1 * (N2 + 1) + N1 + N2 + M); };
    size_t smem = need(F);
    while (smem > 100 * 1024 && F > 1) { F /= 2; smem = need(F); }
    if (smem > 200 * 1024) return fail(TB200_E_ARG, "channelizer: M=%d too large for shared memory", M);
    cudaError_t e = cudaFuncSetAttribute(channelizer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(TB200_E_CUDA, "cudaFuncSetAttribute(channelizer): %s", cudaGetErrorString(e));
    unsigned blocks = (unsigned)((n_frames + F - 1) / F);
    channelizer_kernel<<<blocks, 256, smem, st>>>(x, n, M, N1, log2_n2, F, n_frames, proto_dev, tw_dev, out, out_stride);
    TB_LAUNCH_CHECK(h);
    return TB200_OK;
}

}  // namespace tb200
