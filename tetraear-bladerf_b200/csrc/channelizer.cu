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

__global__ void __launch_bounds__(256)
channelizer_kernel(const float2 *__restrict__ x, long n, int M, int N1, int N2, int F, long n_frames,
                   const float *__restrict__ proto, const float2 *__restrict__ tw, float2 *__restrict__ out,
                   long out_stride)
{
    extern __shared__ float2 ch_smem[];
    float2 *v = ch_smem;                      // [F][M]
    float2 *a = v + (long)F * M;              // [F][M]  (k1-major: a[f][k1*N2 + r2])
    float2 *w1 = a + (long)F * M;             // [N1] exp(+j 2 pi i / N1)
    float2 *w2 = w1 + N1;                     // [N2]
    float2 *wm = w2 + N2;                     // [M]  exp(+j 2 pi i / M)
    const int D = M / 2;
    const long m0 = (long)blockIdx.x * F;
    for (int i = threadIdx.x; i < N1 + N2 + M; i += blockDim.x) w1[i] = tw[i];
    // ---- polyphase FIR
    for (int i = threadIdx.x; i < F * M; i += blockDim.x) {
        int f = i / M, r = i - f * M;
        long m = m0 + f;
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
        v[i] = acc;
    }
    __syncthreads();
    // ---- stage A: N2 DFTs of size N1 over r1 (r = N2 r1 + r2), then twiddle exp(+j 2 pi k1 r2 / M)
    for (int i = threadIdx.x; i < F * M; i += blockDim.x) {
        int f = i / M, j = i - f * M;
        int k1 = j / N2, r2 = j - k1 * N2;
        const float2 *vf = v + (long)f * M;
        float2 acc = make_float2(0.f, 0.f);
        int e = 0;                                        // (k1 r1) mod N1
        for (int r1 = 0; r1 < N1; ++r1) {
            float2 s = vf[N2 * r1 + r2], w = w1[e];
            acc.x += s.x * w.x - s.y * w.y; acc.y += s.x * w.y + s.y * w.x;
            e += k1; if (e >= N1) e -= N1;
        }
        float2 t = wm[(k1 * r2) % M];
        a[i] = make_float2(acc.x * t.x - acc.y * t.y, acc.x * t.y + acc.y * t.x);
    }
    __syncthreads();
    // ---- stage B: N1 DFTs of size N2 over r2; k = k1 + N1 k2; sign (-1)^(k m); store [k][m]
    for (int i = threadIdx.x; i < F * M; i += blockDim.x) {
        int k = i / F, f = i - k * F;                     // f fastest: consecutive threads -> consecutive m of one channel
        long m = m0 + f;
        if (m >= n_frames) continue;
        int k1 = k % N1, k2 = k / N1;
        const float2 *af = a + (long)f * M + k1 * N2;
        float2 acc = make_float2(0.f, 0.f);
        int e = 0;
        for (int r2 = 0; r2 < N2; ++r2) {
            float2 s = af[r2], w = w2[e];
            acc.x += s.x * w.x - s.y * w.y; acc.y += s.x * w.y + s.y * w.x;
            e += k2; if (e >= N2) e -= N2;
        }
        if ((k & 1) && (m & 1)) { acc.x = -acc.x; acc.y = -acc.y; }
        out[(long)k * out_stride + m] = acc;
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

int channelizer_factor(int M, int *N1, int *N2)
{
    int best = 1;
    for (int a = 1; a * a <= M; ++a) if (M % a == 0) best = a;
    *N1 = best; *N2 = M / best;
    return (*N2 <= 128) ? TB200_OK : fail(TB200_E_ARG, "channelizer: M=%d has no factorisation with both factors <= 128", M);
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
    size_t smem = sizeof(float2) * ((size_t)2 * F * M + N1 + N2 + M);
    while (smem > 200 * 1024 && F > 1) { F /= 2; smem = sizeof(float2) * ((size_t)2 * F * M + N1 + N2 + M); }
    if (smem > 200 * 1024) return fail(TB200_E_ARG, "channelizer: M=%d too large for shared memory", M);
    cudaError_t e = cudaFuncSetAttribute(channelizer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(TB200_E_CUDA, "cudaFuncSetAttribute(channelizer): %s", cudaGetErrorString(e));
    unsigned blocks = (unsigned)((n_frames + F - 1) / F);
    channelizer_kernel<<<blocks, 256, smem, st>>>(x, n, M, N1, N2, F, n_frames, proto_dev, tw_dev, out, out_stride);
    TB_LAUNCH_CHECK(h);
    return TB200_OK;
}

}  // namespace tb200
