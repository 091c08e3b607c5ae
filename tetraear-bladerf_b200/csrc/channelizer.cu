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
// The DFT is a two-stage Cooley-Tukey M = N1 x N2; for M = 800 = 25 x 32 both small transforms run in registers (5 x 5
// Cooley-Tukey and a radix-2 FFT, refchan.cu), otherwise as direct small DFTs from shared-memory twiddle tables.  A CTA
// handles 8 consecutive frames so that each channel row receives one 64-byte segment.  The kernel is shared with the
// reference-mode channel bank (launch_polyphase_bank in refchan.cu).
#include <cmath>
#include <vector>
#include "common.h"
#include "kernels.h"

namespace tb200 {

constexpr int CH_P = 6;          // prototype taps per polyphase branch

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
    // same kernel as the reference-mode bank (refchan.cu): causal prototype (l0 = 0), frames M / 2 samples apart; the rotation
    // e^{-j 2 pi k (M/2) m / M} is the (-1)^(k m) of the defining sum
    rc = launch_polyphase_bank(h, x, n, M, N1, N2, CH_P, 0, M / 2, n_frames, proto_dev, tw_dev, out, out_stride, st);
    if (rc) return rc;
    return TB200_OK;
}

}  // namespace tb200
