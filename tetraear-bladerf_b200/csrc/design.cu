// design.cu — host-side filter design of libtetra_b200 (no device code).
//
// The reference designs nothing itself: it calls scipy.signal.decimate(x, q)
// (/root/reference/tetraear/signal/processor.py:254) and butter(4, wn) + filtfilt
// (processor.py:78-79).  To reproduce its float32 decimator bit for bit the coefficients
// must be the float32 roundings of scipy's float64 design, so the evaluation order of
// scipy 1.18.1 / numpy 2.3.5 is followed operation by operation, including numpy's complex
// quotient (Smith's method with a reciprocal, two roundings for a real divisor).
#include <cmath>
#include <complex>
#include <algorithm>
#include <vector>
#include "common.h"

namespace {

typedef std::complex<double> cd;

inline cd cmul(cd a, cd b)
{
    return cd(a.real() * b.real() - a.imag() * b.imag(), a.real() * b.imag() + a.imag() * b.real());
}

// numpy loops.c.src @TYPE@_divide
template <typename T>
inline std::complex<T> np_cdiv(std::complex<T> a, std::complex<T> b)
{
    T ar = a.real(), ai = a.imag(), br = b.real(), bi = b.imag();
    if (std::fabs(br) >= std::fabs(bi)) {
        T rat = bi / br;
        T scl = T(1) / (br + bi * rat);
        return std::complex<T>((ar + ai * rat) * scl, (ai - ar * rat) * scl);
    }
    T rat = br / bi;
    T scl = T(1) / (bi + br * rat);
    return std::complex<T>((ar * rat + ai) * scl, (ai * rat - ar) * scl);
}

// lp2lp_zpk(wo = 4 tan(pi wn / 2)) then bilinear_zpk(fs = 2) for an all-pole prototype
void analog_to_digital(std::vector<cd> &p, double &k, double wn)
{
    const double fs = 2.0;
    double warped = 2 * fs * std::tan(M_PI * wn / fs);
    int degree = (int)p.size();
    for (auto &v : p) v = cd(warped * v.real(), warped * v.imag());
    k = k * std::pow(warped, (double)degree);
    const double fs2 = 2.0 * fs;
    cd prod(1.0, 0.0);
    for (int i = 0; i < degree; ++i) {
        cd den = cd(fs2 - p[i].real(), -p[i].imag());
        prod = (i == 0) ? den : cmul(prod, den);
    }
    for (auto &v : p) v = np_cdiv<double>(cd(fs2 + v.real(), v.imag()), cd(fs2 - v.real(), -v.imag()));
    k = k * np_cdiv<double>(cd(1.0, 0.0), prod).real();
}

// scipy.signal.lfilter_zi in the arithmetic of numpy's complex type with zero imaginary
// parts: a quotient by a real number is a multiplication by its rounded reciprocal.
template <typename T>
void lfilter_zi_t(const T *b, const T *a, int n, T *zi)
{
    T sb = b[0], sa = a[0];
    for (int i = 1; i < n; ++i) { sb = sb + b[i]; sa = sa + a[i]; }
    T y_inf = sb * (T(1) / sa);
    T acc = T(0);
    for (int i = n - 1; i >= 1; --i) {
        T t = b[i] - y_inf * a[i];
        acc = (i == n - 1) ? t : acc + t;
        zi[i - 1] = acc;
    }
}

template <typename T>
void sosfilt_zi_t(const T *sos, int nsec, T *zi)
{
    T scale = T(1);
    for (int s = 0; s < nsec; ++s) {
        const T *b = sos + 6 * s, *a = sos + 6 * s + 3;
        T z[2];
        lfilter_zi_t<T>(b, a, 3, z);
        zi[2 * s] = scale * z[0];
        zi[2 * s + 1] = scale * z[1];
        T sb = (b[0] + b[1]) + b[2], sa = (a[0] + a[1]) + a[2];
        scale = scale * (sb * (T(1) / sa));
    }
}

}  // namespace

extern "C" int tb200_design_decimator(int q, double sos64[24], double zi64[8], float sos32[24], float zi32[8])
{
    if (q < 2) return tb200::fail(TB200_E_ARG, "tb200_design_decimator: q must be >= 2");
    const int N = 8;
    const double rp = 0.05;
    // cheb1ap
    double eps = std::sqrt(std::pow(10.0, 0.1 * rp) - 1.0);
    double mu = 1.0 / N * std::asinh(1 / eps);
    std::vector<cd> p(N);
    cd prod;
    for (int i = 0; i < N; ++i) {
        int m = -N + 1 + 2 * i;
        double theta = M_PI * m / (2 * N);
        p[i] = -std::sinh(cd(mu, theta));
        prod = (i == 0) ? -p[i] : cmul(prod, -p[i]);
    }
    double k = prod.real();
    k = k / std::sqrt(1 + eps * eps);
    analog_to_digital(p, k, 0.8 / q);
    // zpk2sos('nearest'): pole pair nearest the unit circle last, gain in the first section
    std::vector<cd> pos;
    for (auto &v : p) if (v.imag() > 0) pos.push_back(v);
    if ((int)pos.size() != N / 2) return tb200::fail(TB200_E_ARG, "tb200_design_decimator: unexpected real pole");
    std::stable_sort(pos.begin(), pos.end(), [](const cd &a, const cd &b) {
        return std::fabs(1 - std::hypot(a.real(), a.imag())) > std::fabs(1 - std::hypot(b.real(), b.imag()));
    });
    double sos[24];
    for (int s = 0; s < 4; ++s) {
        double pr = pos[s].real(), pi = pos[s].imag();
        sos[6 * s + 0] = 1.0; sos[6 * s + 1] = 2.0; sos[6 * s + 2] = 1.0;
        sos[6 * s + 3] = 1.0;
        sos[6 * s + 4] = -pr - pr;
        sos[6 * s + 5] = pr * pr + pi * pi;
    }
    sos[0] *= k; sos[1] *= k; sos[2] *= k;
    float s32[24];
    for (int i = 0; i < 24; ++i) s32[i] = (float)sos[i];
    if (sos64) for (int i = 0; i < 24; ++i) sos64[i] = sos[i];
    if (sos32) for (int i = 0; i < 24; ++i) sos32[i] = s32[i];
    if (zi64) sosfilt_zi_t<double>(sos, 4, zi64);
    if (zi32) sosfilt_zi_t<float>(s32, 4, zi32);
    return TB200_OK;
}

namespace tb200 {
// digital poles and gain of butter(4, clip((bw/2)/(fs/2))) (zeros: four at -1)
int channel_filter_poles(double bandwidth, double fs, std::complex<double> poles[4], double *gain)
{
    if (!(fs > 0)) return fail(TB200_E_ARG, "channel filter: fs must be > 0");
    double cutoff = (bandwidth / 2) / (fs / 2);
    cutoff = std::min(0.99, std::max(0.01, cutoff));
    const int N = 4;
    std::vector<cd> p(N);
    for (int i = 0; i < N; ++i) {
        int m = -N + 1 + 2 * i;
        // 1j*pi*m/(2N): numpy divides the complex by (8+0j) -> multiply by the reciprocal 0.125
        double theta = (M_PI * m) * (1.0 / (2 * N));
        p[i] = -std::exp(cd(0.0, theta));
    }
    double k = 1.0;
    analog_to_digital(p, k, cutoff);
    for (int i = 0; i < N; ++i) poles[i] = p[i];
    *gain = k;
    return TB200_OK;
}
}  // namespace tb200

extern "C" int tb200_design_channel_filter(double bandwidth, double fs, double b[5], double a[5], double zi[4])
{
    const int N = 4;
    std::vector<cd> p(N);
    double k = 1.0;
    {
        cd pz[4];
        int rc = tb200::channel_filter_poles(bandwidth, fs, pz, &k);
        if (rc) return rc;
        for (int i = 0; i < N; ++i) p[i] = pz[i];
    }
    // a = real(np.poly(p)): sequential convolution with [1, -root]
    cd acc[N + 1];
    acc[0] = cd(1.0, 0.0);
    int len = 1;
    for (int r = 0; r < N; ++r) {
        cd nxt[N + 1];
        for (int i = 0; i <= len; ++i) {
            cd v(0.0, 0.0);
            if (i < len) v = acc[i];
            if (i >= 1) v = (i < len) ? v + cmul(acc[i - 1], -p[r]) : cmul(acc[i - 1], -p[r]);
            nxt[i] = v;
        }
        ++len;
        for (int i = 0; i < len; ++i) acc[i] = nxt[i];
    }
    const double binom[5] = {1.0, 4.0, 6.0, 4.0, 1.0};
    double bb[5], aa[5];
    for (int i = 0; i < 5; ++i) { aa[i] = acc[i].real(); bb[i] = k * binom[i]; }
    if (b) for (int i = 0; i < 5; ++i) b[i] = bb[i];
    if (a) for (int i = 0; i < 5; ++i) a[i] = aa[i];
    if (zi) {
        // float64 lfilter_zi: real division here (numpy float64, not complex)
        double sb = bb[0], sa = aa[0];
        for (int i = 1; i < 5; ++i) { sb += bb[i]; sa += aa[i]; }
        double y_inf = sb / sa;
        double accum = 0.0;
        for (int i = 4; i >= 1; --i) {
            double t = bb[i] - y_inf * aa[i];
            accum = (i == 4) ? t : accum + t;
            zi[i - 1] = accum;
        }
    }
    return TB200_OK;
}

extern "C" int tb200_plan_process(double sample_rate, int chunk_len, tb200_plan *plan)
{
    if (!plan || chunk_len < 0 || !(sample_rate > 0)) return tb200::fail(TB200_E_ARG, "tb200_plan_process: bad argument");
    tb200_plan P;
    P.q = 1;
    P.rate = sample_rate;
    P.n_dec = chunk_len;
    const double target = 240000;
    if (sample_rate > target * 2) {
        int q = (int)(sample_rate / target);
        if (q > 1 && chunk_len > 27) {  // scipy raises for len <= padlen; processor.py:256 keeps going undecimated
            P.q = q;
            P.rate = sample_rate / q;
            P.n_dec = (chunk_len + q - 1) / q;
        }
    }
    P.filtered = P.n_dec > 15 ? 1 : 0;
    P.sps = (int)(P.rate / 18000);
    P.phase_step = std::max(1, P.sps / 8);
    P.max_symbols = (P.sps > 1) ? (P.n_dec / P.sps) : P.n_dec;
    *plan = P;
    return TB200_OK;
}
