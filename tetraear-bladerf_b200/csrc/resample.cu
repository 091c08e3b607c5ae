// resample.cu — SignalProcessor.resample (/root/reference/tetraear/signal/processor.py:35-49), i.e.
// scipy.signal.resample(x, num) for complex input: X = fft(x); keep the m = min(num, n) lowest-frequency
// bins (with scipy's rule for the unpaired bin m/2); y = ifft(Y) * num / n.
// No caller on the receive path uses it, so this is a plain O(n m) direct DFT in
// float64 from a twiddle table — exact index arithmetic (k t mod n), any length, no FFT-size limits.
#include "common.h"
#include "kernels.h"

namespace tb200 {

__global__ void rs_table_kernel(long n, double sign, double2 *__restrict__ w)
{
    long r = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (r >= n) return;
    double s, c;
    sincospi(sign * 2.0 * (double)r / (double)n, &s, &c);
    w[r] = make_double2(c, s);
}

// forward DFT of the kept bins: bin index b in [0, m) -> frequency k = b (b < m2) or n - (m - b)
template <typename TI>
__global__ void rs_forward_kernel(const TI *__restrict__ x, long n, long m, long m2, const double2 *__restrict__ w,
                                  double2 *__restrict__ X)
{
    long b = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (b >= m) return;
    long k = (b < m2) ? b : n - (m - b);
    double ar = 0.0, ai = 0.0;
    long e = 0;
    for (long t = 0; t < n; ++t) {
        double xr = (double)x[2 * t], xi = (double)x[2 * t + 1];
        double2 ww = w[e];
        ar += xr * ww.x - xi * ww.y;
        ai += xr * ww.y + xi * ww.x;
        e += k; if (e >= n) e -= n;
    }
    X[b] = make_double2(ar, ai);
}

// y[t] = (1/n) sum over kept bins of Y e^{+2 pi i j t / num}; bin b sits at output frequency j(b)
template <typename TO>
__global__ void rs_inverse_kernel(const double2 *__restrict__ X, long n, long num, long m, long m2,
                                  const double2 *__restrict__ w2, TO *__restrict__ y)
{
    long t = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (t >= num) return;
    double ar = 0.0, ai = 0.0;
    const bool even = (m % 2) == 0;
    for (long b = 0; b < m; ++b) {
        long j = (b < m2) ? b : num - (m - b);
        double2 v = X[b];
        long e = (long)(((__int128)j * t) % num);
        double2 ww = w2[e];
        ar += v.x * ww.x - v.y * ww.y;
        ai += v.x * ww.y + v.y * ww.x;
    }
    if (even && m >= 2) {
        // scipy's unpaired-bin rule (two-sided path).  Kept bins: [0, m2) and the last m - m2 of X.
        // down-sampling: Y[-m/2] += X[-m/2]  -> the bin at output frequency num - m/2 gets X[n - m/2] added; but in the
        //   copy above Y[num - m/2] is NOT among the negative bins when m2 - m = -(m/2 - 1), so it is the +m/2 bin
        //   (index m/2 == num - m/2 when num == m): handled below as an extra term.
        // up-sampling: Y[m/2] /= 2 and Y[num - m/2] = Y[m/2].
        long h = m / 2;
        if (num < n) {
            // extra term X[n - h] at output frequency (num - h) mod num == h (since num == m)
            // (computed by the caller into X[m] to keep this kernel branch-free in the loop)
            double2 v = X[m];
            long e = (long)(((__int128)((num - h) % num) * t) % num);
            double2 ww = w2[e];
            ar += v.x * ww.x - v.y * ww.y;
            ai += v.x * ww.y + v.y * ww.x;
        } else if (n < num) {
            // undo half of the +h bin and mirror it to num - h
            double2 v = X[h];
            long e1 = (long)(((__int128)h * t) % num), e2 = (long)(((__int128)(num - h) * t) % num);
            double2 wa = w2[e1], wb = w2[e2];
            ar += -0.5 * (v.x * wa.x - v.y * wa.y) + 0.5 * (v.x * wb.x - v.y * wb.y);
            ai += -0.5 * (v.x * wa.y + v.y * wa.x) + 0.5 * (v.x * wb.y + v.y * wb.x);
        }
    }
    const double sc = 1.0 / (double)n;
    y[2 * t] = (TO)(ar * sc);
    y[2 * t + 1] = (TO)(ai * sc);
}

// one extra forward bin (frequency k) -> X[slot]
template <typename TI>
__global__ void rs_single_bin_kernel(const TI *__restrict__ x, long n, long k, const double2 *__restrict__ w, double2 *__restrict__ X,
                                     long slot)
{
    __shared__ double sr[256], si[256];
    double ar = 0.0, ai = 0.0;
    for (long t = threadIdx.x; t < n; t += blockDim.x) {
        long e = (long)(((__int128)k * t) % n);
        double xr = (double)x[2 * t], xi = (double)x[2 * t + 1];
        double2 ww = w[e];
        ar += xr * ww.x - xi * ww.y;
        ai += xr * ww.y + xi * ww.x;
    }
    sr[threadIdx.x] = ar; si[threadIdx.x] = ai;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) { sr[threadIdx.x] += sr[threadIdx.x + o]; si[threadIdx.x] += si[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) X[slot] = make_double2(sr[0], si[0]);
}

int launch_resample(tb200_handle h, const void *x, int dtype, long n, long num, void *y, double *scratch, size_t scratch_bytes,
                    cudaStream_t st)
{
    if (n <= 0 || num <= 0) return TB200_OK;
    long m = num < n ? num : n, m2 = m / 2 + 1;
    size_t need = sizeof(double2) * (size_t)(n + num + m + 2);
    if (need > scratch_bytes) return fail(TB200_E_ARG, "tb200_resample: %ld -> %ld samples exceed the handle capacity", n, num);
    double2 *w = reinterpret_cast<double2 *>(scratch), *w2 = w + n, *X = w2 + num;
    rs_table_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, -1.0, w); TB_LAUNCH_CHECK(h);
    rs_table_kernel<<<(unsigned)((num + 255) / 256), 256, 0, st>>>(num, 1.0, w2); TB_LAUNCH_CHECK(h);
    unsigned fb = (unsigned)((m + 127) / 128);
    if (dtype == TB200_DTYPE_C64) rs_forward_kernel<float><<<fb, 128, 0, st>>>((const float *)x, n, m, m2, w, X);
    else rs_forward_kernel<double><<<fb, 128, 0, st>>>((const double *)x, n, m, m2, w, X);
    TB_LAUNCH_CHECK(h);
    if (m % 2 == 0 && num < n) {             // X[-m/2] of the input spectrum, needed by the down-sampling rule
        long k = n - m / 2;
        if (dtype == TB200_DTYPE_C64) rs_single_bin_kernel<float><<<1, 256, 0, st>>>((const float *)x, n, k, w, X, m);
        else rs_single_bin_kernel<double><<<1, 256, 0, st>>>((const double *)x, n, k, w, X, m);
        TB_LAUNCH_CHECK(h);
    }
    unsigned ib = (unsigned)((num + 127) / 128);
    if (dtype == TB200_DTYPE_C64) rs_inverse_kernel<float><<<ib, 128, 0, st>>>(X, n, num, m, m2, w2, (float *)y);
    else rs_inverse_kernel<double><<<ib, 128, 0, st>>>(X, n, num, m, m2, w2, (double *)y);
    TB_LAUNCH_CHECK(h);
    return TB200_OK;
}

}  // namespace tb200
