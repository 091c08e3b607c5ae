/*
 * tetra_oracle.c — CPU restatement of the sequential loops of the reference's IQ->bursts
 * path.  TEST INFRASTRUCTURE ONLY: loaded by oracle/oracle.py, which only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import.
 * The product (tetraear-bladerf_b200/) never links or calls anything in here.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC  (oracle/build.py).
 * -ffp-contract=off matters: scipy's wheels are built for baseline x86-64 (no FMA), and
 * the float32 decimator is only reproducible bit-for-bit with separate mul/add roundings.
 *
 * Every function cites the reference (or scipy 1.18.1) code it follows.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------
 * scipy.signal._sosfilt._sosfilt (Cython, direct-form II transposed cascade), called by
 * scipy.signal.decimate <- /root/reference/tetraear/signal/processor.py:254.
 *   x_new = b0*x_cur + z0 ; z0 = b1*x_cur - a1*x_new + z1 ; z1 = b2*x_cur - a2*x_new
 * In place on x (contiguous), state zi[nsec][2] updated.
 * ---------------------------------------------------------------------------------- */
#define DEF_SOSFILT(NAME, T)                                                              \
    void NAME(const T *sos, int nsec, T *zi, T *x, long n)                                \
    {                                                                                     \
        for (long i = 0; i < n; ++i) {                                                    \
            T xc = x[i];                                                                  \
            for (int s = 0; s < nsec; ++s) {                                              \
                const T *c = sos + 6 * s;                                                 \
                T xn = c[0] * xc + zi[2 * s];                                             \
                zi[2 * s] = (c[1] * xc - c[4] * xn) + zi[2 * s + 1];                      \
                zi[2 * s + 1] = c[2] * xc - c[5] * xn;                                    \
                xc = xn;                                                                  \
            }                                                                             \
            x[i] = xc;                                                                    \
        }                                                                                 \
    }
DEF_SOSFILT(orc_sosfilt_f32, float)
DEF_SOSFILT(orc_sosfilt_f64, double)

/* ------------------------------------------------------------------------------------
 * scipy.signal.sosfiltfilt (padtype='odd', padlen=3*(2*nsec+1)) on ONE real lane, then
 * y[::q] (scipy.signal.decimate, zero_phase=True).  x is read with a stride so the I and
 * Q lanes of interleaved complex64 can be filtered independently — the complex filter
 * with real coefficients is bit-identical to two real ones.
 *   ext = [2 x0 - x[edge..1], x, 2 x[-1] - x[-2..-edge-1]]   (odd_ext)
 *   fwd : sosfilt(ext, zi * ext[0]) ; bwd : sosfilt(reverse(y), zi * y[-1]) ; reverse ; trim
 * out receives ceil(n/q) samples with stride out_stride.  Returns 0, or -1 if n <= edge
 * (scipy raises ValueError there; processor.py:256 then skips the decimation).
 * ---------------------------------------------------------------------------------- */
#define DEF_SOSFILTFILT(NAME, SOSFILT, T)                                                 \
    int NAME(const T *sos, int nsec, const T *zi, const T *x, long n, long xstride,       \
             int q, T *out, long out_stride)                                              \
    {                                                                                     \
        long edge = 3 * (2 * nsec + 1);                                                   \
        if (n <= edge) return -1;                                                         \
        long m = n + 2 * edge;                                                            \
        T *ext = (T *)malloc(sizeof(T) * (size_t)m);                                      \
        T *st = (T *)malloc(sizeof(T) * 2 * (size_t)nsec);                                \
        T x0 = x[0], xl = x[(n - 1) * xstride];                                           \
        for (long i = 0; i < edge; ++i) ext[i] = (T)2 * x0 - x[(edge - i) * xstride];     \
        for (long i = 0; i < n; ++i) ext[edge + i] = x[i * xstride];                      \
        for (long i = 0; i < edge; ++i) ext[edge + n + i] = (T)2 * xl - x[(n - 2 - i) * xstride]; \
        for (int i = 0; i < 2 * nsec; ++i) st[i] = zi[i] * ext[0];                        \
        SOSFILT(sos, nsec, st, ext, m);                                                   \
        for (long i = 0; i < m / 2; ++i) { T t = ext[i]; ext[i] = ext[m - 1 - i]; ext[m - 1 - i] = t; } \
        for (int i = 0; i < 2 * nsec; ++i) st[i] = zi[i] * ext[0];                        \
        SOSFILT(sos, nsec, st, ext, m);                                                   \
        /* ext is now the reversed result: sample k of the trimmed output is ext[m-1-edge-k] */ \
        long k = 0;                                                                       \
        for (long i = 0; i < n; i += q, ++k) out[k * out_stride] = ext[m - 1 - edge - i]; \
        free(ext); free(st);                                                              \
        return 0;                                                                         \
    }
DEF_SOSFILTFILT(orc_decimate_lane_f32, orc_sosfilt_f32, float)
DEF_SOSFILTFILT(orc_decimate_lane_f64, orc_sosfilt_f64, double)

/* ------------------------------------------------------------------------------------
 * scipy.signal.lfilter (C: scipy/signal/_lfilter.c.in, CDOUBLE_filt with real b, a and
 * a[0]==1 reduces per component to the real DOUBLE_filt):
 *   y = Z[0] + b[0] x ; Z[k] = Z[k+1] + b[k+1] x ; Z[k] -= a[k+1] y ; Z[last] = b[nb-1] x - a[nb-1] y
 * ---------------------------------------------------------------------------------- */
static void lfilter_f64(const double *b, const double *a, int nb, double *Z, double *x, long n)
{
    for (long i = 0; i < n; ++i) {
        double xi = x[i];
        double y = Z[0] + b[0] * xi;
        for (int k = 0; k < nb - 2; ++k) {
            double t = Z[k + 1] + b[k + 1] * xi;
            Z[k] = t - a[k + 1] * y;
        }
        Z[nb - 2] = b[nb - 1] * xi - a[nb - 1] * y;
        x[i] = y;
    }
}

/* scipy.signal.filtfilt(b, a, x) (method='pad', odd, padlen = 3*max(len a, len b)) on one
 * real lane <- processor.py:79.  ext_f32 != 0 reproduces the case freq_offset == 0, where
 * the complex64 decimator output reaches filtfilt un-promoted and odd_ext rounds the
 * extension samples to float32 before lfilter promotes to double.
 * Returns -1 when n <= padlen (scipy raises; processor.py:81-83 returns the input). */
int orc_filtfilt_lane_f64(const double *b, const double *a, int nb, const double *zi,
                          const double *x, long n, long xstride, int ext_f32,
                          double *out, long out_stride)
{
    long edge = 3L * nb;
    if (n <= edge) return -1;
    long m = n + 2 * edge;
    double *ext = (double *)malloc(sizeof(double) * (size_t)m);
    double Z[16];
    double x0 = x[0], xl = x[(n - 1) * xstride];
    for (long i = 0; i < edge; ++i) {
        double l = x[(edge - i) * xstride], r = x[(n - 2 - i) * xstride];
        if (ext_f32) {
            ext[i] = (double)((float)2 * (float)x0 - (float)l);
            ext[edge + n + i] = (double)((float)2 * (float)xl - (float)r);
        } else {
            ext[i] = 2 * x0 - l;
            ext[edge + n + i] = 2 * xl - r;
        }
    }
    for (long i = 0; i < n; ++i) ext[edge + i] = x[i * xstride];
    for (int k = 0; k < nb - 1; ++k) Z[k] = zi[k] * ext[0];
    lfilter_f64(b, a, nb, Z, ext, m);
    for (long i = 0; i < m / 2; ++i) { double t = ext[i]; ext[i] = ext[m - 1 - i]; ext[m - 1 - i] = t; }
    for (int k = 0; k < nb - 1; ++k) Z[k] = zi[k] * ext[0];
    lfilter_f64(b, a, nb, Z, ext, m);
    for (long i = 0; i < n; ++i) out[i * out_stride] = ext[m - 1 - edge - i];
    free(ext);
    return 0;
}

/* ------------------------------------------------------------------------------------
 * SignalProcessor.demodulate_dqpsk slicer loop, processor.py:131-164.  s = normalised
 * complex128 symbols (interleaved re,im), n >= 2; out gets n-1 dibits.
 *   diff = s[k] * conj(s[k-1]) ; phi = arctan2(im, re) ; thresholds +-3pi/8, +-5pi/8
 * ---------------------------------------------------------------------------------- */
void orc_slicer(const double *s, long n, uint8_t *out)
{
    const double PI = 3.141592653589793;
    const double t0 = -5 * PI / 8, t1 = -3 * PI / 8, t2 = 3 * PI / 8, t3 = 5 * PI / 8;
    for (long k = 1; k < n; ++k) {
        double ar = s[2 * k], ai = s[2 * k + 1];
        double br = s[2 * k - 2], bi = -s[2 * k - 1];
        double re = ar * br - ai * bi;
        double im = ar * bi + ai * br;
        double ph = atan2(im, re);
        uint8_t sym;
        if (ph < t0) sym = 3;
        else if (ph < t1) sym = 2;
        else if (ph < t2) sym = 0;
        else if (ph < t3) sym = 1;
        else sym = 3;
        out[k - 1] = sym;
    }
}

/* ------------------------------------------------------------------------------------
 * TetraDecoder.find_sync, /root/reference/tetraear/core/decoder.py:171-295.
 * TS1/TS2 are the 22-bit patterns at decoder.py:197-198.  thr is compared as
 * match/22 >= thr in double, exactly like the Python (`correlation >= threshold`).
 * pos_out must hold n entries (upper bound).  Returns the number of sync positions;
 * *max_corr_out = max over everything visited (skipped positions are NOT visited).
 * ---------------------------------------------------------------------------------- */
static const uint8_t TS1[22] = {1,1,0,1,0,0,0,0,1,1,1,0,1,0,0,1,1,1,0,1,0,0};
static const uint8_t TS2[22] = {0,1,1,1,1,0,1,0,0,1,0,0,0,0,1,1,0,1,1,1,0,0};

long orc_find_sync(const uint8_t *bits, long n, double thr, int32_t *pos_out, double *max_corr_out)
{
    long nfound = 0;
    double max_corr = 0.0;
    *max_corr_out = 0.0;
    if (n < 22) return 0;
    long nwin = n - 22 + 1;
    /* all_correlations (decoder.py:229,251-252): (pos, best corr at pos) for visited pos with corr>0 */
    int32_t *vis_pos = (int32_t *)malloc(sizeof(int32_t) * (size_t)nwin);
    uint8_t *vis_m = (uint8_t *)malloc((size_t)nwin);
    long nvis = 0;
    long i = 0;
    while (i < nwin) {
        int m1 = 0, m2 = 0;
        for (int k = 0; k < 22; ++k) { m1 += bits[i + k] == TS1[k]; }
        double c1 = (double)m1 / 22.0;
        int found = 0, best = m1;
        if (c1 > max_corr) max_corr = c1;
        if (c1 >= thr) { found = 1; }
        else {
            for (int k = 0; k < 22; ++k) { m2 += bits[i + k] == TS2[k]; }
            double c2 = (double)m2 / 22.0;
            if (m2 > best) best = m2;
            if (c2 > max_corr) max_corr = c2;
            if (c2 >= thr) found = 1;
        }
        if (best > 0) { vis_pos[nvis] = (int32_t)i; vis_m[nvis] = (uint8_t)best; ++nvis; }
        if (found) { pos_out[nfound++] = (int32_t)i; i += 250; continue; }
        i += 1;
    }
    /* adaptive re-pick, decoder.py:265-281 */
    if (nfound == 0 && max_corr > 0.75 && max_corr >= (thr - 0.15)) {
        double ad = max_corr - 0.02;
        if (ad < 0.75) ad = 0.75;
        if (ad < thr) {
            long seen_until = -1; /* positions [pos-250, pos+250) get marked; scan is ascending */
            for (long v = 0; v < nvis; ++v) {
                double c = (double)vis_m[v] / 22.0;
                long p = vis_pos[v];
                if (c >= ad && p >= seen_until) {
                    pos_out[nfound++] = (int32_t)p;
                    long hi = p + 250; if (hi > nwin) hi = nwin;
                    seen_until = hi;
                }
            }
        }
    }
    free(vis_pos); free(vis_m);
    *max_corr_out = max_corr;
    return nfound;
}

/* ------------------------------------------------------------------------------------
 * TetraProtocolParser._calculate_crc16, /root/reference/tetraear/core/protocol.py:331-347:
 * CRC-16-CCITT, poly 0x1021, init 0xFFFF, MSB-first bit feed, no final XOR.
 * ---------------------------------------------------------------------------------- */
uint16_t orc_crc16(const uint8_t *bits, long n, long stride)
{
    uint32_t crc = 0xFFFF;
    for (long i = 0; i < n; ++i) {
        crc ^= ((uint32_t)(bits[i * stride] & 1)) << 15;
        if (crc & 0x8000) crc = (crc << 1) ^ 0x1021; else crc <<= 1;
        crc &= 0xFFFF;
    }
    return (uint16_t)crc;
}

/* protocol.py:292-329 _check_crc on nb data bits */
static int check_crc(const uint8_t *d, long nb)
{
    if (nb < 16) return 0;
    long ones = 0;
    for (long i = 0; i < nb; ++i) ones += d[i];
    if (ones == 0 || ones == nb) return 0;
    uint16_t rx = 0;
    for (int i = 0; i < 16; ++i) rx = (uint16_t)((rx << 1) | d[nb - 16 + i]);
    uint16_t c = orc_crc16(d, nb - 16, 1);
    if (__builtin_popcount((unsigned)(c ^ rx)) <= 2) return 1;
    c = orc_crc16(d + nb - 17, nb - 16, -1); /* payload[::-1] */
    if (__builtin_popcount((unsigned)(c ^ rx)) <= 2) return 1;
    return 0;
}

/* ------------------------------------------------------------------------------------
 * TetraProtocolParser.parse_burst integer path, protocol.py:192-290, over a batch:
 * dibits[nb][255] -> burst_type (5 = Synchronization, 2 = NormalDownlink, BurstType enum
 * protocol.py:34-41), crc_ok, data_bits (216 used for normal, 510 for sync; row pitch 510),
 * n_data.  Patterns: protocol.py:162-163.
 * ---------------------------------------------------------------------------------- */
static const uint8_t SYNC_CONT[22] = {1,1,0,1,0,0,0,0,1,1,1,0,1,0,0,1,1,1,0,1,0,0};
static const uint8_t SYNC_DISC[22] = {0,0,1,1,1,0,1,0,0,1,0,0,0,0,1,1,0,1,0,0,1,1};

void orc_parse_bursts(const uint8_t *dibits, long nbursts, uint8_t *burst_type,
                      uint8_t *crc_ok, uint8_t *data_bits, int32_t *n_data)
{
    uint8_t bits[510];
    for (long b = 0; b < nbursts; ++b) {
        const uint8_t *s = dibits + 255 * b;
        for (int k = 0; k < 255; ++k) { bits[2 * k] = (s[k] >> 1) & 1; bits[2 * k + 1] = s[k] & 1; }
        int mc = 0, md = 0;
        for (int k = 0; k < 22; ++k) { mc += bits[255 + k] == SYNC_CONT[k]; md += bits[255 + k] == SYNC_DISC[k]; }
        int mx = mc > md ? mc : md;
        int is_sync = ((double)mx / 22.0) > 0.8;
        uint8_t *d = data_bits + 510 * b;
        long nd;
        if (is_sync) { memcpy(d, bits, 510); nd = 510; }
        else { memcpy(d, bits, 108); memcpy(d + 108, bits + 122, 108); memset(d + 216, 0, 510 - 216); nd = 216; }
        burst_type[b] = is_sync ? 5 : 2;
        n_data[b] = (int32_t)nd;
        crc_ok[b] = (uint8_t)check_crc(d, nd);
    }
}
