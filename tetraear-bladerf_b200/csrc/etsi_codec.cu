// etsi_codec.cu — `etsi` mode channel decoder over batches of bursts: descrambling, block
// de-interleaving (K, a), RCPC rate-2/3 de-puncturing, 16-state Viterbi, CRC-16
// (ETSI EN 300 392-2 clause 8).  The reference has no implementation of this stage
// (/root/reference/tetraear/core/protocol.py:300: "we do not do full channel decoding"), so
// parity is unpinned; the kernel is checked bit for bit against oracle/etsi_oracle.py and by
// encode -> decode round trips.
//
// Mapping: 16 trellis states = 16 lanes, two bursts per warp.  Path metrics never leave the
// registers: the add-compare-select of state ns reads its two predecessors ns>>1 and (ns>>1)|8
// with two shuffles; because the predecessors differ only in the oldest register bit, which
// enters every generator polynomial, their branch metrics are +bm and -bm.  Survivor decisions
// are one bit per (state, step) packed into K2/32 registers per lane; the traceback walks them
// with one shuffle per step.  A burst is K3 bytes in and 42 bytes out: the kernel is
// issue-bound (~20 warp instructions per trellis step), not HBM-bound.
#include "common.h"
#include "kernels.h"

namespace tb200 {

struct EtsiParams {
    uint32_t scr[14];      // scrambling sequence bits p(k), bit k%32 of word k/32
    int K1, K2, K3, a;
};

template <int K2>
__global__ void __launch_bounds__(128)
etsi_decode_kernel(const int8_t *__restrict__ soft, long n, EtsiParams P, uint8_t *__restrict__ bits_out,
                   uint8_t *__restrict__ crc_ok, int32_t *__restrict__ metric)
{
    constexpr int NW = (K2 + 31) / 32;
    __shared__ int8_t s4[8][432];
    const int hw = threadIdx.x >> 4;            // half-warp in block (0..7)
    const int l16 = threadIdx.x & 15;           // trellis state owned by this lane
    const int base = threadIdx.x & 16;          // first lane of this half-warp inside the warp
    const long burst = (long)blockIdx.x * 8 + hw;
    const bool live = burst < n;
    const int K3 = P.K3;
    // ---- load + descramble (sign flip where p(k) = 1)
    if (live) {
        const int8_t *src = soft + burst * K3;
        for (int k = l16; k < K3; k += 16) {
            int v = src[k];
            int pk = (P.scr[k >> 5] >> (k & 31)) & 1;
            s4[hw][k] = (int8_t)(pk ? -v : v);
        }
    }
    __syncwarp();
    // ---- forward pass
    const int ns = l16, b = ns & 1, p0 = ns >> 1, p1 = p0 | 8;
    // mother-code outputs g1, g2 for (predecessor p0, input b); p1 flips both
    const int g1 = b ^ (p0 & 1) ^ ((p0 >> 3) & 1);
    const int g2 = b ^ ((p0 >> 1) & 1) ^ ((p0 >> 2) & 1) ^ ((p0 >> 3) & 1);
    const int c1 = g1 ? 1 : -1, c2 = g2 ? 1 : -1;
    int pm = (l16 == 0) ? 0 : -(1 << 28);
    uint32_t dec[NW];
    int k = 0;                                   // interleaver walk: k_of_i = a (i+1) mod K3
    const int a = P.a;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        uint32_t word = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int t = w * 32 + j;
            if (t < K2) {
                int bm;
                if ((t & 1) == 0) {
                    k += a; if (k >= K3) k -= K3;
                    int v1 = s4[hw][k];
                    k += a; if (k >= K3) k -= K3;
                    int v2 = s4[hw][k];
                    bm = c1 * v1 + c2 * v2;
                } else {
                    k += a; if (k >= K3) k -= K3;
                    bm = c1 * (int)s4[hw][k];
                }
                int m0 = __shfl_sync(0xffffffffu, pm, base + p0) + bm;
                int m1 = __shfl_sync(0xffffffffu, pm, base + p1) - bm;
                bool take1 = m1 > m0;
                pm = take1 ? m1 : m0;
                word |= (take1 ? 1u : 0u) << j;
            }
        }
        dec[w] = word;
    }
    // ---- traceback from state 0 (the four tail bits flush the register)
    const int final_metric = __shfl_sync(0xffffffffu, pm, base);
    int state = 0;
    uint32_t outw[NW];
#pragma unroll
    for (int w = NW - 1; w >= 0; --w) {
        uint32_t ow = 0;
#pragma unroll
        for (int j = 31; j >= 0; --j) {
            const int t = w * 32 + j;
            if (t < K2) {
                ow |= (uint32_t)(state & 1) << (31 - j);                 // MSB-first packing
                uint32_t word = __shfl_sync(0xffffffffu, dec[w], base + state);
                int d = (word >> j) & 1;
                state = (state >> 1) | (d << 3);
            }
        }
        outw[w] = ow;
    }
    if (!live) return;
    // ---- CRC-16 over the K1 type-1 bits, compared with the 16 received parity bits (complemented)
    uint32_t crc = 0xFFFF, rx = 0;
    const int K1 = P.K1;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int t = w * 32 + j;
            if (t < K2) {
                uint32_t bit = (outw[w] >> (31 - j)) & 1u;
                if (t < K1) {
                    uint32_t fb = (bit ^ (crc >> 15)) & 1u;
                    crc = ((crc << 1) & 0xFFFF) ^ (fb ? 0x1021u : 0u);
                } else if (t < K1 + 16) {
                    rx = (rx << 1) | bit;
                }
            }
        }
    }
    crc = (~crc) & 0xFFFF;
    // ---- store: lane j of the half-warp writes word j (big-endian bytes), lane 15 the flags
    uint8_t *ob = bits_out + burst * 40;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        if (l16 == w) {
            uint32_t v = outw[w];
            // clear bits beyond K1
            int lo = w * 32;
            if (lo + 32 > K1) v &= (lo >= K1) ? 0u : (0xFFFFFFFFu << (lo + 32 - K1));
            ob[4 * w] = (uint8_t)(v >> 24); ob[4 * w + 1] = (uint8_t)(v >> 16);
            ob[4 * w + 2] = (uint8_t)(v >> 8); ob[4 * w + 3] = (uint8_t)v;
        }
    }
    if (l16 == 15) {
        for (int i = 4 * NW; i < 40; ++i) ob[i] = 0;
        crc_ok[burst] = (uint8_t)(crc == rx);
        metric[burst] = final_metric;
    }
}

static void scramble_words(uint32_t init, int n, uint32_t *words)
{
    static const int taps[14] = {32, 26, 23, 22, 16, 12, 11, 10, 8, 7, 5, 4, 2, 1};
    for (int i = 0; i < 14; ++i) words[i] = 0;
    uint32_t lfsr = init;
    for (int k = 0; k < n; ++k) {
        uint32_t bit = 0;
        for (int t = 0; t < 14; ++t) bit ^= (lfsr >> (32 - taps[t])) & 1u;
        lfsr = (lfsr >> 1) | (bit << 31);
        words[k >> 5] |= bit << (k & 31);
    }
}

int launch_etsi_decode(tb200_handle h, const int8_t *soft, long n, int channel, uint32_t scramb_init, uint8_t *bits,
                       uint8_t *crc_ok, int32_t *metric, cudaStream_t st)
{
    static const int tab[3][4] = {{124, 144, 216, 101}, {268, 288, 432, 103}, {60, 80, 120, 11}};
    if (channel < 0 || channel > 2) return fail(TB200_E_ARG, "tb200_etsi_decode_batch: unknown logical channel %d", channel);
    if (n <= 0) return TB200_OK;
    EtsiParams P;
    P.K1 = tab[channel][0]; P.K2 = tab[channel][1]; P.K3 = tab[channel][2]; P.a = tab[channel][3];
    scramble_words(scramb_init, P.K3, P.scr);
    unsigned blocks = (unsigned)((n + 7) / 8);
    if (channel == 0) etsi_decode_kernel<144><<<blocks, 128, 0, st>>>(soft, n, P, bits, crc_ok, metric);
    else if (channel == 1) etsi_decode_kernel<288><<<blocks, 128, 0, st>>>(soft, n, P, bits, crc_ok, metric);
    else etsi_decode_kernel<80><<<blocks, 128, 0, st>>>(soft, n, P, bits, crc_ok, metric);
    TB_LAUNCH_CHECK(h);
    return TB200_OK;
}

}  // namespace tb200
