// etsi_codec.cu — `etsi` mode channel decoder over batches of bursts: descrambling, block
// de-interleaving (K, a), RCPC de-puncturing (rate 2/3 specialised; 1/3, 292/432, 148/432 general), 16-state Viterbi, CRC-16
// (ETSI EN 300 392-2 clause 8).  The reference has no implementation of this stage
// (/root/reference/tetraear/core/protocol.py:300: "we do not do full channel decoding"), so
// parity is unpinned; the kernel is checked bit for bit against oracle/etsi_oracle.py and by
// encode -> decode round trips.
//
// Mapping: ONE THREAD PER BLOCK.  The 16 path metrics live in the thread's registers and the add-compare-select
// of a step is 16 independent, fully unrolled register operations (the two predecessors of state ns are ns>>1
// and (ns>>1)|8; they differ only in the oldest register bit, which enters every generator polynomial, so their
// branch metrics are +bm and -bm, and bm is one of +-(v1+v2), +-(v1-v2) chosen at compile time per state).
// A CTA stages its blocks' soft bits in shared memory with coalesced loads, already descrambled and
// DE-INTERLEAVED ([position][thread], pitch THREADS+4 bytes: conflict-free both ways), and keeps the survivor
// decisions there as one uint16 per (step, thread).  No shuffles, no inter-thread traffic: ~110 instructions per
// trellis step per block instead of ~320 lane-instructions with the former 16-lanes-per-block mapping.
#include <climits>
#include <vector>
#include "common.h"
#include "kernels.h"

namespace tb200 {

struct EtsiParams {
    uint32_t scr[14];      // scrambling sequence bits p(k), bit k%32 of word k/32
    uint16_t inv[432];     // de-interleaver: received position k feeds trellis input inv[k]  (k = a (i+1) mod K3)
    int K1, K2, K3, a;
};

// +-1 weights of the two kept mother-code outputs for next state ns reached from predecessor ns>>1
__host__ __device__ constexpr int etsi_c1(int ns) { return ((ns & 1) ^ ((ns >> 1) & 1) ^ ((ns >> 4) & 1)) ? 1 : -1; }
__host__ __device__ constexpr int etsi_c2(int ns) { return ((ns & 1) ^ ((ns >> 2) & 1) ^ ((ns >> 3) & 1) ^ ((ns >> 4) & 1)) ? 1 : -1; }

template <bool PAIR>
__device__ __forceinline__ uint32_t acs_step(int (&pm)[16], int v1, int v2)
{
    const int s = v1 + v2, d = v1 - v2;
    int np[16];
    uint32_t word = 0;
#pragma unroll
    for (int ns = 0; ns < 16; ++ns) {
        const int c1 = etsi_c1(ns), c2 = etsi_c2(ns);
        const int bm = PAIR ? (c1 == c2 ? (c1 > 0 ? s : -s) : (c1 > 0 ? d : -d)) : (c1 > 0 ? v1 : -v1);
        const int m0 = pm[ns >> 1] + bm, m1 = pm[(ns >> 1) | 8] - bm;
        const bool take1 = m1 > m0;
        np[ns] = take1 ? m1 : m0;
        word |= (take1 ? 1u : 0u) << ns;
    }
#pragma unroll
    for (int ns = 0; ns < 16; ++ns) pm[ns] = np[ns];
    return word;
}

template <int K2, int K3, int THREADS>
__global__ void __launch_bounds__(THREADS)
etsi_decode_kernel(const int8_t *__restrict__ soft, long n, EtsiParams P, uint8_t *__restrict__ bits_out,
                   uint8_t *__restrict__ crc_ok, int32_t *__restrict__ metric)
{
    constexpr int NW = (K2 + 31) / 32;
    constexpr int PITCH = THREADS + 4;                      // bytes per trellis input row: (THREADS + 4) / 4 is odd
    extern __shared__ __align__(16) unsigned char etsi_sm[];
    int8_t *T = reinterpret_cast<int8_t *>(etsi_sm);        // [K3][PITCH] descrambled, de-interleaved soft bits
    uint16_t *D = reinterpret_cast<uint16_t *>(etsi_sm + ((K3 * PITCH + 15) & ~15));   // [K2][THREADS] survivor decisions
    const int tid = threadIdx.x;
    const long blk0 = (long)blockIdx.x * THREADS;
    const int nb = (int)((n - blk0 < THREADS) ? (n - blk0) : THREADS);
    // the de-interleaver table and the scrambling words live in the kernel's parameter block: indexed per lane they are
    // 32 serialised constant-bank reads per warp instruction (52 % of this kernel's stall samples in round 1's form),
    // so a CTA copies them to shared memory once
    __shared__ uint16_t inv_s[K3];
    __shared__ uint32_t scr_s[14];
    for (int i = tid; i < K3; i += THREADS) inv_s[i] = P.inv[i];
    if (tid < 14) scr_s[tid] = P.scr[tid];
    __syncthreads();
    // ---- coalesced load of nb * K3 contiguous bytes; descramble (sign flip where p(k) = 1) and de-interleave
    {
        const int8_t *src = soft + blk0 * K3;
        const int tot = nb * K3;
        int r = tid / K3, k = tid - r * K3;                 // row / column of this thread's first byte
        constexpr int DR = THREADS / K3, DK = THREADS - DR * K3;
        for (int idx = tid; idx < tot; idx += THREADS) {
            const int v = src[idx];
            const int pk = (scr_s[k >> 5] >> (k & 31)) & 1;
            T[(int)inv_s[k] * PITCH + r] = (int8_t)(pk ? -v : v);
            r += DR; k += DK;
            if (k >= K3) { k -= K3; ++r; }
        }
    }
    __syncthreads();
    const long burst = blk0 + tid;
    if (tid >= nb) return;
    // ---- forward pass: steps come in (pair, single) couples — rate 2/3 keeps mother bits {1, 2, 5} of every 8
    int pm[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) pm[i] = (i == 0) ? 0 : -(1 << 28);
    const int8_t *tp = T + tid;
    for (int t = 0; t < K2; t += 2) {
        const int v1 = tp[0], v2 = tp[PITCH], v3 = tp[2 * PITCH];
        tp += 3 * PITCH;
        D[t * THREADS + tid] = (uint16_t)acs_step<true>(pm, v1, v2);
        D[(t + 1) * THREADS + tid] = (uint16_t)acs_step<false>(pm, v3, 0);
    }
    // ---- traceback from state 0 (the four tail bits flush the register)
    const int final_metric = pm[0];
    int state = 0;
    uint32_t outw[NW];
#pragma unroll
    for (int w = NW - 1; w >= 0; --w) {
        uint32_t ow = 0;
        const int jhi = (K2 - w * 32 < 32) ? (K2 - w * 32 - 1) : 31;
        for (int j = jhi; j >= 0; --j) {
            const int t = w * 32 + j;
            ow |= (uint32_t)(state & 1) << (31 - j);                 // MSB-first packing
            const int dbit = (D[t * THREADS + tid] >> state) & 1;
            state = (state >> 1) | (dbit << 3);
        }
        outw[w] = ow;
    }
    // ---- CRC-16 over the K1 type-1 bits, compared with the 16 received parity bits (complemented)
    uint32_t crc = 0xFFFF, rx = 0;
    const int K1 = P.K1;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        const int jn = (K2 - w * 32 < 32) ? (K2 - w * 32) : 32;
        for (int j = 0; j < jn; ++j) {
            const int t = w * 32 + j;
            const uint32_t bit = (outw[w] >> (31 - j)) & 1u;
            if (t < K1) {
                const uint32_t fb = (bit ^ (crc >> 15)) & 1u;
                crc = ((crc << 1) & 0xFFFF) ^ (fb ? 0x1021u : 0u);
            } else if (t < K1 + 16) {
                rx = (rx << 1) | bit;
            }
        }
    }
    crc = (~crc) & 0xFFFF;
    // ---- store: 40 bytes per block (type-1 bits MSB first, zero padded)
    uint32_t *ob = reinterpret_cast<uint32_t *>(bits_out + burst * 40);
#pragma unroll
    for (int w = 0; w < 10; ++w) {
        uint32_t v = 0;
        if (w < NW) {
            v = outw[w];
            const int lo = w * 32;
            if (lo + 32 > K1) v &= (lo >= K1) ? 0u : (0xFFFFFFFFu << (lo + 32 - K1));     // clear bits beyond K1
        }
        ob[w] = __byte_perm(v, 0, 0x0123);                 // big-endian bytes
    }
    crc_ok[burst] = (uint8_t)(crc == rx);
    metric[burst] = final_metric;
}


// ---- general form: any RCPC rate of cl. 8.2.3.1 (1/3, 292/432, 148/432 — and 2/3 for block sizes without a
// specialised instantiation).  None of the standard's puncturing tables ever keeps the fourth mother bit, so a trellis
// step has three inputs (g1, g2, g3); punctured positions are erasures (0), which the load phase leaves in the
// zero-filled array.  Otherwise the same mapping: one thread per block, metrics in registers, inputs and survivor
// words in shared memory.
struct EtsiGenParams {
    uint32_t scr[14];      // scrambling sequence bits
    uint16_t row[432];     // received position k -> row of the de-interleaved array (= type-3 index j - 1)
    uint16_t step[960];    // [K2][3]: rows feeding g1, g2, g3 of every trellis step (K3 = the zero row: punctured)
    int K1, K2, K3, crc;
};

__host__ __device__ constexpr int etsi_c3(int ns) { return ((ns & 1) ^ ((ns >> 1) & 1) ^ ((ns >> 2) & 1) ^ ((ns >> 4) & 1)) ? 1 : -1; }

__device__ __forceinline__ uint32_t acs_step3(int (&pm)[16], int v1, int v2, int v3)
{
    const int a = v1 + v2, b = v1 - v2;
    const int s0 = a + v3, s1 = a - v3, s2 = b + v3, s3 = b - v3;       // v1 + c2 v2 + c3 v3 for (c2, c3) = ++, +-, -+, --
    int np[16];
    uint32_t word = 0;
#pragma unroll
    for (int ns = 0; ns < 16; ++ns) {
        const int c1 = etsi_c1(ns), c2 = etsi_c2(ns) * c1, c3 = etsi_c3(ns) * c1;     // relative to c1
        const int mag = (c2 > 0) ? (c3 > 0 ? s0 : s1) : (c3 > 0 ? s2 : s3);
        const int bm = (c1 > 0) ? mag : -mag;
        const int m0 = pm[ns >> 1] + bm, m1 = pm[(ns >> 1) | 8] - bm;
        const bool take1 = m1 > m0;
        np[ns] = take1 ? m1 : m0;
        word |= (take1 ? 1u : 0u) << ns;
    }
#pragma unroll
    for (int ns = 0; ns < 16; ++ns) pm[ns] = np[ns];
    return word;
}

// Shared memory per CTA: the K3 received values per block ([K3 + 1][PITCH] int8, row K3 = zeros = "punctured"), the
// survivor words ([K2][THREADS] uint16) and a per-step descriptor (three row numbers, the same for every thread:
// broadcast reads).  The first form kept a [3 K2][PITCH] input array with the erasures stored: 1.5 KB per block at
// K2 = 292, two CTAs of 64 threads per SM.
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
etsi_decode_gen_kernel(const int8_t *__restrict__ soft, long n, const __grid_constant__ EtsiGenParams P,
                       uint8_t *__restrict__ bits_out, uint8_t *__restrict__ crc_ok, int32_t *__restrict__ metric)
{
    constexpr int PITCH = THREADS + 4;
    extern __shared__ __align__(16) unsigned char etsi_sm[];
    const int K1 = P.K1, K2 = P.K2, K3 = P.K3;
    int8_t *T = reinterpret_cast<int8_t *>(etsi_sm);                                        // [K3 + 1][PITCH]
    uint16_t *D = reinterpret_cast<uint16_t *>(etsi_sm + (((K3 + 1) * PITCH + 15) & ~15));   // [K2][THREADS]
    uint16_t *desc = D + K2 * THREADS;                                                      // [K2][3] row numbers
    __shared__ uint16_t row_s[432];
    __shared__ uint32_t scr_s[14];
    const int tid = threadIdx.x;
    const long blk0 = (long)blockIdx.x * THREADS;
    const int nb = (int)((n - blk0 < THREADS) ? (n - blk0) : THREADS);
    for (int i = tid; i < K3; i += THREADS) row_s[i] = P.row[i];
    if (tid < 14) scr_s[tid] = P.scr[tid];
    for (int i = tid; i < 3 * K2; i += THREADS) desc[i] = P.step[i];
    for (int i = tid; i < PITCH; i += THREADS) T[K3 * PITCH + i] = 0;
    __syncthreads();
    {
        const int8_t *src = soft + blk0 * K3;
        const int tot = nb * K3;
        int r = tid / K3, k = tid - r * K3;
        const int DR = THREADS / K3, DK = THREADS - DR * K3;
        for (int idx = tid; idx < tot; idx += THREADS) {
            const int v = src[idx];
            const int pk = (scr_s[k >> 5] >> (k & 31)) & 1;
            T[(int)row_s[k] * PITCH + r] = (int8_t)(pk ? -v : v);
            r += DR; k += DK;
            if (k >= K3) { k -= K3; ++r; }
        }
    }
    __syncthreads();
    const long burst = blk0 + tid;
    if (tid >= nb) return;
    int pm[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) pm[i] = (i == 0) ? 0 : -(1 << 28);
    const int8_t *tp = T + tid;
    for (int t = 0; t < K2; ++t) {
        const int r1 = desc[3 * t], r2 = desc[3 * t + 1], r3 = desc[3 * t + 2];
        const int v1 = tp[r1 * PITCH], v2 = tp[r2 * PITCH], v3 = tp[r3 * PITCH];
        D[t * THREADS + tid] = (uint16_t)acs_step3(pm, v1, v2, v3);
    }
    const int final_metric = pm[0];
    // traceback: decoded bit t = the newest register bit of the state after step t; packed MSB first
    uint32_t outw[10];
#pragma unroll
    for (int w = 0; w < 10; ++w) outw[w] = 0u;
    int state = 0;
    for (int t = K2 - 1; t >= 0; --t) {
        const uint32_t bit = (uint32_t)(state & 1) << (31 - (t & 31));
        const int w = t >> 5;
#pragma unroll
        for (int u = 0; u < 10; ++u) if (u == w) outw[u] |= bit;
        const int dbit = (D[t * THREADS + tid] >> state) & 1;
        state = (state >> 1) | (dbit << 3);
    }
    uint32_t crc = 0xFFFF, rx = 0;
    if (P.crc) {
#pragma unroll
        for (int w = 0; w < 10; ++w) {
            for (int j = 0; j < 32; ++j) {
                const int t = w * 32 + j;
                if (t >= K1 + 16) break;
                const uint32_t bit = (outw[w] >> (31 - j)) & 1u;
                if (t < K1) {
                    const uint32_t fb = (bit ^ (crc >> 15)) & 1u;
                    crc = ((crc << 1) & 0xFFFF) ^ (fb ? 0x1021u : 0u);
                } else {
                    rx = (rx << 1) | bit;
                }
            }
        }
        crc = (~crc) & 0xFFFF;
    }
    uint32_t *ob = reinterpret_cast<uint32_t *>(bits_out + burst * 40);
#pragma unroll
    for (int w = 0; w < 10; ++w) {
        uint32_t v = outw[w];
        const int lo = w * 32;
        if (lo + 32 > K1) v &= (lo >= K1) ? 0u : (0xFFFFFFFFu << (lo + 32 - K1));
        ob[w] = __byte_perm(v, 0, 0x0123);
    }
    crc_ok[burst] = P.crc ? (uint8_t)(crc == rx) : (uint8_t)2;          // 2: the channel has no block code to check
    metric[burst] = final_metric;
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

// ---- AACH: (30, 14) shortened second-order Reed-Muller block code (cl. 8.2.3.2), soft maximum-likelihood decoding.
// One warp per block: lane L correlates the 512 codewords whose five leading information bits spell L, walking the
// other nine in Gray-code order (one generator row XORed per step; the three fastest bits are unrolled, their rows sit in
// registers).  The correlation of a codeword is 2 * (sum of the soft values where it has a 1) - (sum of all), and the
// sum over its ones is taken five code bits at a time from six 32-entry tables of partial sums that live ONE ENTRY PER
// LANE in six registers and are read with warp shuffles (index = five bits of the codeword): no shared memory, no bank
// conflicts (a first version with byte-indexed tables in shared memory spent its time on 32-way-random 16-bit reads:
// 1.0 G bank conflicts and 70 instructions per codeword, 32 ms per 1 M blocks).  Winner and runner-up are tracked as
// one integer key = correlation * 2^14 + (16383 - word): max = best correlation, ties to the smaller word.
// The generator is a parameter (14 rows, code bit j in word bit 29 - j): any (30, 14) linear code decodes the same way.
struct AachParams {
    uint32_t gen[14];      // gen[k]: row of information bit k + 1 (MSB of the 14-bit result first)
    uint32_t scr;          // scrambling bits of the 30 positions, bit j = p(j + 1)
};

__global__ void __launch_bounds__(128)
etsi_aach_kernel(const int8_t *__restrict__ soft, long n, AachParams P, uint16_t *__restrict__ bits_out,
                 int32_t *__restrict__ metric, int32_t *__restrict__ margin)
{
    const int lane = threadIdx.x & 31;
    const long blk = (long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (blk >= n) return;
    // lane w < 30 holds the descrambled soft value of WORD bit w = code bit 29 - w
    int v = 0;
    if (lane < 30) {
        const int j = 29 - lane;
        v = soft[blk * 30 + j];
        if ((P.scr >> j) & 1u) v = -v;
    }
    int tot = v;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    // T[g] on lane L = sum of the soft values of word bits 5g + u over the set bits u of L
    int T[6];
#pragma unroll
    for (int g = 0; g < 6; ++g) {
        int acc = 0;
#pragma unroll
        for (int u = 0; u < 5; ++u) {
            const int sv = __shfl_sync(0xffffffffu, v, 5 * g + u);
            if ((lane >> u) & 1) acc += sv;
        }
        T[g] = acc;
    }
    uint32_t cw = 0;
#pragma unroll
    for (int k = 0; k < 5; ++k) if ((lane >> (4 - k)) & 1) cw ^= P.gen[k];
    const uint32_t r0 = P.gen[13], r1 = P.gen[12], r2 = P.gen[11];
    // key = corr * 16384 + 16383 - word = ones * 32768 + (16383 - tot * 16384 - (lane << 9)) - gray
    const int k0 = 16383 - tot * 16384 - (lane << 9);
    int best = INT_MIN, second = INT_MIN;
    int gray = 0;
    auto eval = [&](uint32_t c, int gr) {
        int ones = __shfl_sync(0xffffffffu, T[0], (int)c);
#pragma unroll
        for (int g = 1; g < 6; ++g) ones += __shfl_sync(0xffffffffu, T[g], (int)(c >> (5 * g)));     // source lane is taken mod 32
        const int key = ones * 32768 + k0 - gr;
        second = max(second, min(best, key));
        best = max(best, key);
    };
    for (int hi = 0; hi < 64; ++hi) {
        if (hi) {
            const int b = __ffs(hi) + 2;                  // Gray code over bits 3..8 of the low nine
            gray ^= 1 << b;
            cw ^= P.gen[13 - b];
        }
        eval(cw, gray);
        cw ^= r0; gray ^= 1; eval(cw, gray);
        cw ^= r1; gray ^= 2; eval(cw, gray);
        cw ^= r0; gray ^= 1; eval(cw, gray);
        cw ^= r2; gray ^= 4; eval(cw, gray);
        cw ^= r0; gray ^= 1; eval(cw, gray);
        cw ^= r1; gray ^= 2; eval(cw, gray);
        cw ^= r0; gray ^= 1; eval(cw, gray);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int ob = __shfl_xor_sync(0xffffffffu, best, o), os = __shfl_xor_sync(0xffffffffu, second, o);
        second = max(max(second, os), min(best, ob));
        best = max(best, ob);
    }
    if (lane == 0) {
        bits_out[blk] = (uint16_t)(16383 - (best & 16383));
        metric[blk] = best >> 14;                         // arithmetic shift = floor: the low 14 bits are non-negative
        if (margin) margin[blk] = (best >> 14) - (second >> 14);
    }
}

int launch_etsi_aach(tb200_handle h, const int8_t *soft, long n, const uint32_t *gen_rows, uint32_t scramb_init, uint16_t *bits,
                     int32_t *metric, int32_t *margin, cudaStream_t st)
{
    if (n <= 0) return TB200_OK;
    AachParams P;
    uint32_t seen = 0;
    for (int k = 0; k < 14; ++k) {
        P.gen[k] = gen_rows[k] & 0x3FFFFFFFu;
        if (!P.gen[k]) return fail(TB200_E_ARG, "tb200_etsi_aach_decode: generator row %d is zero", k);
        seen |= P.gen[k];
    }
    (void)seen;
    uint32_t words[14];
    scramble_words(scramb_init, 30, words);
    P.scr = words[0] & 0x3FFFFFFFu;
    etsi_aach_kernel<<<(unsigned)((n + 3) / 4), 128, 0, st>>>(soft, n, P, bits, metric, margin);
    TB_LAUNCH_CHECK(h);
    return TB200_OK;
}

// any (K1, K2, K3, a) block with one of the four RCPC rates of cl. 8.2.3.1: rate 0 = 2/3, 1 = 1/3, 2 = 292/432, 3 = 148/432
int launch_etsi_decode_custom(tb200_handle h, const int8_t *soft, long n, int K1, int K2, int K3, int a, int rate, int crc,
                              uint32_t scramb_init, uint8_t *bits, uint8_t *crc_ok, int32_t *metric, cudaStream_t st)
{
    static const int P23[3] = {1, 2, 5}, P13[6] = {1, 2, 3, 5, 6, 7};
    if (rate < 0 || rate > 3) return fail(TB200_E_ARG, "tb200_etsi_decode_custom: unknown RCPC rate id %d", rate);
    if (K1 < 1 || K2 > 320 || K3 < 1 || K3 > 432 || a < 1 || K2 != K1 + (crc ? 16 : 0) + 4)
        return fail(TB200_E_ARG, "tb200_etsi_decode_custom: unsupported block geometry K1 = %d, K2 = %d, K3 = %d", K1, K2, K3);
    const int *Pt = (rate == 0 || rate == 2) ? P23 : P13;
    const int t = (rate == 0 || rate == 2) ? 3 : 6;
    const int div = rate == 2 ? 65 : rate == 3 ? 35 : 0;
    EtsiGenParams P;
    P.K1 = K1; P.K2 = K2; P.K3 = K3; P.crc = crc ? 1 : 0;
    scramble_words(scramb_init, K3, P.scr);
    // type-3 bit j (1-based) is mother bit k = 8 ((i-1) div t) + P(i - t ((i-1) div t)), i = j + (j-1) div d;
    // it was sent at type-4 position 1 + (a j mod K3)
    std::vector<char> seen((size_t)K3, 0);
    for (int i = 0; i < 3 * K2; ++i) P.step[i] = (uint16_t)K3;
    for (int j = 1; j <= K3; ++j) {
        const int i = div ? j + (j - 1) / div : j;
        const int blk = (i - 1) / t;
        const int k = 8 * blk + Pt[i - t * blk - 1];                     // 1-based mother bit
        const int step = (k - 1) / 4, g = (k - 1) % 4;
        if (step >= K2 || g > 2) return fail(TB200_E_ARG, "tb200_etsi_decode_custom: K3 = %d type-3 bits do not fit rate id %d with K2 = %d", K3, rate, K2);
        const int pos = (int)(((long)a * j) % K3);
        if (seen[pos]) return fail(TB200_E_ARG, "tb200_etsi_decode_custom: interleaver (%d, %d) is not a permutation", K3, a);
        seen[pos] = 1;
        P.row[pos] = (uint16_t)(j - 1);
        P.step[3 * step + g] = (uint16_t)(j - 1);
    }
    if ((reinterpret_cast<uintptr_t>(bits) & 3) != 0) return fail(TB200_E_ARG, "tb200_etsi_decode_custom: d_bits must be 4-byte aligned");
    if (n <= 0) return TB200_OK;
    constexpr int T = 64;
    const size_t sm = (size_t)((((K3 + 1) * (T + 4) + 15) & ~15) + K2 * T * 2 + 3 * K2 * 2);
    cudaError_t e = cudaFuncSetAttribute(etsi_decode_gen_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return fail(TB200_E_CUDA, "cudaFuncSetAttribute(etsi_decode_gen_kernel): %s", cudaGetErrorString(e));
    etsi_decode_gen_kernel<T><<<(unsigned)((n + T - 1) / T), T, sm, st>>>(soft, n, P, bits, crc_ok, metric);
    TB_LAUNCH_CHECK(h);
    return TB200_OK;
}

int launch_etsi_decode(tb200_handle h, const int8_t *soft, long n, int channel, uint32_t scramb_init, uint8_t *bits,
                       uint8_t *crc_ok, int32_t *metric, cudaStream_t st)
{
    static const int tab[4][4] = {{124, 144, 216, 101}, {268, 288, 432, 103}, {60, 80, 120, 11}, {92, 112, 168, 13}};
    if (channel == 4) return launch_etsi_decode_custom(h, soft, n, 288, 292, 432, 103, 2, 0, scramb_init, bits, crc_ok, metric, st);
    if (channel == 5) return launch_etsi_decode_custom(h, soft, n, 144, 148, 432, 103, 3, 0, scramb_init, bits, crc_ok, metric, st);
    if (channel < 0 || channel > 5) return fail(TB200_E_ARG, "tb200_etsi_decode_batch: unknown logical channel %d", channel);
    if (n <= 0) return TB200_OK;
    EtsiParams P;
    P.K1 = tab[channel][0]; P.K2 = tab[channel][1]; P.K3 = tab[channel][2]; P.a = tab[channel][3];
    scramble_words(scramb_init, P.K3, P.scr);
    for (int i = 0; i < P.K3; ++i) P.inv[(int)(((long)P.a * (i + 1)) % P.K3)] = (uint16_t)i;
    if ((reinterpret_cast<uintptr_t>(bits) & 3) != 0) return fail(TB200_E_ARG, "tb200_etsi_decode_batch: d_bits must be 4-byte aligned");
    auto smem_of = [](int K2, int K3, int T) { return (size_t)(((K3 * (T + 4) + 15) & ~15) + K2 * T * 2); };
    cudaError_t e = cudaSuccess;
    if (channel == 0) {
        constexpr int T = 128;
        const size_t sm = smem_of(144, 216, T);
        e = cudaFuncSetAttribute(etsi_decode_kernel<144, 216, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e == cudaSuccess) etsi_decode_kernel<144, 216, T><<<(unsigned)((n + T - 1) / T), T, sm, st>>>(soft, n, P, bits, crc_ok, metric);
    } else if (channel == 1) {
        constexpr int T = 64;
        const size_t sm = smem_of(288, 432, T);
        e = cudaFuncSetAttribute(etsi_decode_kernel<288, 432, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e == cudaSuccess) etsi_decode_kernel<288, 432, T><<<(unsigned)((n + T - 1) / T), T, sm, st>>>(soft, n, P, bits, crc_ok, metric);
    } else if (channel == 3) {
        constexpr int T = 128;
        const size_t sm = smem_of(112, 168, T);
        e = cudaFuncSetAttribute(etsi_decode_kernel<112, 168, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e == cudaSuccess) etsi_decode_kernel<112, 168, T><<<(unsigned)((n + T - 1) / T), T, sm, st>>>(soft, n, P, bits, crc_ok, metric);
    } else {
        constexpr int T = 128;
        const size_t sm = smem_of(80, 120, T);
        e = cudaFuncSetAttribute(etsi_decode_kernel<80, 120, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e == cudaSuccess) etsi_decode_kernel<80, 120, T><<<(unsigned)((n + T - 1) / T), T, sm, st>>>(soft, n, P, bits, crc_ok, metric);
    }
    if (e != cudaSuccess) return fail(TB200_E_CUDA, "cudaFuncSetAttribute(etsi_decode_kernel): %s", cudaGetErrorString(e));
    TB_LAUNCH_CHECK(h);
    return TB200_OK;
}

}  // namespace tb200
