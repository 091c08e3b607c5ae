// etsi_codec.cu — `etsi` mode channel decoder over batches of bursts: descrambling, block
// de-interleaving (K, a), RCPC rate-2/3 de-puncturing, 16-state Viterbi, CRC-16
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
