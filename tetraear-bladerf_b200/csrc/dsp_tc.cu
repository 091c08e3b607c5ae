// dsp_tc.cu — tensor-core ("TC") form of the time-parallel zero-phase decimator of dsp_fast.cu, for
// scipy.signal.decimate(x, q) as called by SignalProcessor.process
// (/root/reference/tetraear/signal/processor.py:254).
//
// Same filter and the same two-sided modal form as dsp_fast.cu (see its header):
//     y[k] = 2 Re sum_i alpha_i ( F_i[n_k] + p_i B_i[n_k + 1] ) - right-edge transient,   n_k = 27 + q k,
// but everything that is a plain linear map of the input samples is one dense contraction on the 5th-generation
// tensor cores.  A lane owns R = 3 consecutive outputs, i.e. 3 q consecutive input samples (its "slot").  For every
// slot one tcgen05.mma row produces, from the slot's own samples only,
//     * Y0[r]   the R outputs with zero state on both sides of the slot (weights h[|t - n_k|] = 2 sum Re alpha p^|d|),
//     * AF_i    the causal lane aggregate   2 alpha_i     sum p_i^(n_last - t) x[t]    (complex, 4 modes),
//     * AB_i    the anti-causal aggregate   2 alpha_i p_i sum p_i^(t - t_first) x[t]   (complex, 4 modes),
// for I and Q at once: the A operand is the interleaved complex64 input exactly as it lies in HBM (one TMA bulk copy
// per group of 128 slots; the threads re-lay it into the canonical K-major shared-memory layout, one 16-byte chunk
// of every slot per core-matrix row), B holds the weights with zeros where I would meet Q.  float32 accuracy
// comes from a split: x = hi + lo with hi, lo representable in TF32, B = Bhi + Blo, three MMAs per k-step
// (hi*Bhi + lo*Bhi + hi*Blo) accumulated in TMEM (missing term ~2^-22).  What is left for the CUDA cores is the
// part that is a recurrence: the Kogge-Stone scans of the aggregates over the 32 lanes of a pass (shuffles), the carry
// of the neighbour pass, and three multiply-adds per output.
//
// A CTA streams over a contiguous range of 384-output groups (one MMA of M = 128 slots): the causal carry chains
// from group to group, the anti-causal carry of a group's last pass arrives one group later (those 96 outputs are
// held back in registers), chunk ends use the analytic steady state of scipy's constant continuation, so there is
// no halo inside a range — only one extra group at a range end that falls inside a chunk.
#include <atomic>
#include <complex>
#include <cstring>
#include <mutex>
#include <vector>
#include "common.h"
#include "kernels.h"
#include "packed.h"
#include "decim_common.h"

namespace tb200 {

constexpr int TC_R = FD_R;                 // outputs per lane slot
constexpr int TC_PASS = 32 * TC_R;         // outputs per warp pass
constexpr int TC_GP = 4;                   // passes per group = MMA rows / 32
constexpr int TC_GROUP = TC_GP * TC_PASS;  // outputs per group
constexpr int TC_N = 48;                   // MMA N: 6 zero-state outputs + 16 + 16 aggregate components (+ 10 unused)

template <int Q> struct TcG {
    static constexpr int KF = 2 + 2 * TC_R * Q;          // floats of a slot that carry weight positions: 1 pad sample in front
    static constexpr int K = (KF + 15) / 16 * 16;        // contraction length (TF32 MMA K = 8; 16 so that lo goes to TMEM in 16-column stores)
    static constexpr int KC = K / 4;                     // 16-byte chunks per slot
    static constexpr int A_BYTES = KC * 128 * 16;
    static constexpr int B_BYTES = KC * TC_N * 16;
    static constexpr int LEAD = 27 / Q;                  // the stream of a chunk starts LEAD rows before output 0
    static constexpr int SLOT_FLOATS = 2 * TC_R * Q;     // distance of consecutive slots in the input
};

struct TcParams {
    float lampow[5][4][2];        // (lambda^R)^(2^s): Kogge-Stone multipliers
    float lanepow[32][4][2];      // (lambda^R)^l: cross-pass carry weights
    float cF[TC_R][4][2];         // lambda^(r+1)
    float cB[TC_R][4][2];         // lambda^(R-1-r)
    float initF[4][2];            // 2 alpha / (1 - p): causal state of a constant signal
    float initB[4][2];            // 2 alpha p / (1 - p)
};

// ---- PTX wrappers: tcgen05 (TMEM allocation, MMA, commit, load), TMA tensor copy ---------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, unsigned cols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, unsigned cols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// shared-memory matrix descriptor, K-major, no swizzle: core matrix = 8 rows x 16 bytes stored contiguously;
// lbo = byte distance of the two 16-byte K chunks of one MMA, sbo = byte distance of consecutive 8-row groups
__device__ __forceinline__ uint64_t umma_desc(unsigned saddr, unsigned lbo, unsigned sbo)
{
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
           (1ull << 46);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                   "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                   "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

__device__ __forceinline__ cp cpadd(cp a, cp b)
{
    cp o;
    o.re = make_float2(a.re.x + b.re.x, a.re.y + b.re.y);
    o.im = make_float2(a.im.x + b.im.x, a.im.y + b.im.y);
    return o;
}
// packed complex (I, Q) state of a constant signal c: v * c for a complex scalar v
__device__ __forceinline__ cp cp_const(cf v, float2 c)
{
    cp o;
    o.re = make_float2(v.x * c.x, v.x * c.y);
    o.im = make_float2(v.y * c.x, v.y * c.y);
    return o;
}

// out of line: chunk ends only; keeps the group loop's code small
__device__ __noinline__ float2 padded_sample_ool(const float2 *__restrict__ x, long n, long t) { return padded_sample(x, n, t); }

// ---- the kernel ---------------------------------------------------------------------------------------------------
// A CTA holds TC_TEAMS teams of 4 warps; every team streams over its own range of groups with its own buffers,
// barriers and carries (the teams only share the weight matrices and the TMEM allocation), so that the latency of one
// team's load / MMA / scan chain is covered by the others: 2 CTAs x 2 teams = 16 warps and 4 groups in flight per SM.
constexpr int TC_TEAMS = 2;
constexpr int TC_TEAM_COLS = 128;          // TMEM columns per team: accumulators at +0 (48 used), lo operand at +64 (K <= 64)

// one lane of a fully active warp (warp-uniform control flow around the single-thread tcgen05 / TMA instructions)
__device__ __forceinline__ bool elect_one()
{
    unsigned pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred)::"memory");
    return pred != 0;
}
__device__ __forceinline__ void team_sync(int team)
{
    asm volatile("bar.sync %0, 128;\n" ::"r"(1 + team) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n"
                 ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]),
                 "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]: the A operand is read from tensor memory (row = TMEM lane, one TF32 element per column)
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

template <int Q>
__global__ void __launch_bounds__(TC_TEAMS * TC_GP * 32, 2)
decim_tc_kernel(const float2 *__restrict__ x, long n, long chunk_stride, int n_chunks, long n_dec, int gpc, long total_floats,
                TcParams P, const float *__restrict__ bmat, float2 *__restrict__ dec_out, long long *__restrict__ timing)
{
    using G = TcG<Q>;
    extern __shared__ uint8_t tc_raw[];
    uint8_t *sm = tc_raw + ((128u - (smem_u32(tc_raw) & 127u)) & 127u);        // 128-byte aligned, still a shared-space pointer
    uint8_t *B_hi = sm, *B_lo = B_hi + G::B_BYTES;
    __shared__ __align__(8) uint64_t full_bar[TC_TEAMS], mma_bar[TC_TEAMS], b_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ cp carryF_s[TC_TEAMS][2][TC_GP][4];
    __shared__ cp carryB_s[TC_TEAMS][2][TC_GP][4];
    __shared__ cf lanepow_s[32][4];

    const int team = threadIdx.x >> 7, tid = threadIdx.x & 127, w = tid >> 5, lane = tid & 31;
    uint8_t *A_hi = sm + 2 * G::B_BYTES + team * G::A_BYTES;                   // raw group, then the canonical hi operand
    const int NG = n_chunks * gpc;                   // (the launcher keeps n_chunks * gpc below 2^30)
    const int n_ranges = TC_TEAMS * gridDim.x, range = TC_TEAMS * blockIdx.x + team;
    const int G0 = (int)((long)NG * range / n_ranges), G1 = (int)((long)NG * (range + 1) / n_ranges);
    const int gs = G0 - ((G0 % gpc) ? 1 : 0);        // a range that starts inside a chunk first runs the group before it (causal carry)
    const int ge = G1 + ((G1 % gpc) ? 1 : 0);        // ... one that ends inside a chunk also the group after it (anti-causal carry)

    if (team == 0) lanepow_s[lane][w] = ld2(P.lanepow[lane][w]);
    if (threadIdx.x == 0) {
        for (int t = 0; t < TC_TEAMS; ++t) { mbar_init(&full_bar[t], 1); mbar_init(&mma_bar[t], 1); }
        mbar_init(&b_bar, 1);
    }
    if (threadIdx.x < 32) tmem_alloc(&tmem_base_s, TC_TEAMS * TC_TEAM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_base_s + team * TC_TEAM_COLS, tmem_a = tmem_d + 64;

    // float offset (from x) of slot (pass 0, lane 0) of group g (a multiple of 4: 16-byte aligned), and whether the whole
    // raw run of the group lies inside the input so that one bulk copy may fetch it
    constexpr long RAW_FLOATS = 127L * G::SLOT_FLOATS + G::K;
    auto group_f0 = [&](int g) -> long {
        const int c = g / gpc, gi = g - c * gpc;
        return 2 * ((long)c * chunk_stride + (long)Q * (gi * TC_GROUP - G::LEAD - 1));
    };
    auto group_tma = [&](int g) -> bool {
        const long f0 = group_f0(g);
        return f0 >= 0 && f0 + RAW_FLOATS <= total_floats;
    };
    auto issue_load = [&](int g) {
        tma_load_1d(A_hi, reinterpret_cast<const float *>(x) + group_f0(g), (unsigned)(RAW_FLOATS * 4), &full_bar[team]);
    };

    if (threadIdx.x < 32 && elect_one()) tma_load_1d(B_hi, bmat, 2u * G::B_BYTES, &b_bar);    // weights: Bhi then Blo, canonical layout
    if (w == 0 && gs < ge && group_tma(gs) && elect_one()) issue_load(gs);

    const int m = tid;                                              // slot = MMA row = TMEM lane
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t lane_base = (uint32_t)(w * 32) << 16;
    // matrix descriptors of k-step 0 (the start-address field advances by bytes >> 4 per k-step; no carry out of its 14 bits)
    const uint64_t desc_a = umma_desc(smem_u32(A_hi), 2048, 128);
    const uint64_t desc_bh = umma_desc(smem_u32(B_hi), TC_N * 16, 128), desc_bl = umma_desc(smem_u32(B_lo), TC_N * 16, 128);
    unsigned ph_full = 0, ph_mma = 0;
    bool pend = false;
    float2 ypend[TC_R];
    long pend_k0 = 0;
    int pend_chunk = 0;
    ypend[0] = ypend[1] = ypend[2] = make_float2(0.f, 0.f);

    // optional phase timing (profiling aid, tb200_debug_tc_timing): cycles of thread 0 of CTA 0 per phase, summed over its groups
    const bool timed = timing != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
    long long tprev = timed ? clock64() : 0;
    auto tick = [&](int phase) { if (timed) { const long long t = clock64(); timing[phase] += t - tprev; tprev = t; } };
    for (int g = gs; g < ge; ++g) {
        tick(0);
        const int buf = (g - gs) & 1;
        const int chunk = g / gpc;
        const int gi = g - chunk * gpc;
        const bool store = g >= G0 && g < G1;
        const float2 *xc = x + (long)chunk * chunk_stride;
        const long k0 = (long)gi * TC_GROUP - G::LEAD + (long)w * TC_PASS + (long)lane * TC_R;   // first output of this slot
        const long i0 = 1 + (long)Q * (k0 - 1);                     // chunk sample under weight position 0 (slot float 2)
        const bool via_tma = group_tma(g);

        // ---- A operand.  The 128 slots of a group are one contiguous run of the input (slot m starts SLOT_FLOATS * m
        // floats in, slots overlap by the pad floats): ONE bulk copy stages it raw in the A_hi region; every thread
        // then takes its slot into registers and, after a barrier, writes hi = x rounded to TF32 back in the canonical
        // K-major layout (16-byte chunk kc of slot m at kc * 2048 + m * 16) and lo = x - hi (exact) into its lane of
        // tensor memory, from where the second MMA term reads its A operand.
        if (via_tma) {
            mbar_wait(&full_bar[team], ph_full);
            ph_full ^= 1;
        } else {
#pragma unroll 1
            for (int j = 0; j < G::K / 2; ++j) {                     // range at an end of the whole input: plain loads
                const long i = i0 - 1 + j;
                const float2 v = (i >= 0 && i < n) ? xc[i] : padded_sample_ool(xc, n, 27 + i);
                *reinterpret_cast<float2 *>(A_hi + m * (G::SLOT_FLOATS * 4) + j * 8) = v;
            }
            team_sync(team);
        }
        tick(1);
        float4 v[G::KC];
#pragma unroll
        for (int kc = 0; kc < G::KC; ++kc) v[kc] = *reinterpret_cast<const float4 *>(A_hi + m * (G::SLOT_FLOATS * 4) + kc * 16);
        if (via_tma && (i0 - 1 < 0 || i0 - 1 + G::K / 2 > n)) {     // slot reaches outside the chunk: scipy's padded signal there
#pragma unroll
            for (int kc = 0; kc < G::KC; ++kc) {
                const long i = i0 - 1 + 2 * kc;
                if (i < 0 || i >= n) { const float2 e = padded_sample_ool(xc, n, 27 + i); v[kc].x = e.x; v[kc].y = e.y; }
                if (i + 1 < 0 || i + 1 >= n) { const float2 e = padded_sample_ool(xc, n, 28 + i); v[kc].z = e.x; v[kc].w = e.y; }
            }
        }
        team_sync(team);                                             // everybody holds its slot: the raw bytes may be overwritten
        tick(2);
#pragma unroll
        for (int q4 = 0; q4 < G::KC / 4; ++q4) {
            float lo[16];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int kc = 4 * q4 + c;
                float4 hi;
                hi.x = __uint_as_float((__float_as_uint(v[kc].x) + 0x1000u) & 0xFFFFE000u);
                hi.y = __uint_as_float((__float_as_uint(v[kc].y) + 0x1000u) & 0xFFFFE000u);
                hi.z = __uint_as_float((__float_as_uint(v[kc].z) + 0x1000u) & 0xFFFFE000u);
                hi.w = __uint_as_float((__float_as_uint(v[kc].w) + 0x1000u) & 0xFFFFE000u);
                lo[4 * c + 0] = v[kc].x - hi.x; lo[4 * c + 1] = v[kc].y - hi.y;
                lo[4 * c + 2] = v[kc].z - hi.z; lo[4 * c + 3] = v[kc].w - hi.w;
                *reinterpret_cast<float4 *>(A_hi + kc * 2048 + m * 16) = hi;
            }
            tmem_st16(tmem_a + lane_base + 16 * q4, lo);
        }
        tmem_st_wait();
        fence_async_smem();                                         // generic-proxy writes -> visible to the tensor core
        tc_fence_before();
        team_sync(team);
        tick(3);

        // ---- one thread issues the MMAs of the group: D[128 x 48] = sum over k-steps of hi*Bhi + lo*Bhi + hi*Blo
        if (w == 0) {                                               // warp-uniform branch, one elected lane issues
            if (g == gs) mbar_wait(&b_bar, 0);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < G::K / 8; ++ks) {
                    const uint64_t dah = desc_a + (uint64_t)(ks * (4096 >> 4));
                    const uint64_t dbh = desc_bh + (uint64_t)(ks * ((2 * TC_N * 16) >> 4)), dbl = desc_bl + (uint64_t)(ks * ((2 * TC_N * 16) >> 4));
                    umma_tf32(tmem_d, dah, dbh, idesc, ks > 0 ? 1u : 0u);
                    umma_tf32_ts(tmem_d, tmem_a + 8 * ks, dbh, idesc, 1u);
                    umma_tf32(tmem_d, dah, dbl, idesc, 1u);
                }
                umma_commit(&mma_bar[team]);
            }
            __syncwarp();
        }
        tick(4);
        mbar_wait(&mma_bar[team], ph_mma);
        ph_mma ^= 1;
        tc_fence_after();
        tick(5);
        if (w == 0 && g + 1 < ge && group_tma(g + 1) && elect_one()) issue_load(g + 1);   // the MMAs are done with A_hi: prefetch the next group

        // ---- accumulators of this slot: TMEM lane m, columns 0..39
        uint32_t va[32], vb[8];
        tmem_ld32(tmem_d + lane_base, va);
        tmem_ld8(tmem_d + lane_base + 32, vb);
        tmem_ld_wait();
        tc_fence_before();
        tick(6);
        float2 y[TC_R];
#pragma unroll
        for (int r = 0; r < TC_R; ++r) y[r] = make_float2(__uint_as_float(va[2 * r]), __uint_as_float(va[2 * r + 1]));

        // ---- Kogge-Stone scans of the aggregates over the lanes of the pass (zero state at the pass ends)
        cp tF[4], tB[4];
        const float2 zero2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int mm = 0; mm < 4; ++mm) {
            auto col = [&](int cc) -> float { return __uint_as_float(cc < 32 ? va[cc < 32 ? cc : 0] : vb[cc >= 32 ? cc - 32 : 0]); };
            cp a, b;
            a.re = make_float2(col(6 + 4 * mm), col(7 + 4 * mm));
            a.im = make_float2(col(8 + 4 * mm), col(9 + 4 * mm));
            b.re = make_float2(col(22 + 4 * mm), col(23 + 4 * mm));
            b.im = make_float2(col(24 + 4 * mm), col(25 + 4 * mm));
#pragma unroll
            for (int s = 0; s < 5; ++s) {
                if (s >= fd_nstage(mm)) continue;
                cp oa, ob;
                oa.re = shfl2_up(a.re, 1 << s); oa.im = shfl2_up(a.im, 1 << s);
                ob.re = shfl2_down(b.re, 1 << s); ob.im = shfl2_down(b.im, 1 << s);
                const cf gm = ld2(P.lampow[s][mm]);
                if (lane >= (1 << s)) a = cpfma(gm, oa, a);
                if (lane + (1 << s) < 32) b = cpfma(gm, ob, b);
            }
            if (lane == 31) carryF_s[team][buf][w][mm] = a;
            if (lane == 0) carryB_s[team][buf][w][mm] = b;
            tF[mm].re = shfl2_up(a.re, 1); tF[mm].im = shfl2_up(a.im, 1);
            tB[mm].re = shfl2_down(b.re, 1); tB[mm].im = shfl2_down(b.im, 1);
            if (lane == 0) tF[mm].re = tF[mm].im = zero2;
            if (lane == 31) tB[mm].re = tB[mm].im = zero2;
        }
        team_sync(team);
        tick(7);

        // ---- the pass held back from the previous group gets its anti-causal carry from this group's first pass
        if (pend) {
            float2 *out = dec_out + (long)pend_chunk * n_dec;
#pragma unroll
            for (int mm = 0; mm < 4; ++mm) {
                const cp gB = cpmul(lanepow_s[31 - lane][mm], carryB_s[team][buf][0][mm]);
#pragma unroll
                for (int r = 0; r < TC_R; ++r) ypend[r] = re_acc(ld2(P.cB[r][mm]), gB, ypend[r]);
            }
#pragma unroll
            for (int r = 0; r < TC_R; ++r) {
                const long k = pend_k0 + r;
                if (k >= 0 && k < n_dec) out[k] = ypend[r];
            }
            pend = false;
        }

        // ---- carries across passes: causal from the pass before (previous group / analytic at a chunk start),
        // anti-causal from the pass after (analytic at a chunk end; the last pass of any other group waits one group)
        const bool last_group = gi == gpc - 1;
        const bool wait_b = (w == TC_GP - 1) && !last_group;
        float2 c0 = zero2, cM = zero2;
        if (w == 0 && gi == 0) c0 = padded_sample_ool(xc, n, 0);
        if (w == TC_GP - 1 && last_group) cM = padded_sample_ool(xc, n, n + 53);
#pragma unroll
        for (int mm = 0; mm < 4; ++mm) {
            cp inF, inB;
            inF.re = inF.im = inB.re = inB.im = zero2;
            if (w > 0) inF = carryF_s[team][buf][w - 1][mm];
            else if (gi == 0) inF = cp_const(ld2(P.initF[mm]), c0);
            else if (g > gs) inF = carryF_s[team][buf ^ 1][TC_GP - 1][mm];
            if (w < TC_GP - 1) inB = carryB_s[team][buf][w + 1][mm];
            else if (last_group) inB = cp_const(ld2(P.initB[mm]), cM);
            const cp sF = cpfma(lanepow_s[lane][mm], inF, tF[mm]);
            const cp sB = cpfma(lanepow_s[31 - lane][mm], inB, tB[mm]);
#pragma unroll
            for (int r = 0; r < TC_R; ++r) y[r] = re_acc(ld2(P.cB[r][mm]), sB, re_acc(ld2(P.cF[r][mm]), sF, y[r]));
        }
        if (store) {
            if (wait_b) {
                pend = true; pend_k0 = k0; pend_chunk = chunk;
#pragma unroll
                for (int r = 0; r < TC_R; ++r) ypend[r] = y[r];
            } else {
                float2 *out = dec_out + (long)chunk * n_dec;
#pragma unroll
                for (int r = 0; r < TC_R; ++r) {
                    const long k = k0 + r;
                    if (k >= 0 && k < n_dec) out[k] = y[r];
                }
            }
        }
    }
    tick(8);
    if (timed) timing[15] += ge - gs;
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem_base_s, TC_TEAMS * TC_TEAM_COLS);
}

// ====================================================================================================================
// Warp-specialised form ("tc4"): ONE CTA per SM whose warps are the stages of a pipeline over the CTA's range of groups,
//   warp 0        loader        one lane issues the TMA bulk copy of group g into a ring of 3 raw stages
//   warps 2..5    re-layout     slot -> registers -> canonical hi (shared memory, 2 buffers) + lo (tensor memory, 2 buffers)
//   warp 1        MMA issuer    one lane: 3 x K/8 tcgen05.mma into THREE accumulators (hi*Bhi, lo*Bhi, hi*Blo: three dependency
//                               chains of K/8 instead of one of 3K/8), tcgen05.commit to "accumulators full" and "A free"
//   warps 6..     scan teams    NS teams of 4 warps take alternate groups: TMEM load, sum of the three accumulators,
//                               Kogge-Stone scans, carries, stores
// connected by mbarriers (raw full/empty, A full/empty, D full/empty) and, between scan teams, a per-group "carries
// published" barrier: the causal carry of group g-1 and the anti-causal carry of group g+1 come from another team.
// Every wait is bounded: a stage that waits too long raises the CTA's abort flag and all stages drain and exit, so a
// protocol error costs a wrong result (caught by the parity tests), never a hung GPU.
constexpr int TC4_NS = 3;                       // scan teams
constexpr int TC4_NR = 1;                       // re-layout teams (2: team t owns operand buffer t; needs an even TC4_RAW so that a raw
                                                // stage is always read by the same team.  Measured: no faster)
constexpr int TC4_RAW = 3;                      // raw ring stages
constexpr int TC4_SCAN0 = 2 + 4 * TC4_NR;       // first scan warp
constexpr int TC4_WARPS = TC4_SCAN0 + 4 * TC4_NS;
static_assert(TC4_NR == 1 || (TC4_NR == 2 && TC4_RAW % 2 == 0), "two re-layout teams need an even raw ring");
constexpr int TC4_DCOLS = 144;                  // three 48-column accumulators per D buffer
constexpr int TC4_ACOL0 = 2 * TC4_DCOLS;        // lo operand buffers behind the two D buffers (64 columns each)

__device__ __forceinline__ bool mbar_try(uint64_t *bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait: false (and the abort flag set) if the phase did not complete in ~1 s
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t *bar, unsigned parity, volatile int *abort_flag)
{
    for (int spin = 0; spin < (1 << 22); ++spin) {
        if (mbar_try(bar, parity)) return true;
        if ((spin & 1023) == 1023 && *abort_flag) return false;
    }
    *abort_flag = 1;
    return false;
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_sync(int id, int threads)
{
    asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(threads) : "memory");
}

template <int Q, bool TIMED>
__global__ void __launch_bounds__(TC4_WARPS * 32, 1)
decim_tc4_kernel(const float2 *__restrict__ x, long n, long chunk_stride, int n_chunks, long n_dec, int gpc, long total_floats,
                 TcParams P, const float *__restrict__ bmat, float2 *__restrict__ dec_out, int *__restrict__ err_out,
                 long long *__restrict__ timing)
{
    using G = TcG<Q>;
    constexpr int RAW_BYTES = (int)(((127L * G::SLOT_FLOATS + G::K) * 4 + 127) & ~127L);
    extern __shared__ uint8_t tc_raw[];
    uint8_t *sm = tc_raw + ((128u - (smem_u32(tc_raw) & 127u)) & 127u);
    uint8_t *B_hi = sm, *B_lo = B_hi + G::B_BYTES;
    uint8_t *A_can = sm + 2 * G::B_BYTES;                         // 2 canonical hi buffers
    uint8_t *raw = A_can + 2 * G::A_BYTES;                        // TC4_RAW raw stages
    // d_full is per scan TEAM (the issuer commits group `it` to d_full[it % NS]): a team then waits consecutive phases of its
    // own barrier — with a barrier per D buffer a team that skips groups could be two phases early and pass on parity alone
    __shared__ __align__(8) uint64_t raw_full[TC4_RAW], raw_empty[TC4_RAW], a_full[2], a_empty[2], d_full[TC4_NS], d_empty[2], pub_bar[4], b_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ cp pubF[4][4], pubB[4][4];                         // per group slot (g & 3): causal carry of its last pass, anti-causal of its first
    __shared__ cp carryF_s[TC4_NS][TC_GP][4], carryB_s[TC4_NS][TC_GP][4];
    __shared__ cf lanepow_s[32][4];
    __shared__ int abort_s;
    volatile int *abort_flag = &abort_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NG = n_chunks * gpc;
    const int G0 = (int)((long)NG * blockIdx.x / gridDim.x), G1 = (int)((long)NG * (blockIdx.x + 1) / gridDim.x);
    const int gs = G0 - ((G0 % gpc) ? 1 : 0), ge = G1 + ((G1 % gpc) ? 1 : 0);

    if (threadIdx.x < 128) lanepow_s[threadIdx.x & 31][threadIdx.x >> 5] = ld2(P.lanepow[threadIdx.x & 31][threadIdx.x >> 5]);
    if (threadIdx.x == 0) {
        abort_s = 0;
        for (int i = 0; i < TC4_RAW; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&raw_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); mbar_init(&d_empty[i], 1); }
        for (int i = 0; i < TC4_NS; ++i) mbar_init(&d_full[i], 1);
        for (int i = 0; i < 4; ++i) mbar_init(&pub_bar[i], 1);
        mbar_init(&b_bar, 1);
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem0 = tmem_base_s;
    // profiling aid (tb200_debug_tc_timing): per-stage cycles of CTA 0 — one thread per role accumulates where it waits / works
    const bool timed = TIMED && timing != nullptr && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 1 || warp == 2 || warp == TC4_SCAN0);
    long long tprev = timed ? clock64() : 0;
    long long tacc[TIMED ? 12 : 1];                   // kept in registers (constant indices), written once at the end
#pragma unroll
    for (int i = 0; i < (TIMED ? 12 : 1); ++i) tacc[i] = 0;
    auto tick = [&](int c) { if (TIMED && timed) { const long long t = clock64(); tacc[TIMED ? c : 0] += t - tprev; tprev = t; } };

    constexpr long RAW_FLOATS = 127L * G::SLOT_FLOATS + G::K;
    auto group_f0 = [&](int g) -> long {
        const int c = g / gpc, gi = g - c * gpc;
        return 2 * ((long)c * chunk_stride + (long)Q * (gi * TC_GROUP - G::LEAD - 1));
    };
    auto group_tma = [&](int g) -> bool {
        const long f0 = group_f0(g);
        return f0 >= 0 && f0 + RAW_FLOATS <= total_floats;
    };

    if (warp == 0) {
        // ================================================= loader
        if (elect_one()) {
            tma_load_1d(B_hi, bmat, 2u * G::B_BYTES, &b_bar);
            for (int g = gs; g < ge && !*abort_flag; ++g) {
                const int it = g - gs, st = it % TC4_RAW;
                if (!mbar_wait_bounded(&raw_empty[st], ((it / TC4_RAW) & 1) ^ 1, abort_flag)) break;
                tick(0);
                if (group_tma(g))
                    tma_load_1d(raw + st * RAW_BYTES, reinterpret_cast<const float *>(x) + group_f0(g), (unsigned)(RAW_FLOATS * 4), &raw_full[st]);
                else
                    mbar_arrive(&raw_full[st]);                    // the re-layout team fills this group with plain loads
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================================================= MMA issuer
        if (elect_one()) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint64_t desc_bh = umma_desc(smem_u32(B_hi), TC_N * 16, 128), desc_bl = umma_desc(smem_u32(B_lo), TC_N * 16, 128);
            bool ok = mbar_wait_bounded(&b_bar, 0, abort_flag);
            for (int g = gs; g < ge && ok && !*abort_flag; ++g) {
                const int it = g - gs, b = it & 1;
                ok = mbar_wait_bounded(&a_full[b], (it >> 1) & 1, abort_flag) && mbar_wait_bounded(&d_empty[b], ((it >> 1) & 1) ^ 1, abort_flag);
                if (!ok) break;
                tick(4);
                tc_fence_after();
                const uint64_t desc_a = umma_desc(smem_u32(A_can + b * G::A_BYTES), 2048, 128);
                const uint32_t d0 = tmem0 + b * TC4_DCOLS, alo = tmem0 + TC4_ACOL0 + 64 * b;
#pragma unroll
                for (int ks = 0; ks < G::K / 8; ++ks) {
                    const uint64_t dah = desc_a + (uint64_t)(ks * (4096 >> 4));
                    const uint64_t dbh = desc_bh + (uint64_t)(ks * ((2 * TC_N * 16) >> 4)), dbl = desc_bl + (uint64_t)(ks * ((2 * TC_N * 16) >> 4));
                    umma_tf32(d0, dah, dbh, idesc, ks > 0 ? 1u : 0u);
                    umma_tf32_ts(d0 + 48, alo + 8 * ks, dbh, idesc, ks > 0 ? 1u : 0u);
                    umma_tf32(d0 + 96, dah, dbl, idesc, ks > 0 ? 1u : 0u);
                }
                umma_commit(&d_full[it % TC4_NS]);
                umma_commit(&a_empty[b]);
                tick(5);
            }
        }
        __syncwarp();
    } else if (warp < TC4_SCAN0) {
        // ================================================= re-layout teams: 128 threads each, slot m = 32 * (warp & 3) + lane;
        // team t takes the groups of parity t, i.e. always operand buffer t and the raw stages of parity t
        const int rteam = (warp - 2) >> 2;
        const int m = 32 * (warp & 3) + lane;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        bool ok = true;
        for (int g = gs + rteam; g < ge && ok; g += TC4_NR) {
            const int it = g - gs, st = it % TC4_RAW, b = it & 1;
            const int chunk = g / gpc, gi = g - chunk * gpc;
            const float2 *xc = x + (long)chunk * chunk_stride;
            const long k0 = (long)gi * TC_GROUP - G::LEAD + (long)m * TC_R;
            const long i0 = 1 + (long)Q * (k0 - 1);
            const bool via_tma = group_tma(g);
            uint8_t *rw = raw + st * RAW_BYTES;
            ok = mbar_wait_bounded(&raw_full[st], (it / TC4_RAW) & 1, abort_flag);
            if (!ok) break;
            tick(1);
            float4 v[G::KC];
            if (via_tma) {
#pragma unroll
                for (int kc = 0; kc < G::KC; ++kc) v[kc] = *reinterpret_cast<const float4 *>(rw + m * (G::SLOT_FLOATS * 4) + kc * 16);
                if (i0 - 1 < 0 || i0 - 1 + G::K / 2 > n) {
#pragma unroll
                    for (int kc = 0; kc < G::KC; ++kc) {
                        const long i = i0 - 1 + 2 * kc;
                        if (i < 0 || i >= n) { const float2 e = padded_sample_ool(xc, n, 27 + i); v[kc].x = e.x; v[kc].y = e.y; }
                        if (i + 1 < 0 || i + 1 >= n) { const float2 e = padded_sample_ool(xc, n, 28 + i); v[kc].z = e.x; v[kc].w = e.y; }
                    }
                }
            } else {
#pragma unroll
                for (int kc = 0; kc < G::KC; ++kc) {
                    const long i = i0 - 1 + 2 * kc;
                    const float2 e0 = (i >= 0 && i < n) ? xc[i] : padded_sample_ool(xc, n, 27 + i);
                    const float2 e1 = (i + 1 >= 0 && i + 1 < n) ? xc[i + 1] : padded_sample_ool(xc, n, 28 + i);
                    v[kc] = make_float4(e0.x, e0.y, e1.x, e1.y);
                }
            }
            named_sync(1 + rteam, 128);                           // all slots are in registers: the raw stage is free
            if ((warp & 3) == 0 && lane == 0) mbar_arrive(&raw_empty[st]);
            tick(3);
            ok = mbar_wait_bounded(&a_empty[b], ((it >> 1) & 1) ^ 1, abort_flag);      // the MMAs of group g-2 are done with buffer b
            if (!ok) break;
            tick(2);
            tc_fence_after();
            uint8_t *Ah = A_can + b * G::A_BYTES;
            const uint32_t alo = tmem0 + TC4_ACOL0 + 64 * b + lane_base;
#pragma unroll
            for (int q4 = 0; q4 < G::KC / 4; ++q4) {
                float lo[16];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int kc = 4 * q4 + c;
                    float4 hi;
                    hi.x = __uint_as_float((__float_as_uint(v[kc].x) + 0x1000u) & 0xFFFFE000u);
                    hi.y = __uint_as_float((__float_as_uint(v[kc].y) + 0x1000u) & 0xFFFFE000u);
                    hi.z = __uint_as_float((__float_as_uint(v[kc].z) + 0x1000u) & 0xFFFFE000u);
                    hi.w = __uint_as_float((__float_as_uint(v[kc].w) + 0x1000u) & 0xFFFFE000u);
                    lo[4 * c + 0] = v[kc].x - hi.x; lo[4 * c + 1] = v[kc].y - hi.y;
                    lo[4 * c + 2] = v[kc].z - hi.z; lo[4 * c + 3] = v[kc].w - hi.w;
                    *reinterpret_cast<float4 *>(Ah + kc * 2048 + m * 16) = hi;
                }
                tmem_st16(alo + 16 * q4, lo);
            }
            tmem_st_wait();
            fence_async_smem();
            tc_fence_before();
            named_sync(1 + rteam, 128);
            if ((warp & 3) == 0 && lane == 0) mbar_arrive(&a_full[b]);
            tick(3);
        }
    } else {
        // ================================================= scan teams
        const int team = (warp - TC4_SCAN0) >> 2, w = warp & 3;
        const uint32_t lane_base = (uint32_t)(w * 32) << 16;
        const float2 zero2 = make_float2(0.f, 0.f);
        bool ok = true;
        for (int g = gs + team; g < ge && ok; g += TC4_NS) {
            const int it = g - gs, b = it & 1, slot = it & 3;        // carry slots and their barrier phases count from the range start
            const int chunk = g / gpc, gi = g - chunk * gpc;
            const bool store = g >= G0 && g < G1;
            const bool last_group = gi == gpc - 1;
            const float2 *xc = x + (long)chunk * chunk_stride;
            const long k0 = (long)gi * TC_GROUP - G::LEAD + (long)w * TC_PASS + (long)lane * TC_R;
            tick(11);
            ok = mbar_wait_bounded(&d_full[team], (it / TC4_NS) & 1, abort_flag);
            if (!ok) break;
            tick(7);
            tc_fence_after();
            // ---- accumulators of this slot: sum of the three partial products, columns 0..39
            uint32_t va[32], vb[8];
            {
                const uint32_t d0 = tmem0 + b * TC4_DCOLS + lane_base;
                uint32_t t1[32], t2[8];
                tmem_ld32(d0, va); tmem_ld8(d0 + 32, vb);
                tmem_ld32(d0 + 48, t1); tmem_ld8(d0 + 48 + 32, t2);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) va[i] = __float_as_uint(__uint_as_float(va[i]) + __uint_as_float(t1[i]));
#pragma unroll
                for (int i = 0; i < 8; ++i) vb[i] = __float_as_uint(__uint_as_float(vb[i]) + __uint_as_float(t2[i]));
                tmem_ld32(d0 + 96, t1); tmem_ld8(d0 + 96 + 32, t2);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) va[i] = __float_as_uint(__uint_as_float(va[i]) + __uint_as_float(t1[i]));
#pragma unroll
                for (int i = 0; i < 8; ++i) vb[i] = __float_as_uint(__uint_as_float(vb[i]) + __uint_as_float(t2[i]));
            }
            tc_fence_before();
            named_sync(1 + TC4_NR + team, 128);
            if (w == 0 && lane == 0) mbar_arrive(&d_empty[b]);     // the MMA issuer may overwrite this D buffer
            tick(8);
            float2 y[TC_R];
#pragma unroll
            for (int r = 0; r < TC_R; ++r) y[r] = make_float2(__uint_as_float(va[2 * r]), __uint_as_float(va[2 * r + 1]));
            cp tF[4], tB[4];
#pragma unroll
            for (int mm = 0; mm < 4; ++mm) {
                auto col = [&](int cc) -> float { return __uint_as_float(cc < 32 ? va[cc < 32 ? cc : 0] : vb[cc >= 32 ? cc - 32 : 0]); };
                cp a, bb;
                a.re = make_float2(col(6 + 4 * mm), col(7 + 4 * mm));
                a.im = make_float2(col(8 + 4 * mm), col(9 + 4 * mm));
                bb.re = make_float2(col(22 + 4 * mm), col(23 + 4 * mm));
                bb.im = make_float2(col(24 + 4 * mm), col(25 + 4 * mm));
#pragma unroll
                for (int s = 0; s < 5; ++s) {
                    if (s >= fd_nstage(mm)) continue;
                    cp oa, ob;
                    oa.re = shfl2_up(a.re, 1 << s); oa.im = shfl2_up(a.im, 1 << s);
                    ob.re = shfl2_down(bb.re, 1 << s); ob.im = shfl2_down(bb.im, 1 << s);
                    const cf gm = ld2(P.lampow[s][mm]);
                    if (lane >= (1 << s)) a = cpfma(gm, oa, a);
                    if (lane + (1 << s) < 32) bb = cpfma(gm, ob, bb);
                }
                if (lane == 31) { carryF_s[team][w][mm] = a; if (w == TC_GP - 1) pubF[slot][mm] = a; }
                if (lane == 0) { carryB_s[team][w][mm] = bb; if (w == 0) pubB[slot][mm] = bb; }
                tF[mm].re = shfl2_up(a.re, 1); tF[mm].im = shfl2_up(a.im, 1);
                tB[mm].re = shfl2_down(bb.re, 1); tB[mm].im = shfl2_down(bb.im, 1);
                if (lane == 0) tF[mm].re = tF[mm].im = zero2;
                if (lane == 31) tB[mm].re = tB[mm].im = zero2;
            }
            named_sync(1 + TC4_NR + team, 128);                             // this team's carries are in shared memory
            if (w == 0 && lane == 0) mbar_arrive(&pub_bar[slot]);  // ... and published to the other teams
            tick(9);
            // ---- carries across passes
            const bool need_prev = (w == 0) && gi != 0 && g > gs;
            const bool need_next = (w == TC_GP - 1) && !last_group && (g + 1 < ge);
            if (need_prev) ok = mbar_wait_bounded(&pub_bar[(it - 1) & 3], ((it - 1) >> 2) & 1, abort_flag);
            if (need_next && ok) ok = mbar_wait_bounded(&pub_bar[(it + 1) & 3], ((it + 1) >> 2) & 1, abort_flag);
            tick(10);
            float2 c0 = zero2, cM = zero2;
            if (w == 0 && gi == 0) c0 = padded_sample_ool(xc, n, 0);
            if (w == TC_GP - 1 && last_group) cM = padded_sample_ool(xc, n, n + 53);
#pragma unroll
            for (int mm = 0; mm < 4; ++mm) {
                cp inF, inB;
                inF.re = inF.im = inB.re = inB.im = zero2;
                if (w > 0) inF = carryF_s[team][w - 1][mm];
                else if (gi == 0) inF = cp_const(ld2(P.initF[mm]), c0);
                else if (g > gs) inF = pubF[(it - 1) & 3][mm];
                if (w < TC_GP - 1) inB = carryB_s[team][w + 1][mm];
                else if (last_group) inB = cp_const(ld2(P.initB[mm]), cM);
                else if (g + 1 < ge) inB = pubB[(it + 1) & 3][mm];
                const cp sF = cpfma(lanepow_s[lane][mm], inF, tF[mm]);
                const cp sB = cpfma(lanepow_s[31 - lane][mm], inB, tB[mm]);
#pragma unroll
                for (int r = 0; r < TC_R; ++r) y[r] = re_acc(ld2(P.cB[r][mm]), sB, re_acc(ld2(P.cF[r][mm]), sF, y[r]));
            }
            if (store && ok) {
                float2 *out = dec_out + (long)chunk * n_dec;
#pragma unroll
                for (int r = 0; r < TC_R; ++r) {
                    const long k = k0 + r;
                    if (k >= 0 && k < n_dec) out[k] = y[r];
                }
            }
            named_sync(1 + TC4_NR + team, 128);                             // carryF_s / carryB_s of this team are free again
        }
    }
    if (TIMED && timed) {
#pragma unroll
        for (int i = 0; i < (TIMED ? 12 : 1); ++i) if (tacc[i]) timing[i] += tacc[i];
        if (warp == 0) timing[15] += ge - gs;
    }
    __syncthreads();
    if (threadIdx.x == 0 && abort_s && err_out) atomicExch(err_out, 1);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem0, 512);
}

// ---- right-edge transient of scipy's backward pass: y[k] -= 2 Re sum_i kappa_i pc_i lambda_i^(n_dec-1-k) on the last
// FD_NFIX outputs of every chunk (dsp_fast.cu header); kappa_i from F_i[M-1] = sum_k p_i^k ext[M-1-k].  One warp per chunk.
__global__ void __launch_bounds__(32)
decim_tc_edge_kernel(const float2 *__restrict__ x, long n, long chunk_stride, long n_dec, FastDecimParams K, float2 *__restrict__ dec_out)
{
    const int lane = threadIdx.x;
    const float2 *xc = x + (long)blockIdx.x * chunk_stride;
    cp S[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) S[j].re = S[j].im = make_float2(0.f, 0.f);
    const long t_end = n + 53;
    const long t0 = t_end - 32 * FD_EDGE_S + 1 + (long)lane * FD_EDGE_S;
    constexpr int EB = 14;                                  // samples fetched per batch: independent loads in flight
    static_assert(FD_EDGE_S % EB == 0, "edge batch");
    for (int i0 = 0; i0 < FD_EDGE_S; i0 += EB) {
        float2 e[EB];
#pragma unroll
        for (int i = 0; i < EB; ++i) e[i] = padded_sample(xc, n, t0 + i0 + i);
#pragma unroll
        for (int i = 0; i < EB; ++i) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                cp nx = cpmul(ld2(K.p[j]), S[j]);
                nx.re.x += e[i].x; nx.re.y += e[i].y;
                S[j] = nx;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int st = 0; st < 5; ++st) {
            cp o;
            o.re = shfl2_up(S[j].re, 1 << st); o.im = shfl2_up(S[j].im, 1 << st);
            if (lane >= (1 << st)) S[j] = cpfma(ld2(K.pS[st][j]), o, S[j]);
        }
    }
    cp kappa[4];
    {
        const float2 xl = padded_sample(xc, n, t_end);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            cp k;
            k.re = make_float2(K.ka[i][0] * xl.x, K.ka[i][0] * xl.y);
            k.im = make_float2(K.ka[i][1] * xl.x, K.ka[i][1] * xl.y);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                k = cpfma(ld2(K.kM[i][j]), S[j], k);
                cp cj; cj.re = S[j].re; cj.im = make_float2(-S[j].im.x, -S[j].im.y);
                k = cpfma(ld2(K.kMc[i][j]), cj, k);
            }
            // lane 31 holds F_j[M-1]: broadcast its kappa
            kappa[i].re = make_float2(__shfl_sync(0xffffffffu, k.re.x, 31), __shfl_sync(0xffffffffu, k.re.y, 31));
            kappa[i].im = make_float2(__shfl_sync(0xffffffffu, k.im.x, 31), __shfl_sync(0xffffffffu, k.im.y, 31));
        }
    }
    float2 *out = dec_out + (long)blockIdx.x * n_dec;
    for (int mm = lane; mm < FD_NFIX; mm += 32) {
        const long k = n_dec - 1 - mm;
        if (k < 0) break;
        const int a = mm / FD_R, b = mm - a * FD_R;
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            cf g = ld2(K.pc[m]);
#pragma unroll
            for (int sft = 0; sft < 6; ++sft) if (a & (1 << sft)) g = cmulf(g, ld2(K.lampow[sft][m]));
            const cf lam = ld2(K.lam[m]);
            for (int i = 0; i < b; ++i) g = cmulf(g, lam);
            acc = re_acc(g, kappa[m], acc);
        }
        float2 v = out[k];
        v.x -= 2.0f * acc.x; v.y -= 2.0f * acc.y;
        out[k] = v;
    }
}

// ---- host: weights and carry constants --------------------------------------------------------------------------
typedef std::complex<double> cd;
static void putc(float dst[2], cd v) { dst[0] = (float)v.real(); dst[1] = (float)v.imag(); }

static float tf32_round(float v)
{
    uint32_t u;
    memcpy(&u, &v, 4);
    u = (u + 0x1000u) & 0xFFFFE000u;
    memcpy(&v, &u, 4);
    return v;
}

// modal constants (poles p_i, alpha_i of H(z)H(1/z)) of the float32-rounded design, as in make_fast_params
static int tc_modes(const DecimDesign &d, cd p4[4], cd alpha[4])
{
    cd p8[8], r8[8];
    const double g = (double)d.sos32[0];
    for (int s = 0; s < 4; ++s) {
        double a1 = (double)d.sos32[6 * s + 4], a2 = (double)d.sos32[6 * s + 5];
        double disc = a1 * a1 - 4 * a2;
        if (!(disc < 0)) return fail(TB200_E_ARG, "tc decimator: real pole pair in section %d", s);
        p8[s] = cd(-a1 / 2, std::sqrt(-disc) / 2);
        p8[s + 4] = std::conj(p8[s]);
    }
    cd prodp(1.0, 0.0);
    for (int i = 0; i < 8; ++i) prodp *= p8[i];
    const double dd = g / prodp.real();
    for (int i = 0; i < 8; ++i) {
        cd num = g * std::pow(1.0 + 1.0 / p8[i], 8), den(1.0, 0.0);
        for (int j = 0; j < 8; ++j) if (j != i) den *= (1.0 - p8[j] / p8[i]);
        r8[i] = num / den;
    }
    for (int j = 0; j < 4; ++j) {
        cd acc(dd, 0.0);
        for (int i = 0; i < 8; ++i) acc += r8[i] / (1.0 - p8[i] * p8[j]);
        alpha[j] = r8[j] * acc;
        p4[j] = p8[j];
    }
    return TB200_OK;
}

// B[k][n] in double (row-major, K x TC_N), the canonical-layout TF32 split (bhi, blo: [KC][TC_N][4] floats each) and the
// carry constants for decimation factor q.  K_out = contraction length.
static int tc_design(const DecimDesign &d, int q, std::vector<double> &B, std::vector<float> &bhi, std::vector<float> &blo,
                     TcParams &P, int &K_out)
{
    cd p[4], alpha[4];
    int rc = tc_modes(d, p, alpha);
    if (rc) return rc;
    const int KF = 2 + 2 * TC_R * q, K = (KF + 15) / 16 * 16, KC = K / 4;
    K_out = K;
    B.assign((size_t)K * TC_N, 0.0);
    for (int f = 2; f < 2 + 2 * TC_R * q; ++f) {
        const int j = f / 2 - 1, comp = f & 1, s = j / q, c = j % q;
        double *row = &B[(size_t)f * TC_N];
        for (int r = 0; r < TC_R; ++r) {                   // zero-state outputs: h[|t - n_k|]
            const int delta = q * (s - r) - (q - 1) + c;
            double hsum = 0.0;
            for (int i = 0; i < 4; ++i) hsum += 2.0 * (alpha[i] * std::pow(p[i], std::abs(delta))).real();
            row[2 * r + comp] = hsum;
        }
        for (int i = 0; i < 4; ++i) {
            const cd vf = 2.0 * alpha[i] * std::pow(p[i], q * (TC_R - 1 - s) + (q - 1) - c);
            const cd vb = 2.0 * alpha[i] * std::pow(p[i], q * s + c + 1);
            row[6 + 4 * i + 0 + comp] = vf.real();
            row[6 + 4 * i + 2 + comp] = vf.imag();
            row[22 + 4 * i + 0 + comp] = vb.real();
            row[22 + 4 * i + 2 + comp] = vb.imag();
        }
    }
    bhi.assign((size_t)KC * TC_N * 4, 0.f);
    blo.assign((size_t)KC * TC_N * 4, 0.f);
    for (int k = 0; k < K; ++k)
        for (int nn = 0; nn < TC_N; ++nn) {
            const double v = B[(size_t)k * TC_N + nn];
            const float hi = tf32_round((float)v);
            const float lo = tf32_round((float)(v - (double)hi));
            const size_t at = ((size_t)(k / 4) * TC_N + nn) * 4 + (k & 3);
            bhi[at] = hi; blo[at] = lo;
        }
    memset(&P, 0, sizeof(P));
    for (int i = 0; i < 4; ++i) {
        const cd lam = std::pow(p[i], q), lr = std::pow(lam, TC_R);
        cd v = lr;
        int nst = 0;
        for (int s = 0; s < 5; ++s) { putc(P.lampow[s][i], v); if (std::abs(v) >= 1e-9) nst = s + 1; v *= v; }
        if (nst > fd_nstage(i)) return fail(TB200_E_ARG, "tc decimator: mode %d of q=%d needs %d scan stages, kernel has %d", i, q, nst, fd_nstage(i));
        for (int l = 0; l < 32; ++l) putc(P.lanepow[l][i], std::pow(lr, l));
        for (int r = 0; r < TC_R; ++r) { putc(P.cF[r][i], std::pow(lam, r + 1)); putc(P.cB[r][i], std::pow(lam, TC_R - 1 - r)); }
        putc(P.initF[i], 2.0 * alpha[i] / (1.0 - p[i]));
        putc(P.initB[i], 2.0 * alpha[i] * p[i] / (1.0 - p[i]));
    }
    return TB200_OK;
}

int tc_decimator_supported(int q) { return q == 10 || q == 8; }

template <int Q>
static int launch_tc_q(tb200_handle h, const void *x, int n_chunks, long n, long chunk_stride, long n_dec,
                       long total_floats, const TcParams &P, const float *bmat, const FastDecimParams &KE, void *dec_out,
                       cudaStream_t st)
{
    using G = TcG<Q>;
    const size_t smem = TC_TEAMS * (size_t)G::A_BYTES + 2 * (size_t)G::B_BYTES + 128;
    static std::atomic<bool> attr_set[64];
    static std::atomic<int> n_sm[64];
    const int dv = h->cfg.device & 63;
    if (!attr_set[dv]) {
        cudaError_t e = cudaFuncSetAttribute(decim_tc_kernel<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(TB200_E_CUDA, "cudaFuncSetAttribute(decim_tc): %s", cudaGetErrorString(e));
        int sms = 0;
        TB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->cfg.device));
        n_sm[dv] = sms > 0 ? sms : 148;
        attr_set[dv] = true;
    }
    const int gpc = (int)((n_dec + G::LEAD + 4 + TC_GROUP - 1) / TC_GROUP);
    const long NG = (long)n_chunks * gpc;
    if (NG >= (1L << 30)) return fail(TB200_E_ARG, "tc decimator: batch too large");
    if (h->decim_tc >= 2) {
        // warp-specialised pipeline: one CTA per SM
        constexpr int RAW_BYTES = (int)(((127L * G::SLOT_FLOATS + G::K) * 4 + 127) & ~127L);
        const size_t smem4 = 2 * (size_t)G::B_BYTES + 2 * (size_t)G::A_BYTES + (size_t)TC4_RAW * RAW_BYTES + 128;
        static std::atomic<bool> attr4_set[64];
        if (!attr4_set[dv]) {
            cudaError_t e = cudaFuncSetAttribute(decim_tc4_kernel<Q, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(decim_tc4_kernel<Q, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4);
            if (e != cudaSuccess) return fail(TB200_E_CUDA, "cudaFuncSetAttribute(decim_tc4): %s", cudaGetErrorString(e));
            attr4_set[dv] = true;
        }
        long grid4 = n_sm[dv];
        if (grid4 > NG) grid4 = NG;
        int *err = reinterpret_cast<int *>((char *)h->tc_bmat + h->tc_bmat_bytes - 16);      // counter slot 14 of the timing block
        if (h->tc_timing)
            decim_tc4_kernel<Q, true><<<(unsigned)grid4, TC4_WARPS * 32, smem4, st>>>((const float2 *)x, n, chunk_stride, n_chunks, n_dec, gpc,
                                                                                    total_floats, P, bmat, (float2 *)dec_out, err, h->tc_timing);
        else
            decim_tc4_kernel<Q, false><<<(unsigned)grid4, TC4_WARPS * 32, smem4, st>>>((const float2 *)x, n, chunk_stride, n_chunks, n_dec, gpc,
                                                                                     total_floats, P, bmat, (float2 *)dec_out, err, nullptr);
        TB_LAUNCH_CHECK(h);
    } else {
    long grid = 2L * n_sm[dv];                      // two resident CTAs per SM (two teams each), one range of groups per team
    if (grid * TC_TEAMS > NG) grid = NG / TC_TEAMS > 0 ? NG / TC_TEAMS : 1;
    decim_tc_kernel<Q><<<(unsigned)grid, TC_TEAMS * TC_GP * 32, smem, st>>>((const float2 *)x, n, chunk_stride, n_chunks, n_dec, gpc,
                                                               total_floats, P, bmat, (float2 *)dec_out, h->tc_timing);
    TB_LAUNCH_CHECK(h);
    }
    decim_tc_edge_kernel<<<n_chunks, 32, 0, st>>>((const float2 *)x, n, chunk_stride, n_dec, KE, (float2 *)dec_out);
    TB_LAUNCH_CHECK(h);
    return TB200_OK;
}

// eligibility beyond q: the streaming form needs a few groups per chunk and 16-byte aligned slots
bool tc_decimator_usable(const void *x, long n, long chunk_stride, int q)
{
    return tc_decimator_supported(q) && n >= 4L * TC_GROUP * q && n < (1L << 28) && (chunk_stride & 1) == 0 &&
           (reinterpret_cast<uintptr_t>(x) & 15) == 0;
}

int launch_decimate_tc(tb200_handle h, const void *x, int n_chunks, long n, long chunk_stride, int q, const DecimDesign &d,
                       void *dec_out, long n_dec, cudaStream_t st)
{
    if (!tc_decimator_usable(x, n, chunk_stride, q)) return fail(TB200_E_ARG, "tc decimator: unsupported geometry");
    // constants depend on (design, q, n): cached; the weight matrices live in the handle's workspace
    struct Cached { int q; long n, n_dec; float sos[24]; TcParams P; FastDecimParams KE; std::vector<float> bhi, blo; };
    static std::mutex mu;
    static std::vector<Cached> cache;
    TcParams P;
    FastDecimParams KE;
    const Cached *hit = nullptr;
    std::unique_lock<std::mutex> lk(mu);
    for (const Cached &c : cache)
        if (c.q == q && c.n == n && c.n_dec == n_dec && !memcmp(c.sos, d.sos32, sizeof(c.sos))) { hit = &c; break; }
    if (!hit) {
        Cached c;
        c.q = q; c.n = n; c.n_dec = n_dec;
        memcpy(c.sos, d.sos32, sizeof(c.sos));
        std::vector<double> B;
        int K = 0;
        int rc = tc_design(d, q, B, c.bhi, c.blo, c.P, K);
        if (rc) return rc;
        rc = make_fast_params(d, q, n, n_dec, c.KE);
        if (rc) return rc;
        if (cache.size() >= 8) cache.erase(cache.begin());
        cache.push_back(std::move(c));
        hit = &cache.back();
    }
    P = hit->P; KE = hit->KE;
    const size_t bbytes = hit->bhi.size() * sizeof(float);
    if (2 * bbytes + 128 > h->tc_bmat_bytes) return fail(TB200_E_ARG, "tc decimator: weight matrices exceed the workspace");
    if (!(h->tc_key_q == q && !memcmp(h->tc_key_sos, d.sos32, sizeof(h->tc_key_sos)))) {
        // first use of this design on this handle: upload Bhi, Blo (synchronous copies from pageable memory)
        TB_CUDA(cudaMemcpyAsync(h->tc_bmat, hit->bhi.data(), bbytes, cudaMemcpyHostToDevice, st));
        TB_CUDA(cudaMemcpyAsync((char *)h->tc_bmat + bbytes, hit->blo.data(), bbytes, cudaMemcpyHostToDevice, st));
        TB_CUDA(cudaStreamSynchronize(st));
        h->tc_key_q = q;
        memcpy(h->tc_key_sos, d.sos32, sizeof(h->tc_key_sos));
    }
    lk.unlock();
    const long total_floats = 2 * ((long)(n_chunks - 1) * chunk_stride + n);
    if (q == 10) return launch_tc_q<10>(h, x, n_chunks, n, chunk_stride, n_dec, total_floats, P, h->tc_bmat, KE, dec_out, st);
    return launch_tc_q<8>(h, x, n_chunks, n, chunk_stride, n_dec, total_floats, P, h->tc_bmat, KE, dec_out, st);
}

}  // namespace tb200

// Debugging / test aid: the weight matrix and constants of the tensor-core decimator for factor q (host only, no device
// needed): B as K x 48 doubles (row-major), the TF32 split in the canonical layout, TcParams as 4-byte floats.
extern "C" int tb200_debug_tc_design(int q, double *B, int b_capacity, float *bhi, float *blo, int split_capacity, float *params,
                                     int params_capacity, int *K_out)
{
    using namespace tb200;
    if (!tc_decimator_supported(q) || !K_out) return fail(TB200_E_ARG, "tb200_debug_tc_design: q=%d not supported", q);
    DecimDesign d;
    int rc = tb200_design_decimator(q, d.sos64, d.zi64, d.sos32, d.zi32);
    if (rc) return rc;
    std::vector<double> Bv;
    std::vector<float> hi, lo;
    TcParams P;
    int K = 0;
    rc = tc_design(d, q, Bv, hi, lo, P, K);
    if (rc) return rc;
    *K_out = K;
    if (B) { if ((size_t)b_capacity < Bv.size()) return fail(TB200_E_ARG, "tb200_debug_tc_design: B capacity"); memcpy(B, Bv.data(), Bv.size() * sizeof(double)); }
    if (bhi && blo) {
        if ((size_t)split_capacity < hi.size()) return fail(TB200_E_ARG, "tb200_debug_tc_design: split capacity");
        memcpy(bhi, hi.data(), hi.size() * sizeof(float));
        memcpy(blo, lo.data(), lo.size() * sizeof(float));
    }
    if (params) {
        if ((size_t)params_capacity * sizeof(float) < sizeof(P)) return fail(TB200_E_ARG, "tb200_debug_tc_design: params capacity");
        memcpy(params, &P, sizeof(P));
    }
    return TB200_OK;
}

// Profiling aid: with a non-zero `enable` the tensor-core decimator of this handle accumulates, for thread 0 of CTA 0,
// the cycles spent per phase (0 loop top, 1 input wait, 2 slot read, 3 re-layout + split, 4 MMA issue, 5 MMA wait,
// 6 TMEM load, 7 scans, 8 carries + stores, 15 groups); reads and clears the counters.
extern "C" int tb200_debug_tc_timing(tb200_handle h, int enable, int64_t out[16])
{
    using namespace tb200;
    if (!h || !h->tc_bmat) return fail(TB200_E_ARG, "tb200_debug_tc_timing: handle has no process workspace");
    long long *buf = reinterpret_cast<long long *>((char *)h->tc_bmat + h->tc_bmat_bytes - 16 * sizeof(long long));
    TB_CUDA(cudaDeviceSynchronize());
    if (out) {
        if (h->tc_timing) TB_CUDA(cudaMemcpy(out, buf, 16 * sizeof(long long), cudaMemcpyDeviceToHost));
        else memset(out, 0, 16 * sizeof(int64_t));
    }
    TB_CUDA(cudaMemset(buf, 0, 16 * sizeof(long long)));
    h->tc_timing = enable ? buf : nullptr;
    return TB200_OK;
}
