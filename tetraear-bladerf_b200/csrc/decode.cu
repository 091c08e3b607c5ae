// decode.cu — the integer half of the path: TetraDecoder.symbols_to_bits / find_sync / decode
// (/root/reference/tetraear/core/decoder.py:140-295, :835-992) and
// TetraProtocolParser.parse_burst with its CRC (/root/reference/tetraear/core/protocol.py:192-347).
//
// Design: byte work, HBM bound.  A burst is 255 dibits = 255 bytes: one warp takes one burst,
// each lane loads 8 consecutive dibits (a coalesced 255-byte row per warp) and keeps its 16
// bits in one register; every later step (pattern match, data-bit gather, CRC, header
// fields) is a warp-shuffle exchange on those registers, nothing goes back to memory until
// the 32-byte record is written.  The CRC-16 is linear over GF(2), so instead of a 200-step
// serial shift register each lane XORs the precomputed contribution of its own 1-bits and
// the warp XOR-reduces (5 shuffles) — forward and bit-reversed payload in the same pass.
#include <mutex>
#include "common.h"
#include "kernels.h"
#include "symbols_dev.h"

namespace tb200 {

__constant__ uint8_t c_ts1[22] = {1,1,0,1,0,0,0,0,1,1,1,0,1,0,0,1,1,1,0,1,0,0};   // decoder.py:197
__constant__ uint8_t c_ts2[22] = {0,1,1,1,1,0,1,0,0,1,0,0,0,0,1,1,0,1,1,1,0,0};   // decoder.py:198
// protocol.py:162-163 as 22-bit words, first bit = MSB
#define SYNC_CONT_W 0x343A74u  /* 1101000011101001110100 */
#define SYNC_DISC_W 0x0E90D3u  /* 0011101001000011010011 */

// crc_tab[k] = state reached from e15 after k register steps (k = 1..511); crc_init[n] is
// folded in by the kernels (M^n * 0xFFFF) for the two payload lengths in use.
__device__ uint16_t g_crc_tab[512];
__device__ uint16_t g_crc_init[512];
// CRC contribution of every 4-bit group of BURST bits, per burst type (0 normal, 1 sync): forward-payload CRC in the low
// half-word, reversed-payload CRC in the high one.  Indexed [type][nibble of the lane 0..3][value 0..15][lane]: the 32
// lanes of a warp read 32 consecutive words — one shared-memory wavefront, no bank conflict (nibble j = 4 lane + q).
__device__ uint32_t g_crc_nib[2][4][16][32];
static bool g_crc_ready[64] = {false};
static std::mutex g_crc_mu;                 // tb200_create may run concurrently for several devices

static void build_crc_tables(uint16_t *tab, uint16_t *init)
{
    // one register step without input: M(s)
    auto stepf = [](uint16_t s) -> uint16_t { return (s & 0x8000) ? (uint16_t)((s << 1) ^ 0x1021) : (uint16_t)(s << 1); };
    // contribution of a 1-bit whose own step is the first of k steps: M^k applied to e15 ... the bit is
    // XORed in BEFORE the shift (protocol.py:337-343), so k steps act on 0x8000.
    uint16_t s = 0x8000;
    tab[0] = 0;
    for (int k = 1; k < 512; ++k) { s = stepf(s); tab[k] = s; }
    uint16_t r = 0xFFFF;
    init[0] = r;
    for (int n = 1; n < 512; ++n) { r = stepf(r); init[n] = r; }
}

int ensure_crc_tables(int device)
{
    if (device < 0 || device >= 64) return fail(TB200_E_ARG, "device ordinal out of range");
    std::lock_guard<std::mutex> lk(g_crc_mu);
    if (g_crc_ready[device]) return TB200_OK;
    uint16_t tab[512], init[512];
    build_crc_tables(tab, init);
    TB_CUDA(cudaMemcpyToSymbol(g_crc_tab, tab, sizeof(tab)));
    TB_CUDA(cudaMemcpyToSymbol(g_crc_init, init, sizeof(init)));
    static uint32_t nib[2][4][16][32];
    for (int type = 0; type < 2; ++type) {
        const int n_pay = type ? 494 : 200;
        for (int j = 0; j < 128; ++j)
            for (int v = 0; v < 16; ++v) {
                uint16_t f = 0, r = 0;
                for (int b = 0; b < 4; ++b) {
                    if (!((v >> (3 - b)) & 1)) continue;
                    int bi = 4 * j + b;                                    // burst bit index
                    int d = type ? (bi < 510 ? bi : -1) : (bi < 108 ? bi : (bi >= 122 && bi < 230) ? bi - 14 : -1);
                    if (d >= 0 && d < n_pay) { f ^= tab[n_pay - d]; r ^= tab[d + 1]; }
                }
                nib[type][j & 3][v][j >> 2] = (uint32_t)f | ((uint32_t)r << 16);
            }
    }
    TB_CUDA(cudaMemcpyToSymbol(g_crc_nib, nib, sizeof(nib)));
    g_crc_ready[device] = true;
    return TB200_OK;
}

// ------------------------------------------------------------------ symbols_to_bits
__global__ void symbols_to_bits_kernel(const uint8_t *__restrict__ sym, long stride, const int32_t *__restrict__ n_sym,
                                       int n_fixed, uint8_t *__restrict__ bits, uint8_t *__restrict__ mapped,
                                       long bits_stride)
{
    __shared__ int s_max;
    int s = blockIdx.x;
    long n = n_sym ? n_sym[s] : n_fixed;
    n = n < 0 ? 0 : (n > stride ? stride : n);          // the count lives on the device: never trust it past the row
    const uint8_t *in = sym + (long)s * stride;
    if (threadIdx.x == 0) s_max = 0;
    __syncthreads();
    int mx = 0;
    for (long i = threadIdx.x; i < n; i += blockDim.x) mx = max(mx, (int)in[i]);
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_down_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) atomicMax(&s_max, mx);
    __syncthreads();
    bool dqpsk = s_max <= 3;                              // decoder.py:149-150
    uint8_t *b = bits + (long)s * bits_stride;
    uint8_t *m = mapped ? mapped + (long)s * (bits_stride / 2) : nullptr;
    for (long i = threadIdx.x; i < n; i += blockDim.x) {
        int v = in[i];
        if (dqpsk) v &= 3;
        else v = (v <= 2) ? 0 : (v <= 4) ? 1 : (v == 5) ? 3 : (v <= 7) ? 2 : 0;   // decoder.py:160-164
        if (m) m[i] = (uint8_t)v;
        b[2 * i] = (uint8_t)(v >> 1);
        b[2 * i + 1] = (uint8_t)(v & 1);
    }
}

// ------------------------------------------------------------------ 22-bit Hamming correlation
__global__ void sync_corr_kernel(const uint8_t *__restrict__ bits, long bits_stride, const int32_t *__restrict__ n_sym,
                                 int n_bits_fixed, uint8_t *__restrict__ m1, uint8_t *__restrict__ m2)
{
    int s = blockIdx.y;
    long nb = n_sym ? 2L * n_sym[s] : n_bits_fixed;
    nb = nb < 0 ? 0 : (nb > bits_stride ? bits_stride : nb);
    long nwin = nb - 21;
    long pos = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (pos >= nwin) return;
    const uint8_t *b = bits + (long)s * bits_stride + pos;
    int a1 = 0, a2 = 0;
#pragma unroll
    for (int k = 0; k < 22; ++k) { int v = b[k]; a1 += (v == c_ts1[k]); a2 += (v == c_ts2[k]); }
    m1[(long)s * bits_stride + pos] = (uint8_t)a1;
    m2[(long)s * bits_stride + pos] = (uint8_t)a2;
}

// smallest match count m with (double)m / 22 >= thr  (decoder.py:240,245)
__device__ __forceinline__ int need_matches(double thr)
{
    int m = 0;
    while (m <= 22 && !((double)m / 22.0 >= thr)) ++m;
    return m;
}

// find_sync executed by one full warp over precomputed match counts (decoder.py:226-281).
// Returns the number of positions written (same on every lane); *max_match = best count
// among the (pattern, position) pairs the reference evaluates.
__device__ int warp_find_sync(const uint8_t *__restrict__ m1, const uint8_t *__restrict__ m2, long nwin, double thr,
                              int32_t *__restrict__ pos_out, int max_pos, int *max_match)
{
    const int lane = threadIdx.x & 31;
    int found = 0, mmax = 0;
    if (nwin <= 0) { *max_match = 0; return 0; }
    const int need = need_matches(thr);
    long i = 0;
    while (i < nwin) {
        long p = i + lane;
        int a1 = 0, a2 = 0;
        if (p < nwin) { a1 = m1[p]; a2 = m2[p]; }
        bool hit = (p < nwin) && (a1 >= need || a2 >= need);
        unsigned ball = __ballot_sync(0xffffffffu, hit);
        int first = ball ? (__ffs(ball) - 1) : 32;
        // visited: lanes < first evaluate both patterns; the hit lane stops after TS1 if TS1 hits
        int v = 0;
        if (p < nwin) {
            if (lane < first) v = max(a1, a2);
            else if (lane == first) v = (a1 >= need) ? a1 : max(a1, a2);
        }
        for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
        mmax = max(mmax, v);
        if (ball) {
            if (lane == 0 && found < max_pos) pos_out[found] = (int32_t)(i + first);
            ++found;
            i = i + first + 250;                         // decoder.py:256
        } else {
            i += 32;
        }
    }
    if (found == 0) {
        double max_corr = (double)mmax / 22.0;
        if (max_corr > 0.75 && max_corr >= (thr - 0.15)) {       // decoder.py:265
            double ad = fmax(0.75, max_corr - 0.02);
            if (ad < thr) {
                const int need_ad = need_matches(ad);
                i = 0;
                while (i < nwin) {
                    long p = i + lane;
                    bool hit = (p < nwin) && (max((int)m1[p < nwin ? p : 0], (int)m2[p < nwin ? p : 0]) >= need_ad);
                    unsigned ball = __ballot_sync(0xffffffffu, hit);
                    if (ball) {
                        int first = __ffs(ball) - 1;
                        if (lane == 0 && found < max_pos) pos_out[found] = (int32_t)(i + first);
                        ++found;
                        i = i + first + 250;             // seen_positions window, decoder.py:278
                    } else {
                        i += 32;
                    }
                }
            }
        }
    }
    *max_match = mmax;
    return found;
}

__global__ void find_sync_kernel(const uint8_t *__restrict__ m1, const uint8_t *__restrict__ m2, int n_bits, double thr,
                                 int32_t *__restrict__ pos, int max_pos, int32_t *__restrict__ count,
                                 int32_t *__restrict__ max_match)
{
    int mm;
    int c = warp_find_sync(m1, m2, (long)n_bits - 21, thr, pos, max_pos, &mm);
    if (threadIdx.x == 0) { *count = c; *max_match = mm; }
}

// ------------------------------------------------------------------ burst parsing, warp-resident
struct BurstOut {
    uint32_t burst_type, crc_ok, mac_ok, hdr, sync_match;
    uint32_t mac_fill, mac_length, mac_address;     // fixed-offset MAC header fields (protocol.py:399-421, :510-517)
};

__device__ __forceinline__ unsigned warp_xor(unsigned v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v ^= __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_isum(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// bits i .. i+nb-1 (nb <= 16) of the 510-bit burst held as one 16-bit word per lane (bit 15 first)
__device__ __forceinline__ unsigned burst_field(unsigned w, int i, int nb)
{
    int l0 = i >> 4, o0 = i & 15;
    unsigned hi = __shfl_sync(0xffffffffu, w, l0 & 31);
    unsigned lo = __shfl_sync(0xffffffffu, w, (l0 + 1) & 31);
    unsigned both = (hi << 16) | lo;                       // 32 bits starting at bit 16*l0
    return (both >> (32 - o0 - nb)) & ((1u << nb) - 1u);
}

// packed byte `byte` (8 data bits, MSB first) of TetraBurst.data_bits (protocol.py:277-290):
// normal burst = burst bits [0,108) ++ [122,230) (216 bits), sync burst = all 510 bits; zero padded.
// Every lane of the warp must call this (it shuffles).
__device__ __forceinline__ unsigned data_byte(unsigned w, bool is_sync, int byte)
{
    int d0 = byte * 8;
    unsigned gap = (burst_field(w, 104, 4) << 4) | burst_field(w, 122, 4);   // byte 13 straddles the TS gap
    int src = is_sync ? d0 : (d0 < 104 ? d0 : d0 + 14);
    bool valid = is_sync ? (d0 < 512) : (d0 < 216);
    unsigned plain = burst_field(w, (valid && src <= 504) ? src : 0, 8);
    if (!valid) return 0u;
    if (!is_sync && d0 == 104) return gap;
    return plain;
}

// Parse one burst held in `w` (lane L has burst bits 16L..16L+15; bits >= 510 are zero).
// crc tables in shared memory.  rec (32 B) / full (64 B) may be null.
__device__ void warp_parse_burst(unsigned w, const uint16_t *__restrict__ tab, const uint16_t *__restrict__ initv,
                                 uint8_t *__restrict__ rec, uint8_t *__restrict__ full, BurstOut &out)
{
    const int lane = threadIdx.x & 31;
    // _detect_burst_type: bits[255:277] vs both patterns, max/22 > 0.8 (protocol.py:246-265)
    unsigned win = (burst_field(w, 255, 11) << 11) | burst_field(w, 266, 11);
    int mc = 22 - __popc(win ^ SYNC_CONT_W), md = 22 - __popc(win ^ SYNC_DISC_W);
    int mx = max(mc, md);
    bool is_sync = mx >= 18;                          // max(match) / 22 > 0.8 (protocol.py:252): 18/22 = 0.818, 17/22 = 0.773
    // per-lane view of the data bits: normal = [0,108) ++ [122,230), sync = all 510
    int ones = 0;
    unsigned cf = 0, cr = 0;           // forward / reversed-payload CRC contributions
    const int n_data = is_sync ? 510 : 216;
    const int n_pay = n_data - 16;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        int bi = lane * 16 + k;                            // burst bit index
        unsigned bit = (w >> (15 - k)) & 1u;
        int d;                                             // data-bit index or -1
        if (is_sync) d = (bi < 510) ? bi : -1;
        else d = (bi < 108) ? bi : (bi >= 122 && bi < 230) ? bi - 14 : -1;
        if (d >= 0 && bit) {
            ++ones;
            if (d < n_pay) { cf ^= tab[n_pay - d]; cr ^= tab[d + 1]; }
        }
    }
    ones = warp_isum(ones);
    cf = warp_xor(cf) ^ initv[n_pay];
    cr = warp_xor(cr) ^ initv[n_pay];
    unsigned rx = is_sync ? burst_field(w, 494, 16) : burst_field(w, 214, 16);
    bool crc_ok = false;
    if (ones != 0 && ones != n_data) {                     // protocol.py:301-304
        crc_ok = (__popc(cf ^ rx) <= 2) || (__popc(cr ^ rx) <= 2);   // :311-325
    }
    // MAC header validity (parse_mac_pdu returns None?) protocol.py:369-526 — data bits 0..34 = burst bits 0..34
    unsigned f0 = burst_field(w, 0, 16), f1 = burst_field(w, 16, 16), f2 = burst_field(w, 32, 16);
    unsigned long long head = ((unsigned long long)f0 << 32) | ((unsigned long long)f1 << 16) | f2;   // 48 bits
    auto fld = [&](int lo, int nb) -> unsigned { return (unsigned)((head >> (48 - lo - nb)) & ((1ull << nb) - 1ull)); };
    unsigned type = fld(0, 2), enc = fld(2, 2);
    bool mac_ok;
    if (type == 0) { unsigned len = fld(29, 6); mac_ok = !((int)(len * 8) > n_data - 35 + 16); }
    else if (type == 1) mac_ok = true;
    else if (type == 2) {
        mac_ok = true;
        if (enc == 0) { unsigned mcc = fld(4, 10), mnc = fld(14, 14); mac_ok = !(mcc < 200 || mcc > 799 || mnc > 999); }
    } else { unsigned len = fld(5, 6); mac_ok = !((int)(len * 8) > n_data - 11 + 16); }
    out.burst_type = is_sync ? 5u : 2u;
    out.crc_ok = crc_ok;
    out.mac_ok = mac_ok;
    out.hdr = (type << 2) | enc;
    out.sync_match = (unsigned)mx;
    out.mac_fill = (type != 2) ? fld(4, 1) : 0u;
    out.mac_address = (type == 0) ? fld(5, 24) : 0u;
    out.mac_length = (type == 0) ? fld(29, 6) : (type == 3) ? fld(5, 6) : 0u;
    if (rec) {
        // bytes 0..26: first 216 data bits packed
        unsigned byte = data_byte(w, is_sync, lane);
        if (lane < 27) rec[lane] = (uint8_t)byte;
        if (lane == 27) rec[27] = (uint8_t)out.burst_type;
        if (lane == 28) rec[28] = (uint8_t)out.crc_ok;
        if (lane == 29) rec[29] = (uint8_t)out.mac_ok;
        if (lane == 30) rec[30] = (uint8_t)out.hdr;
        if (lane == 31) rec[31] = (uint8_t)out.sync_match;
    }
    if (full) {
        full[2 * lane] = (uint8_t)(w >> 8);
        full[2 * lane + 1] = (uint8_t)(w & 0xff);
    }
}

// Payload statistics of the MAC PDU parse_mac_pdu would build from the data bits (protocol.py:349-537): where
// MacPDU.data starts and how many bits it takes per PDU type, its length in bytes (BitArray.tobytes pads the last byte
// with zeros), the number of distinct byte values (the entropy heuristic of decoder.py:1036-1053 divides it by the
// length) and its first two bytes (the protocol-id tests in front of parse_sds_data, protocol.py:820-900).
// Every lane of the warp must call this.  w = burst word of the lane, o = the parsed header.
struct MacPayload { uint32_t n_bytes, n_unique, b0, b1; };
__device__ MacPayload warp_mac_payload(unsigned w, const BurstOut &o)
{
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    const bool is_sync = o.burst_type == 5;
    const int n_data = is_sync ? 510 : 216;
    const unsigned type = o.hdr >> 2;
    int pos, nbits;
    if (type == 0) { pos = 35; const int lb = 8 * (int)o.mac_length; nbits = (lb > 0 && pos + lb <= n_data) ? lb : n_data - pos; }
    else if (type == 1) { pos = 5; nbits = n_data - pos; }
    else if (type == 2) { pos = 4; nbits = n_data - pos; }
    else { pos = 11; const int lb = 8 * (int)o.mac_length; nbits = (lb > 0 && pos + lb <= n_data) ? lb : n_data - pos; }
    MacPayload r = {0u, 0u, 0u, 0u};
    if (!o.mac_ok) nbits = 0;                                   // parse_mac_pdu returned None: no PDU, no data
    r.n_bytes = (uint32_t)((nbits + 7) >> 3);
    const int sh = pos & 7;
    unsigned pres[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int j = lane + 32 * half;                         // payload byte index of this lane
        const int a = (pos >> 3) + j;                           // aligned data byte that holds its first bit
        const unsigned hi = data_byte(w, is_sync, a < 64 ? a : 63), lo = data_byte(w, is_sync, a + 1 < 64 ? a + 1 : 63);
        unsigned v = ((((a < 64 ? hi : 0u) << 8) | (a + 1 < 64 ? lo : 0u)) >> (8 - sh)) & 0xFFu;
        const int left = nbits - 8 * j;                         // payload bits from this byte on
        const bool present = left > 0;
        if (left < 8 && left > 0) v &= (0xFFu << (8 - left)) & 0xFFu;      // zero padding of the last byte
        if (!present) v = 0u;
        if (j == 0) r.b0 = v;
        if (j == 1) r.b1 = v;
#pragma unroll
        for (int k = 0; k < 8; ++k) pres[k] |= (present && (int)(v >> 5) == k) ? (1u << (v & 31u)) : 0u;
    }
    unsigned uniq = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) uniq += __popc(__reduce_or_sync(FULL, pres[k]));
    r.n_unique = uniq;
    r.b0 = __shfl_sync(FULL, r.b0, 0);
    r.b1 = __shfl_sync(FULL, r.b1, 1);
    return r;
}

// lane's 16 bits of a burst given its 255 dibits (each dibit -> (s>>1)&1, s&1: protocol.py:211-214)
__device__ __forceinline__ unsigned load_burst_word(const uint8_t *__restrict__ dib)
{
    const int lane = threadIdx.x & 31;
    unsigned w = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        int idx = lane * 8 + k;
        unsigned s = (idx < 255) ? dib[idx] : 0u;
        w = (w << 2) | (s & 3u);
    }
    return w;
}

// same for a row known to live in a buffer padded to a multiple of 8 bytes: two aligned 64-bit loads and
// a funnel shift instead of eight byte loads
__device__ __forceinline__ unsigned load_burst_word_fast(const uint8_t *__restrict__ base, long row_byte0, long total_bytes)
{
    const int lane = threadIdx.x & 31;
    long a = row_byte0 + lane * 8;
    long a0 = a & ~7L;
    int sh = (int)(a & 7) * 8;
    unsigned long long v;
    if (a0 + 16 <= total_bytes) {
        const unsigned long long *p = reinterpret_cast<const unsigned long long *>(base + a0);
        unsigned long long lo = p[0], hi = p[1];
        v = sh ? ((lo >> sh) | (hi << (64 - sh))) : lo;                     // bytes a .. a+7, little endian
    } else {                                                                // last words of the buffer: stay in bounds
        v = 0;
        for (int k = 0; k < 8; ++k) if (a + k < total_bytes) v |= (unsigned long long)base[a + k] << (8 * k);
    }
    if (lane == 31) v &= 0x00FFFFFFFFFFFFFFull;                            // dibit 255 does not exist
    v &= 0x0303030303030303ull;
    // byte k (dibit k) -> bits 15-2k..14-2k: reverse the bytes (dibit 0 on top), then gather the eight 2-bit fields of
    // each 32-bit half by folding (|, >> 6) and (|, >> 12)
    const unsigned hi = __byte_perm((unsigned)v, 0, 0x0123), lo = __byte_perm((unsigned)(v >> 32), 0, 0x0123);
    auto gather4 = [](unsigned z) -> unsigned { const unsigned t = (z | (z >> 6)) & 0x000F000Fu; return (t | (t >> 12)) & 0xFFu; };
    return (gather4(hi) << 8) | gather4(lo);
}

// Leaner burst parse for the batch kernel: nibble CRC tables, hardware warp reductions, direct shuffles.
__device__ __forceinline__ void warp_parse_burst_v2(unsigned w, const uint32_t (*__restrict__ nibtab)[4][16][32],
                                                    const uint16_t *__restrict__ initv, uint8_t *__restrict__ rec,
                                                    uint8_t *__restrict__ full, BurstOut &out)
{
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    // bits 255..276: lane 15 bit 0, lane 16 all, lane 17 top 5
    unsigned w15 = __shfl_sync(FULL, w, 15), w16 = __shfl_sync(FULL, w, 16), w17 = __shfl_sync(FULL, w, 17);
    unsigned win = ((w15 & 1u) << 21) | (w16 << 5) | (w17 >> 11);
    int mx = max(22 - __popc(win ^ SYNC_CONT_W), 22 - __popc(win ^ SYNC_DISC_W));
    const bool is_sync = mx >= 18;                    // max(match) / 22 > 0.8 (protocol.py:252): 18/22 = 0.818, 17/22 = 0.773
    const int type = is_sync ? 1 : 0;
    const int n_data = is_sync ? 510 : 216, n_pay = n_data - 16;
    // data-bit mask of this lane's 16 burst bits (normal: [0,108) ++ [122,230); sync: [0,510))
    unsigned dmask;
    {
        int lo = lane * 16;
        if (is_sync) dmask = (lane == 31) ? 0xFFFCu : 0xFFFFu;
        else {
            dmask = 0;
            if (lo + 16 <= 108) dmask = 0xFFFFu;
            else if (lo < 108) dmask = (0xFFFFu << (lo + 16 - 108)) & 0xFFFFu;         // lane 6: bits 96..107
            if (lo >= 122 && lo + 16 <= 230) dmask = 0xFFFFu;
            else if (lo < 122 && lo + 16 > 122) dmask |= 0xFFFFu >> (122 - lo);          // lane 7: bits 122..127
            else if (lo < 230 && lo + 16 > 230) dmask = (0xFFFFu << (lo + 16 - 230)) & 0xFFFFu;   // lane 14: 224..229
        }
    }
    int ones = __reduce_add_sync(FULL, __popc(w & dmask));
    unsigned c = 0;
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
        unsigned v = (w >> (12 - 4 * q4)) & 15u;
        c ^= nibtab[type][q4][v][lane];
    }
    c = __reduce_xor_sync(FULL, c);
    unsigned cf = (c & 0xFFFFu) ^ initv[n_pay], cr = (c >> 16) ^ initv[n_pay];
    unsigned w13 = __shfl_sync(FULL, w, 13), w14 = __shfl_sync(FULL, w, 14);
    unsigned w30 = __shfl_sync(FULL, w, 30), w31 = __shfl_sync(FULL, w, 31);
    unsigned rx = is_sync ? (((w30 & 3u) << 14) | (w31 >> 2)) : (((w13 & 0x3FFu) << 6) | (w14 >> 10));
    bool crc_ok = false;
    if (ones != 0 && ones != n_data) crc_ok = (__popc(cf ^ rx) <= 2) || (__popc(cr ^ rx) <= 2);
    unsigned f0 = __shfl_sync(FULL, w, 0), f1 = __shfl_sync(FULL, w, 1), f2 = __shfl_sync(FULL, w, 2);
    unsigned long long head = ((unsigned long long)f0 << 32) | ((unsigned long long)f1 << 16) | f2;
    auto fld = [&](int lo, int nb) -> unsigned { return (unsigned)((head >> (48 - lo - nb)) & ((1ull << nb) - 1ull)); };
    unsigned type2 = fld(0, 2), enc = fld(2, 2);
    bool mac_ok;
    if (type2 == 0) mac_ok = !((int)(fld(29, 6) * 8) > n_data - 35 + 16);
    else if (type2 == 1) mac_ok = true;
    else if (type2 == 2) { mac_ok = true; if (enc == 0) { unsigned mcc = fld(4, 10), mnc = fld(14, 14); mac_ok = !(mcc < 200 || mcc > 799 || mnc > 999); } }
    else mac_ok = !((int)(fld(5, 6) * 8) > n_data - 11 + 16);
    out.burst_type = is_sync ? 5u : 2u; out.crc_ok = crc_ok; out.mac_ok = mac_ok; out.hdr = (type2 << 2) | enc;
    out.sync_match = (unsigned)mx;
    out.mac_fill = (type2 != 2) ? fld(4, 1) : 0u;
    out.mac_address = (type2 == 0) ? fld(5, 24) : 0u;
    out.mac_length = (type2 == 0) ? fld(29, 6) : (type2 == 3) ? fld(5, 6) : 0u;
    // record: 27 packed data bytes + 5 flag bytes, one byte per lane -> the warp writes 32 contiguous bytes
    unsigned byte;
    {
        int d0 = lane * 8;
        int src = is_sync ? d0 : (d0 < 104 ? d0 : d0 + 14);
        src = (lane < 27) ? src : 0;
        unsigned hi = __shfl_sync(FULL, w, (src >> 4) & 31), lo = __shfl_sync(FULL, w, ((src >> 4) + 1) & 31);
        unsigned both = (hi << 16) | lo;
        byte = (both >> (24 - (src & 15))) & 0xFFu;
        unsigned w6 = __shfl_sync(FULL, w, 6), w7 = __shfl_sync(FULL, w, 7);
        if (!is_sync && lane == 13) byte = (((w6 >> 4) & 0xFu) << 4) | ((w7 >> 2) & 0xFu);   // bits 104..107 ++ 122..125
        if (lane == 27) byte = out.burst_type;
        if (lane == 28) byte = out.crc_ok;
        if (lane == 29) byte = out.mac_ok;
        if (lane == 30) byte = out.hdr;
        if (lane == 31) byte = out.sync_match;
    }
    rec[lane] = (uint8_t)byte;
    if (full) { full[2 * lane] = (uint8_t)(w >> 8); full[2 * lane + 1] = (uint8_t)(w & 0xff); }
}

__global__ void __launch_bounds__(256)
burst_batch_kernel(const uint8_t *__restrict__ dibits, long n, uint8_t *__restrict__ rec,
                   uint8_t *__restrict__ full, unsigned long long *__restrict__ stats)
{
    __shared__ uint32_t nibtab[2][4][16][32];
    __shared__ uint16_t initv[512];
    {
        const uint32_t *src = &g_crc_nib[0][0][0][0];
        uint32_t *dst = &nibtab[0][0][0][0];
        for (int i = threadIdx.x; i < 2 * 4 * 16 * 32; i += blockDim.x) dst[i] = src[i];
        for (int i = threadIdx.x; i < 512; i += blockDim.x) initv[i] = g_crc_init[i];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    long warp = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    unsigned pass = 0, total = 0;
    const bool aligned = ((reinterpret_cast<uintptr_t>(dibits) & 7) == 0);
    const long total_bytes = n * 255;
    for (long b = warp; b < n; b += nwarps) {
        unsigned w = aligned ? load_burst_word_fast(dibits, b * 255, total_bytes) : load_burst_word(dibits + b * 255);
        BurstOut o;
        warp_parse_burst_v2(w, nibtab, initv, rec + b * 32, full ? full + b * 64 : nullptr, o);
        ++total;
        pass += o.crc_ok;
    }
    if (stats && lane == 0 && total) {
        atomicAdd(&stats[0], (unsigned long long)total);
        atomicAdd(&stats[1], (unsigned long long)pass);
        atomicAdd(&stats[2], (unsigned long long)(total - pass));
    }
}

// ------------------------------------------------------------------ burst batch, one THREAD per burst
// The warp-per-burst form above spends ~85 % of its instructions on scalars every lane recomputes (type, CRC compare,
// header fields, record assembly) and on the shuffles / reductions that hand them round.  Here a CTA takes 128 bursts:
//   phase 1 (cooperative): the tile's 128 x 255 contiguous bytes come in with 16-byte loads (16 in flight per thread), are
//     cut into 8-byte pieces (two aligned 64-bit shared loads + funnel shift) and packed 2 bits per dibit into shared
//     memory, 16 words per burst, [word][burst] with pitch 129;
//   phase 2: thread t owns burst t — its 510 bits in 16 registers, everything else is plain per-thread code: the
//     22-bit sync window, the data-bit popcount, both CRCs from the position-dependent nibble table (re-laid
//     [type][position][value] so that the 32 lanes of a warp, same position / different value, hit different banks),
//     header fields, the 32-byte record as two 16-byte stores (a warp writes 1 KB contiguous).
// Bit-identical to the warp form (tests/test_gpu_parity.py, test_gpu_fuzz.py, the 1M-burst test).
constexpr int BT_THREADS = 128;
constexpr int BT_PITCH = BT_THREADS + 1;

__device__ __forceinline__ unsigned pack8_dibits(unsigned long long v)
{
    v &= 0x0303030303030303ull;
    const unsigned hi = __byte_perm((unsigned)v, 0, 0x0123), lo = __byte_perm((unsigned)(v >> 32), 0, 0x0123);
    auto gather4 = [](unsigned z) -> unsigned { const unsigned t = (z | (z >> 6)) & 0x000F000Fu; return (t | (t >> 12)) & 0xFFu; };
    return (gather4(hi) << 8) | gather4(lo);
}

// bits [start, start + len) of the burst (bit 0 = MSB of W[0]), len <= 32, as the low `len` bits
__device__ __forceinline__ unsigned burst_bits(const unsigned (&W)[16], int start, int len)
{
    const int q = start >> 5, sh = start & 31;
    const unsigned hi = W[q], lo = (q + 1 < 16) ? W[q + 1] : 0u;
    const unsigned x = __funnelshift_l(lo, hi, sh);              // bits start .. start+31
    return len == 32 ? x : (x >> (32 - len));
}

__global__ void __launch_bounds__(BT_THREADS)
burst_batch_tpb_kernel(const uint8_t *__restrict__ dibits, long n, uint8_t *__restrict__ rec,
                       uint8_t *__restrict__ full, unsigned long long *__restrict__ stats)
{
    __shared__ uint32_t nib2[2][128][16];            // [type][16-bit lane * 4 + q4][value]: (forward | reversed << 16) CRC contribution
    __shared__ uint16_t initv[512];
    extern __shared__ __align__(16) uint8_t raw[];               // [BT_THREADS * 255 + 32] the tile's dibits as they lie in memory
    __shared__ __align__(4) uint16_t Wsm[16 * 2 * BT_PITCH];     // halfword g of burst t at (g >> 1) * 2 * PITCH + 2 t + (1 - (g & 1))  (word q = hi:lo)
    {
        for (int i = threadIdx.x; i < 2 * 128 * 16; i += blockDim.x) {
            const int v = i & 15, pos = (i >> 4) & 127, type = i >> 11;
            nib2[type][pos][v] = g_crc_nib[type][pos & 3][v][pos >> 2];
        }
        for (int i = threadIdx.x; i < 512; i += blockDim.x) initv[i] = g_crc_init[i];
    }
    const bool aligned16 = ((reinterpret_cast<uintptr_t>(dibits) & 15) == 0);
    unsigned pass = 0, total = 0;
    for (long blk0 = (long)blockIdx.x * BT_THREADS; blk0 < n; blk0 += (long)gridDim.x * BT_THREADS) {
        __syncthreads();                              // tables ready / previous tile consumed
        const int nb = (int)((n - blk0 < BT_THREADS) ? (n - blk0) : BT_THREADS);
        // ---- phase 1a: the tile's bytes (contiguous: bursts are 255-byte rows) into shared memory, 16 bytes per load, all of a
        // thread's loads in flight before the first store
        const long tile_byte0 = blk0 * 255;
        const int tile_bytes = nb * 255;
        if (aligned16) {
            const uint4 *src = reinterpret_cast<const uint4 *>(dibits + tile_byte0);
            uint4 *dst = reinterpret_cast<uint4 *>(raw);
            const int n16 = tile_bytes >> 4;
            constexpr int PER = (BT_THREADS * 255 / 16 + BT_THREADS - 1) / BT_THREADS;     // 16, in two rounds of 8
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint4 tmp[PER / 2];
#pragma unroll
                for (int i = 0; i < PER / 2; ++i) { const int j = threadIdx.x + (h * (PER / 2) + i) * BT_THREADS; if (j < n16) tmp[i] = src[j]; }
#pragma unroll
                for (int i = 0; i < PER / 2; ++i) { const int j = threadIdx.x + (h * (PER / 2) + i) * BT_THREADS; if (j < n16) dst[j] = tmp[i]; }
            }
            for (int j = (n16 << 4) + threadIdx.x; j < tile_bytes; j += BT_THREADS) raw[j] = dibits[tile_byte0 + j];
        } else {
            for (int j = threadIdx.x; j < tile_bytes; j += BT_THREADS) raw[j] = dibits[tile_byte0 + j];
        }
        for (int j = tile_bytes + threadIdx.x; j < tile_bytes + 16; j += BT_THREADS) raw[j] = 0;      // the last piece reads past the tile
        __syncthreads();
        // ---- phase 1b: pack 8 dibits -> 16 bits (two aligned 64-bit shared loads + funnel shift per piece)
        for (int task = threadIdx.x; task < nb * 32; task += BT_THREADS) {
            const int t = task >> 5, g = task & 31;
            const int a = t * 255 + g * 8, a0 = a & ~7, sh = (a & 7) * 8;
            const unsigned long long lo = *reinterpret_cast<const unsigned long long *>(raw + a0);
            const unsigned long long hi = *reinterpret_cast<const unsigned long long *>(raw + a0 + 8);
            unsigned long long v = sh ? ((lo >> sh) | (hi << (64 - sh))) : lo;
            if (g == 31) v &= 0x00FFFFFFFFFFFFFFull;  // dibit 255 does not exist
            Wsm[(g >> 1) * 2 * BT_PITCH + 2 * t + (1 - (g & 1))] = (uint16_t)pack8_dibits(v);
        }
        __syncthreads();
        // ---- phase 2: one burst per thread
        const int t = threadIdx.x;
        if (t < nb) {
            unsigned W[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) W[q] = reinterpret_cast<const uint32_t *>(Wsm)[q * BT_PITCH + t];
            // bits 255..276 against the two sync patterns (protocol.py:246-265)
            const unsigned win = ((W[7] & 1u) << 21) | (W[8] >> 11);
            const int mx = max(22 - __popc(win ^ SYNC_CONT_W), 22 - __popc(win ^ SYNC_DISC_W));
            const bool is_sync = mx >= 18;
            const int type = is_sync ? 1 : 0;
            const int n_data = is_sync ? 510 : 216, n_pay = n_data - 16;
            // data bits: normal [0,108) ++ [122,230); sync [0,510)
            int ones;
            if (is_sync) {
                ones = __popc(W[15] & 0xFFFFFFFCu);
#pragma unroll
                for (int q = 0; q < 15; ++q) ones += __popc(W[q]);
            } else {
                ones = __popc(W[0]) + __popc(W[1]) + __popc(W[2]) + __popc(W[3] & 0xFFF0003Fu) + __popc(W[4]) + __popc(W[5]) +
                       __popc(W[6]) + __popc(W[7] & 0xFC000000u);
            }
            // CRC-16 of the payload, forward and over the reversed payload, from per-position nibble contributions
            unsigned c = 0;
            const uint32_t (*nt)[16] = nib2[type];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                if (q >= 8 && !is_sync) break;         // normal: nibbles beyond bit 255 contribute nothing
                const unsigned w = W[q];
#pragma unroll
                for (int k = 0; k < 8; ++k) c ^= nt[8 * q + k][(w >> (28 - 4 * k)) & 15u];
            }
            const unsigned cf = (c & 0xFFFFu) ^ initv[n_pay], cr = (c >> 16) ^ initv[n_pay];
            const unsigned rx = is_sync ? ((W[15] >> 2) & 0xFFFFu) : (((W[6] & 0x3FFu) << 6) | (W[7] >> 26));
            bool crc_ok = false;
            if (ones != 0 && ones != n_data) crc_ok = (__popc(cf ^ rx) <= 2) || (__popc(cr ^ rx) <= 2);
            const unsigned long long head = ((unsigned long long)W[0] << 16) | (W[1] >> 16);
            auto fld = [&](int lo, int nbits) -> unsigned { return (unsigned)((head >> (48 - lo - nbits)) & ((1ull << nbits) - 1ull)); };
            const unsigned type2 = fld(0, 2), enc = fld(2, 2);
            bool mac_ok;
            if (type2 == 0) mac_ok = !((int)(fld(29, 6) * 8) > n_data - 35 + 16);
            else if (type2 == 1) mac_ok = true;
            else if (type2 == 2) { mac_ok = true; if (enc == 0) { unsigned mcc = fld(4, 10), mnc = fld(14, 14); mac_ok = !(mcc < 200 || mcc > 799 || mnc > 999); } }
            else mac_ok = !((int)(fld(5, 6) * 8) > n_data - 11 + 16);
            // record: 27 packed data bytes (data bits 0..215) + burst type, crc_ok, mac_ok, header, sync match
            unsigned R[8];
            if (is_sync) {
#pragma unroll
                for (int q = 0; q < 7; ++q) R[q] = W[q];
            } else {
                R[0] = W[0]; R[1] = W[1]; R[2] = W[2];
                // data bits 96..107 = burst bits 96..107, data bits 108.. = burst bits 122..
                R[3] = (W[3] & 0xFFF00000u) | burst_bits(W, 122, 20);
#pragma unroll
                for (int q = 4; q < 7; ++q) R[q] = burst_bits(W, 32 * q + 14, 32);
            }
            // bytes 24..26 = data bits 192..215 (top 24 bits of R[6]); bytes 27..31 = flags
            const unsigned btype = is_sync ? 5u : 2u, hdr = (type2 << 2) | enc;
            R[6] = (R[6] & 0xFFFFFF00u) | btype;
            R[7] = ((crc_ok ? 1u : 0u) << 24) | ((mac_ok ? 1u : 0u) << 16) | (hdr << 8) | (unsigned)mx;
            uint4 *ro = reinterpret_cast<uint4 *>(rec + (blk0 + t) * 32);
            ro[0] = make_uint4(__byte_perm(R[0], 0, 0x0123), __byte_perm(R[1], 0, 0x0123), __byte_perm(R[2], 0, 0x0123), __byte_perm(R[3], 0, 0x0123));
            ro[1] = make_uint4(__byte_perm(R[4], 0, 0x0123), __byte_perm(R[5], 0, 0x0123), __byte_perm(R[6], 0, 0x0123), __byte_perm(R[7], 0, 0x0123));
            if (full) {
                uint4 *fo = reinterpret_cast<uint4 *>(full + (blk0 + t) * 64);
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4)
                    fo[q4] = make_uint4(__byte_perm(W[4 * q4], 0, 0x0123), __byte_perm(W[4 * q4 + 1], 0, 0x0123),
                                        __byte_perm(W[4 * q4 + 2], 0, 0x0123), __byte_perm(W[4 * q4 + 3], 0, 0x0123));
            }
            ++total;
            pass += crc_ok ? 1u : 0u;
        }
    }
    if (stats) {
        total = __reduce_add_sync(0xffffffffu, total);
        pass = __reduce_add_sync(0xffffffffu, pass);
        if ((threadIdx.x & 31) == 0 && total) {
            atomicAdd(&stats[0], (unsigned long long)total);
            atomicAdd(&stats[1], (unsigned long long)pass);
            atomicAdd(&stats[2], (unsigned long long)(total - pass));
        }
    }
}

// ------------------------------------------------------------------ TetraDecoder.decode, one warp per stream
// decode() of ONE stream by one full warp: b = bit row, mp = mapped symbols, c1/c2 = match counts
// (global or shared memory), pos = scratch for max_frames sync positions.
__device__ void decode_stream_warp(const uint8_t *b, const uint8_t *mp, const uint8_t *c1, const uint8_t *c2, long nsym,
                                   const uint16_t *tab, const uint16_t *initv, tb200_frame *__restrict__ fo, int max_frames,
                                   int32_t *__restrict__ n_frames_out, unsigned long long *__restrict__ stats, int32_t *pos)
{
    const int lane = threadIdx.x & 31;
    long nb = 2 * nsym;
    long nwin = nb - 21;
    // threshold cascade, decoder.py:845-856
    int mm;
    int cnt = warp_find_sync(c1, c2, nwin, 0.90, pos, max_frames, &mm);
    if (cnt == 0) {
        cnt = warp_find_sync(c1, c2, nwin, 0.85, pos, max_frames, &mm);
        if (cnt == 0) {
            cnt = warp_find_sync(c1, c2, nwin, 0.80, pos, max_frames, &mm);
            double mc = (double)mm / 22.0;
            if (cnt == 0 && mc >= 0.75) cnt = warp_find_sync(c1, c2, nwin, fmax(0.75, mc - 0.02), pos, max_frames, &mm);
        }
    }
    if (cnt > max_frames) cnt = max_frames;
    __syncwarp();
    unsigned total = 0, pass = 0;
    for (int f = 0; f < cnt; ++f) {
        int p = pos[f];
        long start_pos = (long)p - 216;
        long start_sym = start_pos >= 0 ? start_pos / 2 : 0;
        // decoder.py:867-871 bounds, and decode_frame's len(bits) < 510 -> None (:894)
        bool ok = start_pos >= 0 && (start_sym + 255 <= nsym) && (start_pos + 510 <= nb);
        tb200_frame *fr = fo + f;
        BurstOut o = {0, 0, 0, 0, 0, 0, 0, 0};
        MacPayload mp_stats = {0u, 0u, 0u, 0u};
        unsigned fw = 0;
        if (ok) {
            // frame bits come from the BIT stream at start_pos, burst symbols from the SYMBOL stream at
            // start_pos // 2 — they differ by one bit when start_pos is odd (decoder.py:870-877); keep that.
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                long bi = start_pos + lane * 16 + k;
                unsigned bit = (lane * 16 + k < 510) ? b[bi] : 0u;
                fw = (fw << 1) | (bit & 1u);
            }
            unsigned w = load_burst_word(mp + start_sym);
            warp_parse_burst(w, tab, initv, nullptr, nullptr, o);
            mp_stats = warp_mac_payload(w, o);
            // TetraBurst.data_bits packed into fr->data (216 or 510 bits)
            bool is_sync = o.burst_type == 5;
            for (int r = 0; r < 2; ++r) fr->data[r * 32 + lane] = (uint8_t)data_byte(w, is_sync, r * 32 + lane);
            fr->bits[2 * lane] = (uint8_t)(fw >> 8);
            fr->bits[2 * lane + 1] = (uint8_t)(fw & 0xff);
            ++total;
            pass += o.crc_ok;
        }
        if (lane == 0) {
            fr->sync_pos = p;
            fr->start_pos = (int32_t)start_pos;
            fr->number = ok ? (int32_t)(start_pos / 510) : 0;
            fr->type = (uint8_t)((fw >> 14) & 3u);          // frame bits [0:2], decoder.py:906
            fr->enc_mode = (uint8_t)((fw >> 12) & 3u);      // frame bits [2:4], decoder.py:909
            fr->burst_type = (uint8_t)o.burst_type;
            fr->crc_ok = (uint8_t)o.crc_ok;
            fr->mac_ok = (uint8_t)o.mac_ok;
            fr->kept = (uint8_t)(ok && (o.crc_ok || o.mac_ok));   // decoder.py:1092-1100
            fr->mac_fill = (uint8_t)o.mac_fill;
            fr->mac_length = (uint8_t)o.mac_length;
            fr->mac_address = o.mac_address;
            fr->mac_data_bytes = (uint8_t)mp_stats.n_bytes;
            fr->mac_unique = (uint8_t)mp_stats.n_unique;
            fr->mac_b0 = (uint8_t)mp_stats.b0;
            fr->mac_b1 = (uint8_t)mp_stats.b1;
        }
    }
    if (lane == 0) {
        *n_frames_out = cnt;
        if (stats && total) {
            atomicAdd(&stats[0], (unsigned long long)total);
            atomicAdd(&stats[1], (unsigned long long)pass);
            atomicAdd(&stats[2], (unsigned long long)(total - pass));
        }
    }
}

__global__ void decode_kernel(const uint8_t *__restrict__ bits, const uint8_t *__restrict__ mapped,
                              const uint8_t *__restrict__ m1, const uint8_t *__restrict__ m2, int n_streams,
                              long bits_stride, const int32_t *__restrict__ n_sym, tb200_frame *__restrict__ frames,
                              int max_frames, int32_t *__restrict__ n_frames, unsigned long long *__restrict__ stats,
                              int32_t *__restrict__ pos_scratch)
{
    __shared__ uint16_t tab[512];
    __shared__ uint16_t initv[512];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) { tab[i] = g_crc_tab[i]; initv[i] = g_crc_init[i]; }
    __syncthreads();
    int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (s >= n_streams) return;
    decode_stream_warp(bits + (long)s * bits_stride, mapped + (long)s * (bits_stride / 2), m1 + (long)s * bits_stride,
                       m2 + (long)s * bits_stride, min(max((long)n_sym[s], 0L), bits_stride / 2), tab, initv,
                       frames + (long)s * max_frames, max_frames,
                       n_frames + s, stats, pos_scratch + (long)s * max_frames);
}

// Fused variant for short streams (one chunk of process() output): symbols_to_bits, both 22-bit
// correlations and decode() in ONE launch, every intermediate in shared memory.  One CTA per stream.
__device__ __forceinline__ void decode_fused_body(int s, const uint8_t *__restrict__ dibits, long stride, int n_sym_s,
                                                  tb200_frame *__restrict__ frames, int max_frames, int32_t *__restrict__ n_frames,
                                                  unsigned long long *__restrict__ stats)
{
    extern __shared__ uint8_t dsm[];
    __shared__ uint16_t tab[512];
    __shared__ uint16_t initv[512];
    __shared__ int s_max;
    __shared__ int32_t pos_sm[64];
    const long n = min(max((long)n_sym_s, 0L), stride);       // clamp: the shared buffers hold `stride` dibits
    uint8_t *mp = dsm;                       // [stride]
    uint8_t *b = mp + stride;                // [2 stride]
    uint8_t *c1 = b + 2 * stride;            // [2 stride]
    uint8_t *c2 = c1 + 2 * stride;           // [2 stride]
    uint32_t *W = reinterpret_cast<uint32_t *>(dsm + ((7 * stride + 3) & ~3L));     // bits packed 32 per word, LSB first, + 2 zero words
    for (int i = threadIdx.x; i < 512; i += blockDim.x) { tab[i] = g_crc_tab[i]; initv[i] = g_crc_init[i]; }
    if (threadIdx.x == 0) s_max = 0;
    __syncthreads();
    const uint8_t *in = dibits + (long)s * stride;
    int mx = 0;
    for (long i = threadIdx.x; i < n; i += blockDim.x) { int v = in[i]; mp[i] = (uint8_t)v; mx = max(mx, v); }
    mx = __reduce_max_sync(0xffffffffu, mx);
    if ((threadIdx.x & 31) == 0) atomicMax(&s_max, mx);
    __syncthreads();
    const bool dqpsk = s_max <= 3;
    for (long i = threadIdx.x; i < n; i += blockDim.x) {
        int v = mp[i];
        if (dqpsk) v &= 3;
        else v = (v <= 2) ? 0 : (v <= 4) ? 1 : (v == 5) ? 3 : (v <= 7) ? 2 : 0;
        mp[i] = (uint8_t)v;
        b[2 * i] = (uint8_t)(v >> 1);
        b[2 * i + 1] = (uint8_t)(v & 1);
    }
    __syncthreads();
    // both 22-bit correlations from packed bits: the bits here are 0/1 by construction, so "matches" = 22 - popcount of the
    // window XOR the pattern (8 instructions per position instead of 22 byte compares x 2)
    const int nW = (int)((2 * n + 31) / 32) + 2;
    for (int j = threadIdx.x; j < nW; j += blockDim.x) {
        uint32_t w = 0;
#pragma unroll
        for (int d = 0; d < 16; ++d) {
            const long i = 16L * j + d;
            const uint32_t v = i < n ? mp[i] : 0u;
            w |= ((v >> 1) & 1u) << (2 * d) | (v & 1u) << (2 * d + 1);
        }
        W[j] = w;
    }
    uint32_t M1 = 0, M2 = 0;
#pragma unroll
    for (int k = 0; k < 22; ++k) { M1 |= (uint32_t)c_ts1[k] << k; M2 |= (uint32_t)c_ts2[k] << k; }
    __syncthreads();
    const long nwin = 2 * n - 21;
    for (long p = threadIdx.x; p < nwin; p += blockDim.x) {
        const uint32_t win = __funnelshift_r(W[p >> 5], W[(p >> 5) + 1], (unsigned)(p & 31)) & 0x3FFFFFu;
        c1[p] = (uint8_t)(22 - __popc(win ^ M1));
        c2[p] = (uint8_t)(22 - __popc(win ^ M2));
    }
    __syncthreads();
    if (threadIdx.x < 32)
        decode_stream_warp(b, mp, c1, c2, n, tab, initv, frames + (long)s * max_frames, max_frames < 64 ? max_frames : 64,
                           n_frames + s, stats, pos_sm);
}

__global__ void __launch_bounds__(128)
decode_fused_kernel(const uint8_t *__restrict__ dibits, int n_streams, long stride, const int32_t *__restrict__ n_sym,
                    tb200_frame *__restrict__ frames, int max_frames, int32_t *__restrict__ n_frames,
                    unsigned long long *__restrict__ stats)
{
    (void)n_streams;
    decode_fused_body(blockIdx.x, dibits, stride, n_sym[blockIdx.x], frames, max_frames, n_frames, stats);
}

// The whole tail of process() + decode() for one chunk in ONE launch: timing pick, symbols and
// dibits (symbols_dev.h), then — same CTA, the dibits it just wrote — bits, both correlations, find_sync cascade and
// the frame records.  One CTA of 256 threads per chunk.
__global__ void __launch_bounds__(256)
tail_fused_kernel(const double *__restrict__ filtered, long n, long in_stride, int sps, int step, double *__restrict__ symbols,
                  uint8_t *__restrict__ dibits, long out_stride, int32_t *__restrict__ n_dibits, int32_t *__restrict__ phase_out,
                  const double *__restrict__ phase_partial, const double *__restrict__ phase_tail, int n_tiles,
                  tb200_frame *__restrict__ frames, int max_frames, int32_t *__restrict__ n_frames,
                  unsigned long long *__restrict__ stats)
{
    const int chunk = blockIdx.x;
    symbols_body(chunk, filtered, n, in_stride, sps, step, symbols, dibits, out_stride, n_dibits, nullptr, phase_out,
                 phase_partial, phase_tail, n_tiles);
    __syncthreads();                                   // the chunk's dibits and their count were written by this CTA
    decode_fused_body(chunk, dibits, out_stride, n_dibits[chunk], frames, max_frames, n_frames, stats);
}

// _calculate_crc16 of bit rows (one byte per bit), one warp per row
__global__ void crc16_rows_kernel(const uint8_t *__restrict__ bits, int n_rows, int n_bits, int row_stride,
                                  uint16_t *__restrict__ crc)
{
    int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    const uint8_t *b = bits + (long)row * row_stride;
    // general length: serial shift register per row chunk is fine here (stage-level API, not the hot path)
    unsigned c = 0xFFFF;
    if (lane == 0) {
        for (int i = 0; i < n_bits; ++i) {
            c ^= ((unsigned)(b[i] & 1)) << 15;
            c = (c & 0x8000) ? ((c << 1) ^ 0x1021) : (c << 1);
            c &= 0xFFFF;
        }
        crc[row] = (uint16_t)c;
    }
}

// ------------------------------------------------------------------ launchers
int launch_symbols_to_bits(tb200_handle h, const uint8_t *sym, int n_streams, long stride, const int32_t *n_sym,
                           int n_fixed, uint8_t *bits, uint8_t *mapped, long bits_stride, cudaStream_t st)
{
    if (n_streams <= 0) return TB200_OK;
    symbols_to_bits_kernel<<<n_streams, 256, 0, st>>>(sym, stride, n_sym, n_fixed, bits, mapped, bits_stride);
    TB_LAUNCH_CHECK(h);
    return TB200_OK;
}

int launch_sync_corr(tb200_handle h, const uint8_t *bits, int n_streams, long bits_stride, const int32_t *n_sym,
                     int n_bits_fixed, uint8_t *m1, uint8_t *m2, cudaStream_t st)
{
    long maxwin = bits_stride;   // upper bound of windows per stream
    if (n_streams <= 0 || maxwin <= 0) return TB200_OK;
    dim3 grid((unsigned)((maxwin + 255) / 256), (unsigned)n_streams);
    sync_corr_kernel<<<grid, 256, 0, st>>>(bits, bits_stride, n_sym, n_bits_fixed, m1, m2);
    TB_LAUNCH_CHECK(h);
    return TB200_OK;
}

int launch_find_sync(tb200_handle h, const uint8_t *m1, const uint8_t *m2, int n_bits, double thr, int32_t *pos,
                     int max_pos, int32_t *count, int32_t *max_match, cudaStream_t st)
{
    find_sync_kernel<<<1, 32, 0, st>>>(m1, m2, n_bits, thr, pos, max_pos, count, max_match);
    TB_LAUNCH_CHECK(h);
    return TB200_OK;
}

int launch_decode(tb200_handle h, const uint8_t *bits, const uint8_t *mapped, const uint8_t *m1, const uint8_t *m2,
                  int n_streams, long bits_stride, const int32_t *n_sym, tb200_frame *frames, int max_frames,
                  int32_t *n_frames, int64_t *stats, cudaStream_t st)
{
    if (n_streams <= 0) return TB200_OK;
    int threads = 128;
    int blocks = (n_streams * 32 + threads - 1) / threads;
    decode_kernel<<<blocks, threads, 0, st>>>(bits, mapped, m1, m2, n_streams, bits_stride, n_sym, frames, max_frames,
                                              n_frames, (unsigned long long *)stats, h->sync_pos);
    TB_LAUNCH_CHECK(h);
    return TB200_OK;
}

int launch_decode_fused(tb200_handle h, const uint8_t *dibits, int n_streams, long stride, const int32_t *n_sym,
                        tb200_frame *frames, int max_frames, int32_t *n_frames, int64_t *stats, cudaStream_t st)
{
    if (n_streams <= 0) return TB200_OK;
    size_t smem = ((7 * (size_t)stride + 3) & ~(size_t)3) + 4 * ((2 * (size_t)stride + 31) / 32 + 2);      // + packed bits
    decode_fused_kernel<<<n_streams, 128, smem, st>>>(dibits, n_streams, stride, n_sym, frames, max_frames, n_frames,
                                                     (unsigned long long *)stats);
    TB_LAUNCH_CHECK(h);
    return TB200_OK;
}

int launch_tail_fused(tb200_handle h, const double *filtered, int n_chunks, long n, long in_stride, int sps, int phase_step,
                      double *symbols, uint8_t *dibits, long out_stride, int32_t *n_dibits, int32_t *phase,
                      const double *phase_partial, const double *phase_tail, int n_tiles, tb200_frame *frames, int max_frames,
                      int32_t *n_frames, int64_t *stats, cudaStream_t st)
{
    if (n_chunks <= 0) return TB200_OK;
    size_t smem = ((7 * (size_t)out_stride + 3) & ~(size_t)3) + 4 * ((2 * (size_t)out_stride + 31) / 32 + 2);
    tail_fused_kernel<<<n_chunks, 256, smem, st>>>(filtered, n, in_stride, sps, phase_step, symbols, dibits, out_stride, n_dibits,
                                                   phase, phase_partial, phase_tail, n_tiles, frames, max_frames, n_frames,
                                                   (unsigned long long *)stats);
    TB_LAUNCH_CHECK(h);
    return TB200_OK;
}

int launch_burst_batch(tb200_handle h, const uint8_t *dibits, int64_t n, uint8_t *rec, uint8_t *full, int64_t *stats,
                       cudaStream_t st)
{
    if (n <= 0) return TB200_OK;
    int threads = 256;
    long warps_needed = n;
    long blocks = (warps_needed * 32 + threads - 1) / threads;
    long cap = 148L * 8 * 4;                  // persistent-ish: 8 CTAs/SM x 4 waves
    if (blocks > cap) blocks = cap;
    // thread-per-burst form: needs 16-byte aligned outputs (two / four 16-byte stores per record)
    if (((reinterpret_cast<uintptr_t>(rec) & 15) == 0) && (!full || (reinterpret_cast<uintptr_t>(full) & 15) == 0)) {
        long tiles = (n + BT_THREADS - 1) / BT_THREADS;
        long grid = tiles < 148L * 8 ? tiles : 148L * 8;
        const size_t smem_bt = BT_THREADS * 255 + 32;          // + 25 KB static: above the 48 KB default, opt in
        cudaError_t e = cudaFuncSetAttribute(burst_batch_tpb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        if (e != cudaSuccess) return fail(TB200_E_CUDA, "cudaFuncSetAttribute(burst_batch): %s", cudaGetErrorString(e));
        burst_batch_tpb_kernel<<<(unsigned)grid, BT_THREADS, smem_bt, st>>>(dibits, n, rec, full, (unsigned long long *)stats);
        TB_LAUNCH_CHECK(h);
        return TB200_OK;
    }
    burst_batch_kernel<<<(unsigned)blocks, threads, 0, st>>>(dibits, n, rec, full, (unsigned long long *)stats);
    TB_LAUNCH_CHECK(h);
    return TB200_OK;
}

int launch_crc16_rows(tb200_handle h, const uint8_t *bits, int n_rows, int n_bits, int row_stride, uint16_t *crc,
                      cudaStream_t st)
{
    if (n_rows <= 0) return TB200_OK;
    int threads = 128;
    int blocks = (n_rows * 32 + threads - 1) / threads;
    crc16_rows_kernel<<<blocks, threads, 0, st>>>(bits, n_rows, n_bits, row_stride, crc);
    TB_LAUNCH_CHECK(h);
    return TB200_OK;
}

}  // namespace tb200
