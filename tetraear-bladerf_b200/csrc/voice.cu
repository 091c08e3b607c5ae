// voice.cu — codec-frame builder on the output side of the path.
//
// The reference hands its ACELP decoder 690 int16 per frame: [0] = 0x6B21, 432 soft values in the runs
// 1..114, 116..229, 231..344, 346..435, zeros elsewhere.  It fills them three ways:
//   SYMBOLS  /root/reference/tetraear/ui/modern.py:2309-2423: slot = dibits[pos/3 .. +255) of the chunk's symbol
//            stream (no frame when it does not fit), symbols 0..107 and 119..226, MSB first, bit -> +-127
//   BITS     /root/reference/continuous_capture.py:100-124: the frame's first 432 bits -> +-127
//   PARAMS   /root/reference/tetraear/ui/modern.py:2143-2200: four 108-bit sub-frames at 1 + 172 k:
//            int(e/63*1000)-500, int(p/127*200)+20, int(g/7*150) (LSB-first fields of 6, 7, 3 bits), 92 x +-300
// SOFT (no reference counterpart; "etsi" mode): 432 real soft bits (+ = 1) -> clamp(rint(scale * s), +-127).
#include "common.h"
#include "kernels.h"

namespace tb200 {

// index into the 432 soft values of output short i, -1 for header / padding
__device__ __forceinline__ int voice_slot(int i)
{
    if (i >= 1 && i < 115) return i - 1;
    if (i >= 116 && i < 230) return i - 116 + 114;
    if (i >= 231 && i < 345) return i - 231 + 228;
    if (i >= 346 && i < 436) return i - 346 + 342;
    return -1;
}

// output short i of frame f (header, gathered soft value or padding)
__device__ __forceinline__ int voice_value(int mode, int i, int idx, bool ok, const uint8_t *__restrict__ row8,
                                           const float *__restrict__ rowf, float scale)
{
    if (i == 0) return ok ? 0x6B21 : 0;
    if (!ok) return 0;
    if (mode == TB200_VOICE_PARAMS) {
        const int k = (i - 1) / 172, j = (i - 1) - 172 * k;          // sub-frame, offset
        if (!(k < 4 && j < 95)) return 0;
        const uint8_t *sub = row8 + 108 * k;
        if (j >= 3) return sub[16 + j - 3] ? 300 : -300;
        const int lo = j == 0 ? 0 : (j == 1 ? 6 : 13), nb = j == 0 ? 6 : (j == 1 ? 7 : 3);
        int e = 0;
        for (int b = 0; b < nb; ++b) e |= (sub[lo + b] ? 1 : 0) << b;
        // Python: int((e / 63.0) * 1000) - 500 etc. — IEEE double division and product, truncation
        if (j == 0) return (int)__dmul_rn(__ddiv_rn((double)e, 63.0), 1000.0) - 500;
        if (j == 1) return (int)__dmul_rn(__ddiv_rn((double)e, 127.0), 200.0) + 20;
        return (int)__dmul_rn(__ddiv_rn((double)e, 7.0), 150.0);
    }
    if (idx < 0) return 0;
    if (mode == TB200_VOICE_SYMBOLS) {
        const int sy = idx < 216 ? (idx >> 1) : 119 + ((idx - 216) >> 1);
        const int d = row8[sy];
        return (((idx & 1) ? d : (d >> 1)) & 1) ? 127 : -127;
    }
    if (mode == TB200_VOICE_BITS) return row8[idx] ? 127 : -127;
    const float t = rintf(rowf[idx] * scale);
    return (int)fminf(127.0f, fmaxf(-127.0f, t));
}

// One WARP per frame, a lane writes two adjacent shorts per store (690 shorts = 345 words, frames are 4-byte aligned);
// the slot of every output index comes from a table a CTA builds once.  (Round 1 ran one 256-thread CTA per frame
// — a million CTAs that each decoded the slot ranges with a chain of branches: 1052 warp instructions per frame.)
__global__ void __launch_bounds__(256)
voice_blocks_kernel(int mode, const void *__restrict__ src, long src_stride, const int32_t *__restrict__ n_src,
                    const int32_t *__restrict__ frame_stream, const int32_t *__restrict__ frame_pos, float scale,
                    int16_t *__restrict__ blocks, uint8_t *__restrict__ valid, long n_frames)
{
    __shared__ int16_t slot[692];
    for (int i = threadIdx.x; i < 692; i += blockDim.x) slot[i] = i < 690 ? (int16_t)voice_slot(i) : (int16_t)-1;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((long)gridDim.x * blockDim.x) >> 5;
    for (long f = warp; f < n_frames; f += n_warps) {
        const uint8_t *row8 = nullptr;
        const float *rowf = nullptr;
        bool ok;
        if (mode == TB200_VOICE_SYMBOLS) {
            const int s = frame_stream ? frame_stream[f] : 0;
            const int sp = frame_pos[f] / 3;                       // bit position -> "symbol" position, modern.py:2325
            const int len = n_src ? n_src[s] : (int)src_stride;
            ok = frame_pos[f] >= 0 && sp + 255 <= len;
            row8 = (const uint8_t *)src + (long)s * src_stride + sp;
        } else if (mode == TB200_VOICE_SOFT) {
            ok = (n_src ? n_src[f] : (int)src_stride) >= 432;
            rowf = (const float *)src + f * src_stride;
        } else {
            ok = (n_src ? n_src[f] : (int)src_stride) >= 432;
            row8 = (const uint8_t *)src + f * src_stride;
        }
        if (lane == 0 && valid) valid[f] = ok ? 1 : 0;
        uint32_t *out = reinterpret_cast<uint32_t *>(blocks + f * 690);
        for (int wi = lane; wi < 345; wi += 32) {
            const int i = 2 * wi;
            const int v0 = voice_value(mode, i, slot[i], ok, row8, rowf, scale);
            const int v1 = voice_value(mode, i + 1, slot[i + 1], ok, row8, rowf, scale);
            out[wi] = ((uint32_t)v0 & 0xFFFFu) | ((uint32_t)v1 << 16);
        }
    }
}

int launch_voice_blocks(tb200_handle h, int mode, const void *src, long src_stride, const int32_t *n_src,
                        const int32_t *frame_stream, const int32_t *frame_pos, long n_frames, float scale,
                        int16_t *blocks, uint8_t *valid, cudaStream_t st)
{
    if (n_frames == 0) return TB200_OK;
    if (reinterpret_cast<uintptr_t>(blocks) & 3) return fail(TB200_E_ARG, "tb200_voice_blocks: d_blocks must be 4-byte aligned");
    long ctas = (n_frames + 7) / 8;                               // 8 warps = 8 frames per CTA pass
    if (ctas > 148L * 32) ctas = 148L * 32;
    voice_blocks_kernel<<<(unsigned)ctas, 256, 0, st>>>(mode, src, src_stride, n_src, frame_stream, frame_pos, scale,
                                                        blocks, valid, n_frames);
    TB_LAUNCH_CHECK(h);
    return TB200_OK;
}

}  // namespace tb200
