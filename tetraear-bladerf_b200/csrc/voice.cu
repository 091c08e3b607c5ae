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
// One CTA per frame, one thread per output short: a pure gather, bound by the 1380-byte store per frame.
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

__global__ void __launch_bounds__(256)
voice_blocks_kernel(int mode, const void *__restrict__ src, long src_stride, const int32_t *__restrict__ n_src,
                    const int32_t *__restrict__ frame_stream, const int32_t *__restrict__ frame_pos, float scale,
                    int16_t *__restrict__ blocks, uint8_t *__restrict__ valid)
{
    const long f = blockIdx.x;
    int16_t *out = blocks + f * 690;
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
    if (threadIdx.x == 0 && valid) valid[f] = ok ? 1 : 0;
    for (int i = threadIdx.x; i < 690; i += blockDim.x) {
        int v = 0;
        if (i == 0) v = ok ? 0x6B21 : 0;
        else if (ok && mode == TB200_VOICE_PARAMS) {
            const int k = (i - 1) / 172, j = (i - 1) - 172 * k;          // sub-frame, offset
            if (k < 4 && j < 95) {
                const uint8_t *sub = row8 + 108 * k;
                if (j >= 3) v = sub[16 + j - 3] ? 300 : -300;
                else {
                    const int lo = j == 0 ? 0 : (j == 1 ? 6 : 13), nb = j == 0 ? 6 : (j == 1 ? 7 : 3);
                    int e = 0;
                    for (int b = 0; b < nb; ++b) e |= (sub[lo + b] ? 1 : 0) << b;
                    // Python: int((e / 63.0) * 1000) - 500 etc. — IEEE double division and product, truncation
                    if (j == 0) v = (int)__dmul_rn(__ddiv_rn((double)e, 63.0), 1000.0) - 500;
                    else if (j == 1) v = (int)__dmul_rn(__ddiv_rn((double)e, 127.0), 200.0) + 20;
                    else v = (int)__dmul_rn(__ddiv_rn((double)e, 7.0), 150.0);
                }
            }
        } else if (ok) {
            const int idx = voice_slot(i);
            if (idx >= 0) {
                if (mode == TB200_VOICE_SYMBOLS) {
                    const int sy = idx < 216 ? (idx >> 1) : 119 + ((idx - 216) >> 1);
                    const int d = row8[sy];
                    v = (((idx & 1) ? d : (d >> 1)) & 1) ? 127 : -127;
                } else if (mode == TB200_VOICE_BITS) {
                    v = row8[idx] ? 127 : -127;
                } else {
                    float t = rintf(rowf[idx] * scale);
                    v = (int)fminf(127.0f, fmaxf(-127.0f, t));
                }
            }
        }
        out[i] = (int16_t)v;
    }
}

int launch_voice_blocks(tb200_handle h, int mode, const void *src, long src_stride, const int32_t *n_src,
                        const int32_t *frame_stream, const int32_t *frame_pos, long n_frames, float scale,
                        int16_t *blocks, uint8_t *valid, cudaStream_t st)
{
    if (n_frames == 0) return TB200_OK;
    voice_blocks_kernel<<<(unsigned)n_frames, 256, 0, st>>>(mode, src, src_stride, n_src, frame_stream, frame_pos, scale,
                                                            blocks, valid);
    TB_LAUNCH_CHECK(h);
    return TB200_OK;
}

}  // namespace tb200
