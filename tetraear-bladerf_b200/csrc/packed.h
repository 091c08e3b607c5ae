// packed.h — float complex helpers and packed (I, Q) float2 arithmetic shared by the FAST-mode kernels.
#pragma once
#include <cuda_runtime.h>

namespace tb200 {

struct cf { float x, y; };
__device__ __forceinline__ cf cmulf(cf a, cf b) { return cf{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ __forceinline__ cf cfma(cf a, cf b, cf c) { return cf{fmaf(a.x, b.x, fmaf(-a.y, b.y, c.x)), fmaf(a.x, b.y, fmaf(a.y, b.x, c.y))}; }
__device__ __forceinline__ cf cshfl_up(cf v, int d) { return cf{__shfl_up_sync(0xffffffffu, v.x, d), __shfl_up_sync(0xffffffffu, v.y, d)}; }
__device__ __forceinline__ cf cshfl_down(cf v, int d) { return cf{__shfl_down_sync(0xffffffffu, v.x, d), __shfl_down_sync(0xffffffffu, v.y, d)}; }
__device__ __forceinline__ cf cshfl(cf v, int l) { return cf{__shfl_sync(0xffffffffu, v.x, l), __shfl_sync(0xffffffffu, v.y, l)}; }
__device__ __forceinline__ cf ld2(const float c[2]) { return cf{c[0], c[1]}; }

// ---- packed float2 arithmetic: the I and Q components of a sample share every coefficient, so each
// quantity is kept as a float2 (I, Q) and updated with ONE fma.rn.f32x2 (SASS FFMA2) instead of two FFMAs.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c)
{
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b),
                       rc = *reinterpret_cast<unsigned long long *>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2 *>(&rd);
}
__device__ __forceinline__ float2 splat(float s) { return make_float2(s, s); }
// complex value per component: re = (re_I, re_Q), im = (im_I, im_Q)
struct cp { float2 re, im; };
// c + g * a  for a scalar complex g and packed complex a, c
__device__ __forceinline__ cp cpfma(cf g, cp a, cp c)
{
    cp o;
    o.re = ffma2(splat(g.x), a.re, ffma2(splat(-g.y), a.im, c.re));
    o.im = ffma2(splat(g.x), a.im, ffma2(splat(g.y), a.re, c.im));
    return o;
}
__device__ __forceinline__ cp cpmul(cf g, cp a)
{
    cp z; z.re = z.im = make_float2(0.f, 0.f);
    return cpfma(g, a, z);
}
__device__ __forceinline__ float2 shfl2_up(float2 v, int d) { return make_float2(__shfl_up_sync(0xffffffffu, v.x, d), __shfl_up_sync(0xffffffffu, v.y, d)); }
__device__ __forceinline__ float2 shfl2_down(float2 v, int d) { return make_float2(__shfl_down_sync(0xffffffffu, v.x, d), __shfl_down_sync(0xffffffffu, v.y, d)); }
// y += Re(g * a)
__device__ __forceinline__ float2 re_acc(cf g, cp a, float2 y) { return ffma2(splat(g.x), a.re, ffma2(splat(-g.y), a.im, y)); }

}  // namespace tb200
