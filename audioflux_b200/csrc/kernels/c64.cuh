// c64.cuh -- complex fp32 values packed in one 64-bit register pair, operated on with the
// sm_100 packed-fp32 instructions (PTX add/sub/mul/fma.rn.f32x2 -> SASS FADD2 / FMUL2 / FFMA2).
// ptxas folds the half swaps / sign flips below into operand modifiers (.LO_HI, .NP), so a complex
// add is one instruction and a complex multiply by a constant is two.
#pragma once
#include <stdint.h>

typedef unsigned long long c64;     // lo 32 bits = real, hi 32 bits = imaginary

__device__ __forceinline__ c64 c_pack(float re, float im) { c64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(re), "f"(im)); return r; }
__device__ __forceinline__ void c_unpack(c64 v, float &re, float &im) { asm("mov.b64 {%0, %1}, %2;" : "=f"(re), "=f"(im) : "l"(v)); }
__device__ __forceinline__ c64 c_from(float2 v) { return c_pack(v.x, v.y); }
__device__ __forceinline__ float c_re(c64 v) { float a, b; c_unpack(v, a, b); return a; }
__device__ __forceinline__ float c_im(c64 v) { float a, b; c_unpack(v, a, b); return b; }

__device__ __forceinline__ c64 c_add(c64 a, c64 b) { c64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ c64 c_sub(c64 a, c64 b) { c64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ c64 v_mul(c64 a, c64 b) { c64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }   // element-wise
__device__ __forceinline__ c64 v_fma(c64 a, c64 b, c64 c) { c64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

__device__ __forceinline__ c64 c_swap(c64 v) { float a, b; c_unpack(v, a, b); return c_pack(b, a); }
__device__ __forceinline__ c64 c_conj(c64 v) { float a, b; c_unpack(v, a, b); return c_pack(a, -b); }
__device__ __forceinline__ c64 c_mul_mi(c64 v) { float a, b; c_unpack(v, a, b); return c_pack(b, -a); }   // * (-i)
__device__ __forceinline__ c64 c_mul_pi(c64 v) { float a, b; c_unpack(v, a, b); return c_pack(-b, a); }   // * (+i)
__device__ __forceinline__ c64 c_scale(c64 v, float s) { return v_mul(v, c_pack(s, s)); }
// v * (c - i s)  (forward twiddle e^{-i theta}, c = cos theta, s = sin theta)
__device__ __forceinline__ c64 c_mul_cs(c64 v, float c, float s) { return v_fma(c_swap(v), c_pack(s, -s), v_mul(v, c_pack(c, c))); }
// general complex product a * w
__device__ __forceinline__ c64 c_mul(c64 a, c64 w) {
    float wr, wi; c_unpack(w, wr, wi);
    return v_fma(c_swap(a), c_pack(-wi, wi), v_mul(a, c_pack(wr, wr)));
}
// |v|^2
__device__ __forceinline__ float c_norm2(c64 v) { float a, b; c_unpack(v_mul(v, v), a, b); return a + b; }
