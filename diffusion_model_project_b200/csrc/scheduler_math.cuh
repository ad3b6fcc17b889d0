// Device-side arithmetic of the DDPM / DDIM sampler update, shared by the stand-alone kernel (scheduler.cu) and the
// fused final_conv epilogue (conv_common.cuh, out_mode 3): Philox4x32-10 + Box-Muller noise and the per-element step
// in the reference's operation order (Diffusion_model/src/diffusion.py:103-125, :152-188, :195-234).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace b2d {

// ---- Philox4x32-10 (Salmon et al. 2011), counter = element index / 4, key = (seed, step row) ----
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__device__ __forceinline__ void philox4x32_10(uint64_t ctr, uint32_t stream_id, uint64_t seed, uint32_t (&out)[4]) {
  uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), stream_id, 0u};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& z0, float& z1) {
  const float u1 = ((float)(a >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0,1)
  const float u2 = ((float)(b >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float r = sqrtf(-2.0f * __logf(u1));
  float s, c;
  __sincosf(6.283185307179586f * u2, &s, &c);
  z0 = r * c;
  z1 = r * s;
}

// two fp32 -> one packed 16-bit pair: IEEE fp16 (saturating) or bf16, round to nearest
__device__ __forceinline__ uint32_t pack16(float a, float b, int f16) {
  uint32_t r;
  if (f16) {
    asm("{\n\t.reg .b16 lo, hi;\n\tcvt.rn.satfinite.f16.f32 lo, %1;\n\tcvt.rn.satfinite.f16.f32 hi, %2;\n\tmov.b32 %0, {lo, hi};\n\t}" : "=r"(r) : "f"(a), "f"(b));
  } else {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    r = *reinterpret_cast<uint32_t*>(&h);
  }
  return r;
}

struct Coef { float a, b, c1, c2, s; double inv_a; };

__device__ __forceinline__ Coef load_coef(const float* __restrict__ coef, int row) {
  Coef k;
  k.a = __ldg(coef + row * 8 + 0); k.b = __ldg(coef + row * 8 + 1); k.c1 = __ldg(coef + row * 8 + 2);
  k.c2 = __ldg(coef + row * 8 + 3); k.s = __ldg(coef + row * 8 + 4);
  k.inv_a = 1.0 / (double)k.a;
  return k;
}

// four N(0,1) values for the elements 4*vec .. 4*vec+3 of step row `row`
__device__ __forceinline__ float4 philox_normal4(uint64_t vec, uint32_t row, uint64_t seed) {
  uint32_t r[4];
  philox4x32_10(vec, row, seed, r);
  float4 z;
  box_muller(r[0], r[1], z.x, z.y);
  box_muller(r[2], r[3], z.z, z.w);
  return z;
}

__device__ __forceinline__ float step_one(float x, float e, float z, const Coef& k, int kind, int clip, float lo, float hi,
                                          bool use_noise) {
  // x0 = (x_t - sqrt(1-abar) * eps) / sqrt(abar)               diffusion.py:124
  // The IEEE fp32 quotient through fp64: n * (1/a) in double is within 2^-52 of n/a, far inside the 2^-49 gap that
  // separates an fp32 quotient from a rounding midpoint, so the final rounding is the correctly rounded n/a --
  // bit-identical to the reference's division at a third of div.rn.f32's instruction count.
  float x0 = __double2float_rn(__dmul_rn((double)__fsub_rn(x, __fmul_rn(k.b, e)), k.inv_a));
  if (clip) x0 = fminf(fmaxf(x0, lo), hi);                      // torch.clamp, diffusion.py:169 / :219
  // DDPM: c1*x0 + c2*x_t (diffusion.py:148); DDIM: sqrt(abar')*x0 + sqrt(1-abar'-s^2)*eps (:225-228)
  const float second = (kind == 0) ? __fmul_rn(k.c2, x) : __fmul_rn(k.c2, e);
  float out = __fadd_rn(__fmul_rn(k.c1, x0), second);
  if (use_noise) out = __fadd_rn(out, __fmul_rn(k.s, z));       // diffusion.py:181 / :232
  return out;
}

}  // namespace b2d
