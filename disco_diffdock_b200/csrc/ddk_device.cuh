// Device helpers shared by the kernels.
#pragma once

#include "ddk_internal.h"

namespace ddk {

// |b - a|^2 evaluated as ((dx*dx + dy*dy) + dz*dz) with every operation rounded to fp32 (no FMA contraction):
// bit-identical to the oracle's torch-CPU evaluation, so radius cut-offs select the same edge set
// (torch_cluster.radius semantics, strict '<'; /root/reference/models/score_model.py:315, 379-384, 430).
__device__ __forceinline__ float dist2_unfused(float ax, float ay, float az, float bx, float by, float bz) {
  float dx = __fsub_rn(bx, ax), dy = __fsub_rn(by, ay), dz = __fsub_rn(bz, az);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// e3nn spherical_harmonics(lmax=1, normalize=True, normalization='component'): (1, sqrt3 * v/|v|)
__device__ __forceinline__ float4 sh_l01(float vx, float vy, float vz, float* norm_out) {
  float n = sqrtf(vx * vx + vy * vy + vz * vz);
  *norm_out = n;
  float inv = 1.7320508075688772f / fmaxf(n, 1e-12f);
  return make_float4(1.f, vx * inv, vy * inv, vz * inv);
}

// GaussianSmearing (/root/reference/models/tensor_layers.py:171-181): exp(coeff * (d - mu_k)^2), k < 32.
// sm = [32 offsets | coeff]
__device__ __forceinline__ float smear1(const float* __restrict__ sm, float d, int k) {
  float t = d - sm[k];
  return expf(sm[32] * (t * t));
}

// (seg, n, base) of a work-list entry stored as int4 (the fourth word is padding).  Loaded as 8 + 4 bytes: with a 16-byte
// load ptxas treats the unused fourth destination register as free and later writes to it stall on the pending load.
__device__ __forceinline__ int4 load_seg_entry(const int4* __restrict__ e) {
  const int2 a = __ldg(reinterpret_cast<const int2*>(e));
  const int b = __ldg(reinterpret_cast<const int*>(e) + 2);
  return make_int4(a.x, a.y, b, 0);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace ddk
