// Declarations shared by the two conv-accumulate implementations (ddk_conv.cu: v1 reference kernel,
// ddk_conv2.cu: persistent warp-specialised kernel).
#pragma once

#include "ddk_device.cuh"

namespace ddk {

struct AccArgs {
  int NL;
  const int* seg_order; const int* seg_sidx;
  const int* seg_base; const int* seg_cnt; const int2* seg_list;
  const float* x;                 // [N][84] layer input
  const float* proj;              // [N][4][72]
  const float* ea_pool; const float4* sh_pool;
  const float* W1[4];             // [72][72] (only the edge-embedding columns 0:24 are read here)
  float* A; float* Bsum;          // scratch [nseg][U][72], [nseg][U]
};

constexpr int KC = 32;            // edges per chunk
constexpr int ACC_THREADS = 288;  // 9 warps: warp w owns columns j in [8w, 8w+8), lane owns rows u = lane + 32 a

// basis function u of level LV (kernel order [0e | 1o comp-major | 1e comp-major | 0o], constant factors folded
// into the packed weights): returns type (0: x[i0]*sh[m], 1: dot(x[i0..i0+2], s), 2: cross(x[i0..], s)[m-1]).
template <int LV>
__device__ __forceinline__ void basis_desc(int u, int& type, int& i0, int& m) {
  constexpr int F0e = LV >= 1 ? 30 : 24;
  constexpr int F1o = LV >= 2 ? 36 : (LV == 1 ? 30 : 24);
  constexpr int F1e = LV >= 3 ? 36 : (LV == 2 ? 12 : (LV == 1 ? 6 : 0));
  constexpr int X1O = 24, X1E = 42, X0O = 60;
  type = 0; i0 = 0; m = 0;
  if (u < F0e) {
    if (u < 24) { type = 0; i0 = u; m = 0; } else { type = 1; i0 = X1O + 3 * (u - 24); }
    return;
  }
  u -= F0e;
  if (u < 3 * F1o) {
    int c = u / F1o, k = u % F1o;
    if (k < 24) { type = 0; i0 = k; m = 1 + c; }
    else if (k < 30) { type = 0; i0 = X1O + 3 * (k - 24) + c; m = 0; }
    else { type = 2; i0 = X1E + 3 * (k - 30); m = 1 + c; }
    return;
  }
  u -= 3 * F1o;
  constexpr int F1eD = F1e > 0 ? F1e : 1;
  if (u < 3 * F1e) {
    int c = u / F1eD, k = u % F1eD;
    if (k < 6) { type = 2; i0 = X1O + 3 * k; m = 1 + c; }
    else if (k < 12) { type = 0; i0 = X1E + 3 * (k - 6) + c; m = 0; }
    else { type = 0; i0 = X0O + (k - 12); m = 1 + c; }
    return;
  }
  u -= 3 * F1e;
  if (u < 6) { type = 1; i0 = X1E + 3 * u; } else { type = 0; i0 = X0O + (u - 6); m = 0; }
}

template <int LV>
struct AccCfg {
  static constexpr int U = LV == 0 ? 96 : (LV == 1 ? 138 : (LV == 2 ? 180 : 276));
  static constexpr int NA = (U + 31) / 32;
  static constexpr int BS = NA * 32;
  static constexpr int DINP = LV == 0 ? 24 : (LV == 1 ? 44 : (LV == 2 ? 60 : 84));   // input width rounded up to 4
};


// ---- contraction
struct ConArgs {
  int NL;
  int lig0, lig1, rec0, rec1;     // node ranges of this chunk (rec indices are within the receptor type)
  const int* seg_sidx; const int* seg_cnt;
  const float* A; const float* Bsum;
  const float* W2p[4]; const float* b2p[4];
  const float* bn_scale; const float* bn_shift;
  const float* x_in; float* x_out;
  LayerInfo li;
};


void launch_conv_contract2(DdkCtx* c, const ConArgs& q, cudaStream_t st);
void launch_conv_accum2(DdkCtx* c, const LayerInfo& li, const Chunk& ch, const AccArgs& a, cudaStream_t st);
cudaError_t conv2_configure();
cudaError_t contract2_configure();

}  // namespace ddk
