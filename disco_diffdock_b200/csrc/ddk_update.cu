// Reverse-step update: perturbation arithmetic + modify_conformer_batch.
// Reference: /root/reference/utils/sampling.py:137-198 (a*score + b*z per component),
//            /root/reference/utils/diffusion_utils.py:37-55 (rigid move, torsions, Kabsch re-alignment),
//            /root/reference/utils/torsion.py:71-86 (sequential torsion rotations),
//            /root/reference/utils/geometry.py:38-85 (axis-angle -> quaternion -> matrix), :126-156 (Kabsch).
// One CTA per graph.  The 3x3 SVD of the reference's Kabsch step is replaced by Horn's closed-form quaternion
// solution (largest eigenvector of a symmetric 4x4, Jacobi in fp64), which yields the same proper rotation
// including the reflection-corrected case (checked against numpy SVD in tests/test_host_logic.py).
#include "ddk_device.cuh"

namespace ddk {

struct UpdArgs {
  const int* lig_ptr; const int* rot_ptr; const int* rot_u; const int* rot_v;
  const int64_t* mr_off; const uint8_t* mask_rotate;
  float* pos;
  const float* tr; const float* rot; const float* tor;
  const float* z_tr; const float* z_rot; const float* z_tor;
  DdkStepCoef cf;
  int has_tor;
};

// geometry.py:38-85: quaternion (cos(a/2), axis * sin(a/2)/a) with the small-angle series, then the matrix
__host__ __device__ void axis_angle_to_matrix(float x, float y, float z, float* R) {
  float ang = sqrtf(x * x + y * y + z * z);
  float half = 0.5f * ang;
  float k = (fabsf(ang) < 1e-6f) ? (0.5f - ang * ang / 48.f) : (sinf(half) / ang);
  float r = cosf(half), i = x * k, j = y * k, kk = z * k;
  float two_s = 2.0f / (r * r + i * i + j * j + kk * kk);
  R[0] = 1.f - two_s * (j * j + kk * kk); R[1] = two_s * (i * j - kk * r);        R[2] = two_s * (i * kk + j * r);
  R[3] = two_s * (i * j + kk * r);        R[4] = 1.f - two_s * (i * i + kk * kk); R[5] = two_s * (j * kk - i * r);
  R[6] = two_s * (i * kk - j * r);        R[7] = two_s * (j * kk + i * r);        R[8] = 1.f - two_s * (i * i + j * j);
}

// largest eigenvector of a symmetric 4x4 (cyclic Jacobi, fp64)
__host__ __device__ void sym4_top_eigvec(double A[4][4], double q[4]) {
  double V[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0;
    for (int a = 0; a < 4; ++a)
      for (int b = a + 1; b < 4; ++b) off += A[a][b] * A[a][b];
    if (off < 1e-300) break;
    for (int a = 0; a < 4; ++a)
      for (int b = a + 1; b < 4; ++b) {
        if (fabs(A[a][b]) < 1e-300) continue;
        double theta = (A[b][b] - A[a][a]) / (2.0 * A[a][b]);
        double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 4; ++k) {
          double aka = A[k][a], akb = A[k][b];
          A[k][a] = c * aka - s * akb;
          A[k][b] = s * aka + c * akb;
        }
        for (int k = 0; k < 4; ++k) {
          double aak = A[a][k], abk = A[b][k];
          A[a][k] = c * aak - s * abk;
          A[b][k] = s * aak + c * abk;
        }
        for (int k = 0; k < 4; ++k) {
          double vka = V[k][a], vkb = V[k][b];
          V[k][a] = c * vka - s * vkb;
          V[k][b] = s * vka + c * vkb;
        }
      }
  }
  int best = 0;
  for (int a = 1; a < 4; ++a)
    if (A[a][a] > A[best][best]) best = a;
  for (int k = 0; k < 4; ++k) q[k] = V[k][best];
}

// R (row-major 3x3), t with R a_n + t ~ b_n in the least-squares sense, proper rotation (Horn 1987)
__host__ __device__ void kabsch_horn(const float* A, const float* Bp, int N, float* R9, float* t3) {
  double ca[3] = {0, 0, 0}, cb[3] = {0, 0, 0};
  for (int n = 0; n < N; ++n)
    for (int d = 0; d < 3; ++d) { ca[d] += A[3 * n + d]; cb[d] += Bp[3 * n + d]; }
  for (int d = 0; d < 3; ++d) { ca[d] /= N; cb[d] /= N; }
  double S[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int n = 0; n < N; ++n)
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) S[i][j] += ((double)A[3 * n + i] - ca[i]) * ((double)Bp[3 * n + j] - cb[j]);
  double Nm[4][4] = {
      {S[0][0] + S[1][1] + S[2][2], S[1][2] - S[2][1], S[2][0] - S[0][2], S[0][1] - S[1][0]},
      {S[1][2] - S[2][1], S[0][0] - S[1][1] - S[2][2], S[0][1] + S[1][0], S[2][0] + S[0][2]},
      {S[2][0] - S[0][2], S[0][1] + S[1][0], -S[0][0] + S[1][1] - S[2][2], S[1][2] + S[2][1]},
      {S[0][1] - S[1][0], S[2][0] + S[0][2], S[1][2] + S[2][1], -S[0][0] - S[1][1] + S[2][2]}};
  double q[4];
  sym4_top_eigvec(Nm, q);
  double nq = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  double r = q[0] / nq, i = q[1] / nq, j = q[2] / nq, k = q[3] / nq;
  double R[9] = {1 - 2 * (j * j + k * k), 2 * (i * j - k * r), 2 * (i * k + j * r),
                 2 * (i * j + k * r), 1 - 2 * (i * i + k * k), 2 * (j * k - i * r),
                 2 * (i * k - j * r), 2 * (j * k + i * r), 1 - 2 * (i * i + j * j)};
  for (int d = 0; d < 9; ++d) R9[d] = (float)R[d];
  for (int d = 0; d < 3; ++d) t3[d] = (float)(-(R[3 * d] * ca[0] + R[3 * d + 1] * ca[1] + R[3 * d + 2] * ca[2]) + cb[d]);
}

constexpr int UPD_THREADS = 128;

__global__ void __launch_bounds__(UPD_THREADS) k_update(UpdArgs p) {
  extern __shared__ float usm[];
  const int g = blockIdx.x, tid = threadIdx.x;
  const int l0 = p.lig_ptr[g], l1 = p.lig_ptr[g + 1], N = l1 - l0;
  float* flex = usm;            // [N][3]
  float* rigid = usm + 3 * N;   // [N][3]
  __shared__ float sR[9], sC[3], sT[3], sPv[3];
  for (int i = tid; i < 3 * N; i += UPD_THREADS) flex[i] = p.pos[(size_t)l0 * 3 + i];
  __syncthreads();
  if (tid == 0) {
    float c[3] = {0.f, 0.f, 0.f};
    for (int n = 0; n < N; ++n) { c[0] += flex[3 * n]; c[1] += flex[3 * n + 1]; c[2] += flex[3 * n + 2]; }
    float tp[3], rp[3];
    for (int d = 0; d < 3; ++d) {
      sC[d] = c[d] / (float)N;
      float zt = p.z_tr ? p.z_tr[g * 3 + d] : 0.f, zr = p.z_rot ? p.z_rot[g * 3 + d] : 0.f;
      tp[d] = __fadd_rn(__fmul_rn(p.cf.a_tr, p.tr[g * 3 + d]), __fmul_rn(p.cf.b_tr, zt));
      rp[d] = __fadd_rn(__fmul_rn(p.cf.a_rot, p.rot[g * 3 + d]), __fmul_rn(p.cf.b_rot, zr));
      sT[d] = tp[d];
    }
    axis_angle_to_matrix(rp[0], rp[1], rp[2], sR);
  }
  __syncthreads();
  // rigid_new_pos = (pos - centre) R^T + tr + centre      (diffusion_utils.py:44-46)
  for (int n = tid; n < N; n += UPD_THREADS) {
    float x = flex[3 * n] - sC[0], y = flex[3 * n + 1] - sC[1], z = flex[3 * n + 2] - sC[2];
#pragma unroll
    for (int d = 0; d < 3; ++d)
      rigid[3 * n + d] = (sR[3 * d] * x + sR[3 * d + 1] * y + sR[3 * d + 2] * z) + sT[d] + sC[d];
  }
  __syncthreads();
  const int b0 = p.rot_ptr[g], b1 = p.rot_ptr[g + 1];
  if (!p.has_tor || b0 == b1) {
    for (int i = tid; i < 3 * N; i += UPD_THREADS) p.pos[(size_t)l0 * 3 + i] = rigid[i];
    return;
  }
  for (int i = tid; i < 3 * N; i += UPD_THREADS) flex[i] = rigid[i];
  __syncthreads();
  const uint8_t* mask = p.mask_rotate + p.mr_off[g];
  for (int b = b0; b < b1; ++b) {           // sequential over rotatable bonds (torsion.py:73-84)
    const int u = p.rot_u[b] - l0, v = p.rot_v[b] - l0;
    if (tid == 0) {
      float zt = p.z_tor ? p.z_tor[b] : 0.f;
      float upd = __fadd_rn(__fmul_rn(p.cf.a_tor, p.tor[b]), __fmul_rn(p.cf.b_tor, zt));
      float ax = flex[3 * u] - flex[3 * v], ay = flex[3 * u + 1] - flex[3 * v + 1], az = flex[3 * u + 2] - flex[3 * v + 2];
      float nrm = sqrtf(ax * ax + ay * ay + az * az);
      axis_angle_to_matrix(ax / nrm * upd, ay / nrm * upd, az / nrm * upd, sR);
      sPv[0] = flex[3 * v]; sPv[1] = flex[3 * v + 1]; sPv[2] = flex[3 * v + 2];
    }
    __syncthreads();
    const uint8_t* mrow = mask + (size_t)(b - b0) * N;
    for (int n = tid; n < N; n += UPD_THREADS) {
      if (mrow[n] && n != v) {
        float x = flex[3 * n] - sPv[0], y = flex[3 * n + 1] - sPv[1], z = flex[3 * n + 2] - sPv[2];
#pragma unroll
        for (int d = 0; d < 3; ++d) flex[3 * n + d] = (sR[3 * d] * x + sR[3 * d + 1] * y + sR[3 * d + 2] * z) + sPv[d];
      }
    }
    __syncthreads();
  }
  // Kabsch: R, t minimising |R flex + t - rigid|  (geometry.py:126-156)
  if (tid == 0) kabsch_horn(flex, rigid, N, sR, sT);
  __syncthreads();
  for (int n = tid; n < N; n += UPD_THREADS) {
    float x = flex[3 * n], y = flex[3 * n + 1], z = flex[3 * n + 2];
#pragma unroll
    for (int d = 0; d < 3; ++d)
      p.pos[(size_t)(l0 + n) * 3 + d] = (sR[3 * d] * x + sR[3 * d + 1] * y + sR[3 * d + 2] * z) + sT[d];
  }
}

void launch_update(DdkCtx* c, float* lig_pos, const float* tr, const float* rot, const float* tor, const float* z_tr,
                   const float* z_rot, const float* z_tor, DdkStepCoef coef, cudaStream_t st) {
  UpdArgs p;
  p.lig_ptr = ptr<int>(c->b_lig_ptr); p.rot_ptr = ptr<int>(c->b_rot_ptr);
  p.rot_u = ptr<int>(c->b_rot_u); p.rot_v = ptr<int>(c->b_rot_v);
  p.mr_off = ptr<int64_t>(c->b_mr_off); p.mask_rotate = c->mask_rotate;
  p.pos = lig_pos; p.tr = tr; p.rot = rot; p.tor = tor;
  p.z_tr = z_tr; p.z_rot = z_rot; p.z_tor = z_tor;
  p.cf = coef;
  p.has_tor = (!c->cfg.no_torsion && c->RB > 0 && tor != nullptr) ? 1 : 0;
  size_t smem = (size_t)6 * c->maxNl * sizeof(float);
  LaunchScope ls(c, PC_UPDATE, st);
  k_update<<<c->B, UPD_THREADS, smem, st>>>(p);
}

void host_kabsch(const float* A, const float* Bp, int N, float* R9, float* t3) { kabsch_horn(A, Bp, N, R9, t3); }
void host_axis_angle(const float* aa, float* R9) { axis_angle_to_matrix(aa[0], aa[1], aa[2], R9); }

}  // namespace ddk
