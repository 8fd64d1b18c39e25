// fp32 GEMM for the fully-connected stacks (K6 in SURVEY.md 2.5): nn.Linear forward, input-gradient and
// weight-gradient of every MLP on the hot path (s2:24,43-48,77-89; mc3:60-69; cad:167-179,240-246,318-326,
// 361-367,407-413,435-461,525-538).  One kernel, three operand layouts:
//   C[i][j] (+)= sum_k A(i,k) * B(k,j)
//   A(i,k) = A[i*lda + k] (a_kmajor) or A[k*lda + i];   B(k,j) = B[j*ldb + k] (b_kmajor) or B[k*ldb + j]
// Epilogue (splits == 1): + bias[j] -> activation -> * mask[i][j]*mask_scale (dropout keep-mask) ; optional
// accumulate into C.  With splits > 1 the K range is split over gridDim.z and partial sums are added atomically
// (C must be pre-zeroed; bias/act/mask are then applied by cvad_bias_act_mask_f32).
#include "common.cuh"
#include "cvad_b200.h"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4, NT = 256;

template <bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__(NT) sgemm_kernel(int M, int N, int K, const float* __restrict__ A, long long lda,
                                                   const float* __restrict__ B, long long ldb, float* __restrict__ C, long long ldc,
                                                   const float* __restrict__ bias, int act, const float* __restrict__ mask,
                                                   float mask_scale, int accumulate, int k_per_split, int atomic_out,
                                                   const float* __restrict__ gate) {
  cvad_pdl_enter();
  if (gate && !(*gate > 0.f)) return;     // device-side "this branch received no gradient" switch
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int i0 = blockIdx.x * BM, j0 = blockIdx.y * BN;
  const int k_begin = blockIdx.z * k_per_split;
  const int k_end = min(K, k_begin + k_per_split);
  const int tm = tid & 15, tn = tid >> 4;

  float acc[TM][TN];
#pragma unroll
  for (int a = 0; a < TM; ++a)
#pragma unroll
    for (int b = 0; b < TN; ++b) acc[a][b] = 0.f;

  float ra[4], rb[4];
  auto load = [&](int k0) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int idx = tid + e * NT;   // 0..1023 over a 64x16 tile
      int ii, kk;
      if (A_KMAJOR) { kk = idx & 15; ii = idx >> 4; } else { ii = idx & 63; kk = idx >> 6; }
      int i = i0 + ii, k = k0 + kk;
      ra[e] = (i < M && k < k_end) ? __ldg(A_KMAJOR ? A + (long long)i * lda + k : A + (long long)k * lda + i) : 0.f;
      int jj;
      if (B_KMAJOR) { kk = idx & 15; jj = idx >> 4; } else { jj = idx & 63; kk = idx >> 6; }
      int j = j0 + jj;
      k = k0 + kk;
      rb[e] = (j < N && k < k_end) ? __ldg(B_KMAJOR ? B + (long long)j * ldb + k : B + (long long)k * ldb + j) : 0.f;
    }
  };
  auto store = [&]() {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int idx = tid + e * NT;
      int ii, kk, jj;
      if (A_KMAJOR) { kk = idx & 15; ii = idx >> 4; } else { ii = idx & 63; kk = idx >> 6; }
      As[kk][ii] = ra[e];
      if (B_KMAJOR) { kk = idx & 15; jj = idx >> 4; } else { jj = idx & 63; kk = idx >> 6; }
      Bs[kk][jj] = rb[e];
    }
  };

  if (k_begin < k_end) {
    load(k_begin);
    store();
    __syncthreads();
    for (int k0 = k_begin; k0 < k_end; k0 += BK) {
      const bool more = k0 + BK < k_end;
      if (more) load(k0 + BK);
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        const float4 av = *reinterpret_cast<const float4*>(&As[kk][tm * TM]);
        const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tn * TN]);
        const float a4[4] = {av.x, av.y, av.z, av.w};
        const float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int a = 0; a < TM; ++a)
#pragma unroll
          for (int b = 0; b < TN; ++b) acc[a][b] = fmaf(a4[a], b4[b], acc[a][b]);
      }
      __syncthreads();
      if (more) {
        store();
        __syncthreads();
      }
    }
  }

#pragma unroll
  for (int a = 0; a < TM; ++a) {
    int i = i0 + tm * TM + a;
    if (i >= M) continue;
#pragma unroll
    for (int b = 0; b < TN; ++b) {
      int j = j0 + tn * TN + b;
      if (j >= N) continue;
      float* c = C + (long long)i * ldc + j;
      if (atomic_out) {
        atomicAdd(c, acc[a][b]);
      } else {
        float v = acc[a][b] + (bias ? __ldg(bias + j) : 0.f);
        v = cvad_act(v, act);
        if (mask) v *= __ldg(mask + (long long)i * ldc + j) * mask_scale;
        if (accumulate) v += *c;
        *c = v;
      }
    }
  }
}

__global__ void bias_act_mask_kernel(float* __restrict__ y, long long rows, int cols, const float* __restrict__ bias, int act,
                                     const float* __restrict__ mask, float mask_scale) {
  cvad_pdl_enter();
  long long n = rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int j = (int)(i % cols);
    float v = y[i] + (bias ? __ldg(bias + j) : 0.f);
    v = cvad_act(v, act);
    if (mask) v *= __ldg(mask + i) * mask_scale;
    y[i] = v;
  }
}

// dz = dy * act'(y) * mask*scale ; also (optionally) column sums db[j] = sum_i dz[i][j]
__global__ void act_mask_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ mask,
                                    float mask_scale, int act, float* __restrict__ dz, long long n) {
  cvad_pdl_enter();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float g = dy[i];
    float yy = y ? y[i] : 0.f;
    if (mask) {
      float mk = __ldg(mask + i) * mask_scale;
      g *= mk;
      // y was stored AFTER the mask; recover the pre-mask activation output for sigmoid/tanh derivatives
      if (act == ACT_SIGMOID || act == ACT_TANH) yy = mk != 0.f ? yy / mk : 0.f;
    }
    dz[i] = g * cvad_act_grad_from_out(yy, act);
  }
}

// column sums over rows: out[j] (+)= sum_i x[i*ld + j]   (bias gradients).  One block = 32 columns x 8 row lanes; the 32
// threads of a warp read 32 consecutive columns of one row (coalesced), row lanes are reduced through shared memory.
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, long long rows, int cols, long long ld, float* __restrict__ out,
                                                     int accumulate) {
  cvad_pdl_enter();
  __shared__ float sh[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + cx;
  const long long chunk = (rows + gridDim.y - 1) / gridDim.y;
  const long long r0 = blockIdx.y * chunk, r1 = r0 + chunk < rows ? r0 + chunk : rows;
  float s = 0.f;
  if (j < cols)
    for (long long i = r0 + ry; i < r1; i += 8) s += x[i * ld + j];
  sh[ry][cx] = s;
  __syncthreads();
  if (ry == 0 && j < cols) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sh[k][cx];
    if (gridDim.y > 1) atomicAdd(out + j, t);
    else out[j] = accumulate ? out[j] + t : t;
  }
}

}  // namespace

CVAD_API int cvad_sgemm_f32(int M, int N, int K, const float* A, long long lda, int a_kmajor, const float* B, long long ldb,
                            int b_kmajor, float* C, long long ldc, const float* bias, int act, const float* mask, float mask_scale,
                            int accumulate, int splits, const float* gate, void* stream) {
  if (M <= 0 || N <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (splits < 1) splits = 1;
  int kps = ((K + splits - 1) / splits + BK - 1) / BK * BK;
  if (kps < BK) kps = BK;
  splits = (K + kps - 1) / kps;
  if (splits < 1) splits = 1;
  int atomic_out = splits > 1;
  dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN, splits);
  cudaError_t le;
#define GO(AK, BKM) le = cvad_launch_pdl(sgemm_kernel<AK, BKM>, grid, dim3(NT), 0, st, M, N, K, A, lda, B, ldb, C, ldc, bias, act, mask, mask_scale, \
                                         accumulate, kps, atomic_out, gate)
  if (a_kmajor && b_kmajor) GO(true, true);
  else if (a_kmajor) GO(true, false);
  else if (b_kmajor) GO(false, true);
  else GO(false, false);
#undef GO
  if (le != cudaSuccess) return (int)le;
  CVAD_LAUNCH_CHECK();
  return 0;
}

CVAD_API int cvad_bias_act_mask_f32(float* y, long long rows, int cols, const float* bias, int act, const float* mask,
                                    float mask_scale, void* stream) {
  long long n = rows * cols;
  if (n <= 0) return 0;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 4 * cvad_num_sms()) blocks = 4 * cvad_num_sms();
  cudaError_t le = cvad_launch_pdl(bias_act_mask_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, y, rows, cols, bias, act, mask, mask_scale);
  if (le != cudaSuccess) return (int)le;
  CVAD_LAUNCH_CHECK();
  return 0;
}

CVAD_API int cvad_act_mask_bwd_f32(const float* dy, const float* y, const float* mask, float mask_scale, int act, float* dz,
                                   long long n, void* stream) {
  if (n <= 0) return 0;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 8 * cvad_num_sms()) blocks = 8 * cvad_num_sms();
  cudaError_t le = cvad_launch_pdl(act_mask_bwd_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, dy, y, mask, mask_scale, act, dz, n);
  if (le != cudaSuccess) return (int)le;
  CVAD_LAUNCH_CHECK();
  return 0;
}

CVAD_API int cvad_colsum_f32(const float* x, long long rows, int cols, long long ld, float* out, int accumulate, void* stream) {
  if (cols <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int bx = (cols + 31) / 32;
  int by = 1;
  if (accumulate) {            // split long reductions over rows; partial sums are added atomically onto the running gradient
    by = (int)((rows + 255) / 256);
    const int cap = (2 * cvad_num_sms() + bx - 1) / bx;
    if (by > cap) by = cap;
    if (by < 1) by = 1;
  }
  cudaError_t le = cvad_launch_pdl(colsum_kernel, dim3(bx, by), dim3(256), 0, st, x, rows, cols, ld, out, accumulate);
  if (le != cudaSuccess) return (int)le;
  CVAD_LAUNCH_CHECK();
  return 0;
}
