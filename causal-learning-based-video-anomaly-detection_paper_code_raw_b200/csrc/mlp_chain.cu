// Fused MLP chains for the dense tail of M-A (and the heads of M-B / M-C / M-E): a whole stack of small nn.Linear layers -- bias,
// activation and dropout mask included -- in ONE launch forward and ONE launch for the data-gradient chain backward
// (cad:167-179 detector, 240-246 re-id, 318-326 factor encoder, 361-367 edge predictor, 407-413 dynamics, 435-461 scorers, 525-538 direct
// classifier; s2:43-48, 77-89; mc3:60-69).
//
// Why: these layers are 4..512 wide on 32..2560 rows -- a few MFLOP each -- so a step spent ~120 dependent launches of ~5 us on them,
// every one a round trip through HBM for a tensor of a few KB.  Here a CTA owns MC_ROWS rows and walks the layers with the activations
// in shared memory; weights stream from L2 (the largest chain holds 0.7 MB of them).  The weight / bias gradients stay what they were
// (split-K GEMMs and column sums on the side stream, off the critical path): the backward kernel leaves every layer's dz in HBM for them.
//
// Thread mapping.  Forward, din >= 64: a warp owns output columns j = warp, warp + 8, ...; lanes stride the reduction index (coalesced
// 128-byte weight reads, conflict-free activation reads), eight row accumulators per lane, and a halving butterfly (9 shuffles instead
// of 40) leaves the eight row sums on eight lanes.  din < 64: one (row, column) dot product per thread.  Backward data-gradient
// dh[r][k] = sum_j dz[r][j] W[j][k]: a thread owns column k for all rows, so W is read coalesced along k and dz is a shared-memory
// broadcast; for narrow layers the j range is split over thread groups and combined through shared memory.
#include "common.cuh"
#include "cvad_b200.h"

namespace {

constexpr int MC_MAX_LAYERS = 8;
constexpr int MC_ROWS = 8;
constexpr int MC_THREADS = 256;
constexpr int MC_MAXDIM = 512;

struct McLayer {
  const float* W;         // (dout, din) row-major, nn.Linear layout
  const float* b;         // (dout) or NULL
  const float* mask;      // (rows, dout) keep-mask or NULL
  float* save;            // forward: (rows, dout) stored output (post activation, post mask) or NULL; backward: the same, read-only
  float* dz;              // backward: (rows, dout) gradient w.r.t. the pre-activation, written for the weight-gradient GEMMs
  float mask_scale;
  int din, dout, act;
};

struct McChain {
  int n;
  long long rows;
  McLayer L[MC_MAX_LAYERS];
};

// eight per-lane partial sums -> the eight full sums, one per 4-lane group (lane = 4*row + x holds row's total)
__device__ __forceinline__ float reduce8(const float (&a)[8], int lane) {
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
  float y[4], z[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) y[i] = (b4 ? a[i + 4] : a[i]) + __shfl_xor_sync(0xffffffffu, b4 ? a[i] : a[i + 4], 16);
#pragma unroll
  for (int i = 0; i < 2; ++i) z[i] = (b3 ? y[i + 2] : y[i]) + __shfl_xor_sync(0xffffffffu, b3 ? y[i] : y[i + 2], 8);
  float v = (b2 ? z[1] : z[0]) + __shfl_xor_sync(0xffffffffu, b2 ? z[0] : z[1], 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;                                   // row index = 4*b4 + 2*b3 + b2
}

__global__ void __launch_bounds__(MC_THREADS) mlp_chain_fwd_kernel(const McChain ch, const float* __restrict__ x, float* __restrict__ out) {
  cvad_pdl_enter();
  __shared__ float hbuf[2][MC_ROWS][MC_MAXDIM + 4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long r0 = (long long)blockIdx.x * MC_ROWS;
  const int nr = (int)(ch.rows - r0 < MC_ROWS ? ch.rows - r0 : MC_ROWS);
  const int d0 = ch.L[0].din;
  for (int i = tid; i < MC_ROWS * d0; i += MC_THREADS) {
    const int r = i / d0, k = i - r * d0;
    hbuf[0][r][k] = r < nr ? __ldg(x + (r0 + r) * d0 + k) : 0.f;
  }
  __syncthreads();
  int cur = 0;
  for (int l = 0; l < ch.n; ++l) {
    const McLayer& L = ch.L[l];
    const int din = L.din, dout = L.dout;
    float(*h)[MC_MAXDIM + 4] = hbuf[cur];
    float(*hn)[MC_MAXDIM + 4] = hbuf[cur ^ 1];
    const bool last = l == ch.n - 1;
    auto finish = [&](int r, int j, float v) {
      v += L.b ? __ldg(L.b + j) : 0.f;
      v = cvad_act(v, L.act);
      if (r < nr) {
        if (L.mask) v *= __ldg(L.mask + (r0 + r) * dout + j) * L.mask_scale;
        if (L.save) L.save[(r0 + r) * dout + j] = v;
        if (last) out[(r0 + r) * dout + j] = v;
      }
      hn[r][j] = v;
    };
    if (din >= 64) {
      for (int j = warp; j < dout; j += MC_THREADS / 32) {
        float acc[MC_ROWS];
#pragma unroll
        for (int r = 0; r < MC_ROWS; ++r) acc[r] = 0.f;
        const float* w = L.W + (long long)j * din;
        for (int k = lane; k < din; k += 32) {
          const float wv = __ldg(w + k);
#pragma unroll
          for (int r = 0; r < MC_ROWS; ++r) acc[r] = fmaf(h[r][k], wv, acc[r]);
        }
        const float v = reduce8(acc, lane);
        if ((lane & 3) == 0) finish(lane >> 2, j, v);
      }
    } else {
      for (int i = tid; i < MC_ROWS * dout; i += MC_THREADS) {
        const int r = i / dout, j = i - r * dout;
        const float* w = L.W + (long long)j * din;
        float acc = 0.f;
        for (int k = 0; k < din; ++k) acc = fmaf(h[r][k], __ldg(w + k), acc);
        finish(r, j, acc);
      }
    }
    __syncthreads();
    cur ^= 1;
  }
}

__global__ void __launch_bounds__(MC_THREADS) mlp_chain_bwd_kernel(const McChain ch, const float* __restrict__ dy, float* __restrict__ dx) {
  cvad_pdl_enter();
  __shared__ float gbuf[2][MC_ROWS][MC_MAXDIM + 4];
  __shared__ float part[MC_THREADS][MC_ROWS + 1];            // partial column sums of narrow layers (j range split over thread groups)
  const int tid = threadIdx.x;
  const long long r0 = (long long)blockIdx.x * MC_ROWS;
  const int nr = (int)(ch.rows - r0 < MC_ROWS ? ch.rows - r0 : MC_ROWS);
  const int dl = ch.L[ch.n - 1].dout;
  for (int i = tid; i < MC_ROWS * dl; i += MC_THREADS) {
    const int r = i / dl, j = i - r * dl;
    gbuf[0][r][j] = r < nr ? __ldg(dy + (r0 + r) * dl + j) : 0.f;
  }
  __syncthreads();
  int cur = 0;
  for (int l = ch.n - 1; l >= 0; --l) {
    const McLayer& L = ch.L[l];
    const int din = L.din, dout = L.dout;
    float(*g)[MC_MAXDIM + 4] = gbuf[cur];
    float(*gp)[MC_MAXDIM + 4] = gbuf[cur ^ 1];
    // dz = g * mask*scale * act'(y)   (y = the stored post-mask output; same convention as cvad_act_mask_bwd_f32)
    for (int i = tid; i < MC_ROWS * dout; i += MC_THREADS) {
      const int r = i / dout, j = i - r * dout;
      float v = 0.f;
      if (r < nr) {
        v = g[r][j];
        if (L.mask) v *= __ldg(L.mask + (r0 + r) * dout + j) * L.mask_scale;
        if (L.act != ACT_NONE) v *= cvad_act_grad_from_out(__ldg(L.save + (r0 + r) * dout + j), L.act);
        L.dz[(r0 + r) * dout + j] = v;
      }
      g[r][j] = v;
    }
    __syncthreads();
    if (l == 0 && dx == nullptr) break;
    // gp[r][k] = sum_j dz[r][j] W[j][k]
    if (din >= MC_THREADS / 2) {
      for (int k = tid; k < din; k += MC_THREADS) {
        float acc[MC_ROWS];
#pragma unroll
        for (int r = 0; r < MC_ROWS; ++r) acc[r] = 0.f;
        for (int j = 0; j < dout; ++j) {
          const float wv = __ldg(L.W + (long long)j * din + k);
#pragma unroll
          for (int r = 0; r < MC_ROWS; ++r) acc[r] = fmaf(g[r][j], wv, acc[r]);
        }
#pragma unroll
        for (int r = 0; r < MC_ROWS; ++r) gp[r][k] = acc[r];
      }
    } else {
      // narrow input: ng thread groups share a column, each walks a slice of j; partial sums meet in shared memory
      int ng = MC_THREADS / din;                       // >= 2
      if (ng > dout) ng = dout;
      const int grp = tid / din, k = tid - grp * din;
      float acc[MC_ROWS];
#pragma unroll
      for (int r = 0; r < MC_ROWS; ++r) acc[r] = 0.f;
      if (grp < ng) {
        for (int j = grp; j < dout; j += ng) {
          const float wv = __ldg(L.W + (long long)j * din + k);
#pragma unroll
          for (int r = 0; r < MC_ROWS; ++r) acc[r] = fmaf(g[r][j], wv, acc[r]);
        }
#pragma unroll
        for (int r = 0; r < MC_ROWS; ++r) part[tid][r] = acc[r];
      }
      __syncthreads();
      for (int i = tid; i < MC_ROWS * din; i += MC_THREADS) {
        const int r = i / din, kk = i - r * din;
        float s = 0.f;
        for (int q = 0; q < ng; ++q) s += part[q * din + kk][r];
        gp[r][kk] = s;
      }
    }
    __syncthreads();
    cur ^= 1;
  }
  if (dx) {
    const int d0 = ch.L[0].din;
    for (int i = tid; i < nr * d0; i += MC_THREADS) {
      const int r = i / d0, k = i - r * d0;
      dx[(r0 + r) * d0 + k] = gbuf[cur][r][k];
    }
  }
}

int fill_chain(McChain& ch, long long rows, int n_layers, const int* dims, const int* acts, const void* const* weights,
               const void* const* biases, const void* const* masks, const float* mask_scales, void* const* saves, void* const* dzs) {
  if (n_layers < 1 || n_layers > MC_MAX_LAYERS || rows < 0) return (int)cudaErrorInvalidValue;
  ch.n = n_layers;
  ch.rows = rows;
  for (int l = 0; l < n_layers; ++l) {
    if (dims[l] < 1 || dims[l] > MC_MAXDIM || dims[l + 1] < 1 || dims[l + 1] > MC_MAXDIM) return (int)cudaErrorInvalidValue;
    McLayer& L = ch.L[l];
    L.W = (const float*)weights[l];
    L.b = biases ? (const float*)biases[l] : nullptr;
    L.mask = masks ? (const float*)masks[l] : nullptr;
    L.save = saves ? (float*)saves[l] : nullptr;
    L.dz = dzs ? (float*)dzs[l] : nullptr;
    L.mask_scale = mask_scales ? mask_scales[l] : 1.f;
    L.din = dims[l];
    L.dout = dims[l + 1];
    L.act = acts[l];
  }
  return 0;
}

}  // namespace

CVAD_API int cvad_mlp_chain_fwd_f32(const float* x, long long rows, int n_layers, const int* dims, const int* acts, const void* const* weights,
                                    const void* const* biases, const void* const* masks, const float* mask_scales, void* const* saves, float* out,
                                    void* stream) {
  if (rows == 0) return 0;
  McChain ch;
  int e = fill_chain(ch, rows, n_layers, dims, acts, weights, biases, masks, mask_scales, saves, nullptr);
  if (e) return e;
  const unsigned grid = (unsigned)((rows + MC_ROWS - 1) / MC_ROWS);
  cudaError_t ce = cvad_launch_pdl(mlp_chain_fwd_kernel, dim3(grid), dim3(MC_THREADS), 0, (cudaStream_t)stream, ch, x, out);
  return (int)ce;
}

CVAD_API int cvad_mlp_chain_bwd_f32(const float* dy, long long rows, int n_layers, const int* dims, const int* acts, const void* const* weights,
                                    const void* const* masks, const float* mask_scales, const void* const* saves, void* const* dzs, float* dx,
                                    void* stream) {
  if (rows == 0) return 0;
  McChain ch;
  int e = fill_chain(ch, rows, n_layers, dims, acts, weights, nullptr, masks, mask_scales, (void* const*)saves, dzs);
  if (e) return e;
  for (int l = 0; l < n_layers; ++l)
    if (!ch.L[l].dz || (ch.L[l].act != ACT_NONE && !ch.L[l].save)) return (int)cudaErrorInvalidValue;
  const unsigned grid = (unsigned)((rows + MC_ROWS - 1) / MC_ROWS);
  cudaError_t ce = cvad_launch_pdl(mlp_chain_bwd_kernel, dim3(grid), dim3(MC_THREADS), 0, (cudaStream_t)stream, ch, dy, dx);
  return (int)ce;
}
