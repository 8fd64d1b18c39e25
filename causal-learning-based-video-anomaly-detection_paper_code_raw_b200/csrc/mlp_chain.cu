// Fused MLP chains for the dense tail of M-A (and the heads of M-B / M-C / M-E): a whole stack of small nn.Linear layers -- bias,
// activation and dropout mask included -- in ONE launch forward and ONE launch for the data-gradient chain backward
// (cad:167-179 detector, 240-246 re-id, 318-326 factor encoder, 361-367 edge predictor, 407-413 dynamics, 435-461 scorers, 525-538 direct
// classifier; s2:43-48, 77-89; mc3:60-69).
//
// Why: these layers are 4..512 wide on 32..2560 rows -- a few MFLOP each -- so a step spent ~120 dependent launches of ~5 us on them,
// every one a round trip through HBM for a tensor of a few KB.  Here a CTA owns MC_ROWS rows and walks the layers with the activations
// in shared memory.  The weight / bias gradients stay what they were
// (split-K GEMMs and column sums on the side stream, off the critical path): the backward kernel leaves every layer's dz in HBM for them.
//
// Per layer a CTA first stages the whole weight matrix in shared memory (cooperative, coalesced, every thread with several loads in
// flight -- a first version that read weights from L2 inside the dot-product loops was latency-bound at 80 us per chain), then every thread
// computes (row, column) dot products from shared memory: forward h'[r][j] = sum_k h[r][k] W[j][k] (weight rows padded to an odd pitch:
// conflict-free across j), backward dh[r][k] = sum_j dz[r][j] W[j][k] (consecutive k: conflict-free).  A layer must fit: (din | 1) * dout <= 33024
// and widths <= 256 -- the 6144 -> 512 -> 256 layers of the detector / classifier stay split-K GEMMs, which spread their weights over many SMs
// instead of streaming 0.5 MB through each of 4..64 CTAs.
#include "common.cuh"
#include "cvad_b200.h"

namespace {

constexpr int MC_MAX_LAYERS = 8;
constexpr int MC_ROWS = 8;
constexpr int MC_THREADS = 256;
constexpr int MC_MAXDIM = 256;
constexpr int MC_MAXW = 33024;                         // (din | 1) * dout of one layer: up to 256 -> 128 (129 KB of shared memory)
constexpr int MC_HPITCH = MC_MAXDIM + 4;
constexpr size_t MC_SMEM = (size_t)(2 * MC_ROWS * MC_HPITCH + MC_MAXW + MC_MAXDIM) * sizeof(float);      // activations x2 + weights (padded pitch)

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

// stage W (dout, din) row-major into shared memory with row pitch wp (odd when din is even: rows land in different banks)
__device__ __forceinline__ void stage_weights(const float* __restrict__ W, int din, int dout, int wp, float* __restrict__ ws) {
  const int n = din * dout;
  if ((din & 3) == 0 && ((uintptr_t)W & 15) == 0) {
    const int n4 = n >> 2, d4 = din >> 2;
    for (int i = threadIdx.x; i < n4; i += MC_THREADS) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(W) + i);
      const int j = i / d4, k = (i - j * d4) << 2;
      float* d = ws + j * wp + k;
      d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
  } else {
    for (int i = threadIdx.x; i < n; i += MC_THREADS) {
      const int j = i / din, k = i - j * din;
      ws[j * wp + k] = __ldg(W + i);
    }
  }
}

__global__ void __launch_bounds__(MC_THREADS) mlp_chain_fwd_kernel(const McChain ch, const float* __restrict__ x, float* __restrict__ out) {
  cvad_pdl_enter();
  extern __shared__ float mc_smem[];
  float* hb[2] = {mc_smem, mc_smem + MC_ROWS * MC_HPITCH};
  float* ws = mc_smem + 2 * MC_ROWS * MC_HPITCH;
  const int tid = threadIdx.x;
  const long long r0 = (long long)blockIdx.x * MC_ROWS;
  const int nr = (int)(ch.rows - r0 < MC_ROWS ? ch.rows - r0 : MC_ROWS);
  const int d0 = ch.L[0].din;
  for (int i = tid; i < MC_ROWS * d0; i += MC_THREADS) {
    const int r = i / d0, k = i - r * d0;
    hb[0][r * MC_HPITCH + k] = r < nr ? __ldg(x + (r0 + r) * d0 + k) : 0.f;
  }
  int cur = 0;
  for (int l = 0; l < ch.n; ++l) {
    const McLayer L = ch.L[l];
    const int din = L.din, dout = L.dout;
    const int wp = din | 1;
    const float* h = hb[cur];
    float* hn = hb[cur ^ 1];
    const bool last = l == ch.n - 1;
    stage_weights(L.W, din, dout, wp, ws);
    __syncthreads();
    // (row, column) items, column fastest: a warp reads 32 weight rows at an odd pitch (conflict-free) and one activation row (broadcast)
    for (int i = tid; i < MC_ROWS * dout; i += MC_THREADS) {
      const int r = i / dout, j = i - r * dout;
      const float* w = ws + j * wp;
      const float* hr = h + r * MC_HPITCH;
      float a0 = 0.f, a1 = 0.f;
      int k = 0;
      for (; k + 1 < din; k += 2) {
        a0 = fmaf(hr[k], w[k], a0);
        a1 = fmaf(hr[k + 1], w[k + 1], a1);
      }
      if (k < din) a0 = fmaf(hr[k], w[k], a0);
      float v = a0 + a1 + (L.b ? __ldg(L.b + j) : 0.f);
      v = cvad_act(v, L.act);
      if (r < nr) {
        if (L.mask) v *= __ldg(L.mask + (r0 + r) * dout + j) * L.mask_scale;
        if (L.save) L.save[(r0 + r) * dout + j] = v;
        if (last) out[(r0 + r) * dout + j] = v;
      }
      hn[r * MC_HPITCH + j] = v;
    }
    __syncthreads();
    cur ^= 1;
  }
}

__global__ void __launch_bounds__(MC_THREADS) mlp_chain_bwd_kernel(const McChain ch, const float* __restrict__ dy, float* __restrict__ dx) {
  cvad_pdl_enter();
  extern __shared__ float mc_smem[];
  float* gb[2] = {mc_smem, mc_smem + MC_ROWS * MC_HPITCH};
  float* ws = mc_smem + 2 * MC_ROWS * MC_HPITCH;
  const int tid = threadIdx.x;
  const long long r0 = (long long)blockIdx.x * MC_ROWS;
  const int nr = (int)(ch.rows - r0 < MC_ROWS ? ch.rows - r0 : MC_ROWS);
  const int dl = ch.L[ch.n - 1].dout;
  for (int i = tid; i < MC_ROWS * dl; i += MC_THREADS) {
    const int r = i / dl, j = i - r * dl;
    gb[0][r * MC_HPITCH + j] = r < nr ? __ldg(dy + (r0 + r) * dl + j) : 0.f;
  }
  __syncthreads();
  int cur = 0;
  for (int l = ch.n - 1; l >= 0; --l) {
    const McLayer L = ch.L[l];
    const int din = L.din, dout = L.dout;
    float* g = gb[cur];
    float* gp = gb[cur ^ 1];
    const bool need_prev = l > 0 || dx != nullptr;
    if (need_prev) stage_weights(L.W, din, dout, din, ws);            // natural pitch: the reads below run along k
    // dz = g * mask*scale * act'(y); y is the stored post-mask output (sigmoid / tanh: the pre-mask value is recovered, as cvad_act_mask_bwd_f32 does)
    for (int i = tid; i < MC_ROWS * dout; i += MC_THREADS) {
      const int r = i / dout, j = i - r * dout;
      float v = 0.f;
      if (r < nr) {
        v = g[r * MC_HPITCH + j];
        float yy = L.act != ACT_NONE ? __ldg(L.save + (r0 + r) * dout + j) : 0.f;
        if (L.mask) {
          const float mk = __ldg(L.mask + (r0 + r) * dout + j) * L.mask_scale;
          v *= mk;
          if (L.act == ACT_SIGMOID || L.act == ACT_TANH) yy = mk != 0.f ? yy / mk : 0.f;
        }
        v *= cvad_act_grad_from_out(yy, L.act);
        L.dz[(r0 + r) * dout + j] = v;
      }
      g[r * MC_HPITCH + j] = v;
    }
    __syncthreads();
    if (!need_prev) break;
    // gp[r][k] = sum_j dz[r][j] W[j][k]   ((row, k) items, k fastest)
    for (int i = tid; i < MC_ROWS * din; i += MC_THREADS) {
      const int r = i / din, k = i - r * din;
      const float* gr = g + r * MC_HPITCH;
      float a0 = 0.f, a1 = 0.f;
      int j = 0;
      for (; j + 1 < dout; j += 2) {
        a0 = fmaf(gr[j], ws[j * din + k], a0);
        a1 = fmaf(gr[j + 1], ws[(j + 1) * din + k], a1);
      }
      if (j < dout) a0 = fmaf(gr[j], ws[j * din + k], a0);
      gp[r * MC_HPITCH + k] = a0 + a1;
    }
    __syncthreads();
    cur ^= 1;
  }
  if (dx) {
    const int d0 = ch.L[0].din;
    for (int i = tid; i < nr * d0; i += MC_THREADS) {
      const int r = i / d0, k = i - r * d0;
      dx[(r0 + r) * d0 + k] = gb[cur][r * MC_HPITCH + k];
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
    if ((long long)(dims[l] | 1) * dims[l + 1] > MC_MAXW) return (int)cudaErrorInvalidValue;
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
  static size_t configured[CVAD_MAX_DEVICES] = {};
  cudaError_t ce = cvad_ensure_dyn_smem(mlp_chain_fwd_kernel, MC_SMEM, configured);
  if (ce != cudaSuccess) return (int)ce;
  ce = cvad_launch_pdl(mlp_chain_fwd_kernel, dim3(grid), dim3(MC_THREADS), MC_SMEM, (cudaStream_t)stream, ch, x, out);
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
  static size_t configured[CVAD_MAX_DEVICES] = {};
  cudaError_t ce = cvad_ensure_dyn_smem(mlp_chain_bwd_kernel, MC_SMEM, configured);
  if (ce != cudaSuccess) return (int)ce;
  ce = cvad_launch_pdl(mlp_chain_bwd_kernel, dim3(grid), dim3(MC_THREADS), MC_SMEM, (cudaStream_t)stream, ch, dy, dx);
  return (int)ce;
}
