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
// A CTA first stages the weight matrices of ALL layers of the chain in shared memory -- one cooperative, coalesced pass with every
// thread keeping several 16-byte loads in flight, issued BEFORE the programmatic-dependent-launch wait: parameters do not depend on the
// stream predecessor, so the staging overlaps the tail of the previous kernel (two earlier versions, weights read from L2 inside the
// dot-product loops / staged layer by layer, were latency-bound at 80 / 20 us per chain).  Then the layers run back to back out of shared
// memory, separated only by __syncthreads: a thread owns one output column (forward) or one input column (backward) for four of the
// CTA's eight rows -- weights are read once per four rows, activations as 16-byte broadcasts.  Forward h'[r][j] = sum_k h[r][k] W[j][k]
// (weight rows padded to an odd pitch: conflict-free across j); backward dh[r][k] = sum_j dz[r][j] W[j][k] (consecutive k: conflict-free).
// A chain must fit: widths <= 256 and all its (padded) weight matrices together <= 50 000 floats (the 6144 -> 512 -> 256 layers of the
// detector / classifier stay split-K GEMMs, which spread their weights over many SMs instead of streaming 0.5 MB through each of 4..64 CTAs).
#include "common.cuh"
#include "cvad_b200.h"

namespace {

constexpr int MC_MAX_LAYERS = 8;
constexpr int MC_ROWS = 8;                             // rows per CTA
constexpr int MC_RB = 4;                               // rows per thread
constexpr int MC_THREADS = 256;
constexpr int MC_MAXDIM = 256;
constexpr int MC_MAXW = 50000;                         // floats of all staged weight matrices of a chain (200 KB)
constexpr int MC_HPITCH = MC_MAXDIM + 4;               // multiple of 4: rows are read as float4

struct McLayer {
  const float* W;         // (dout, din) row-major, nn.Linear layout
  const float* b;         // (dout) or NULL
  const float* mask;      // (rows, dout) keep-mask or NULL
  float* save;            // forward: (rows, dout) stored output (post activation, post mask) or NULL; backward: the same, read-only
  float* dz;              // backward: (rows, dout) gradient w.r.t. the pre-activation, written for the weight-gradient GEMMs
  float mask_scale;
  int din, dout, act;
  int woff;               // offset (floats) of this layer's staged weights in shared memory
};

struct McChain {
  int n;
  long long rows;
  McLayer L[MC_MAX_LAYERS];
};

// stage W (dout, din) row-major into shared memory with row pitch wp
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
  extern __shared__ __align__(16) float mc_smem[];
  float* hb[2] = {mc_smem, mc_smem + MC_ROWS * MC_HPITCH};
  float* wsm = mc_smem + 2 * MC_ROWS * MC_HPITCH;
  const int tid = threadIdx.x;
  // parameters first: they do not depend on the kernel in front of this one
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  for (int l = 0; l < ch.n; ++l) stage_weights(ch.L[l].W, ch.L[l].din, ch.L[l].dout, ch.L[l].din | 1, wsm + ch.L[l].woff);
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const long long r0 = (long long)blockIdx.x * MC_ROWS;
  const int nr = (int)(ch.rows - r0 < MC_ROWS ? ch.rows - r0 : MC_ROWS);
  const int d0 = ch.L[0].din;
  for (int i = tid; i < MC_ROWS * MC_HPITCH; i += MC_THREADS) {
    const int r = i / MC_HPITCH, k = i - r * MC_HPITCH;
    hb[0][i] = (r < nr && k < d0) ? __ldg(x + (r0 + r) * d0 + k) : 0.f;       // zero fill: the float4 reads below run to a multiple of 4
    hb[1][i] = 0.f;
  }
  __syncthreads();
  int cur = 0;
  for (int l = 0; l < ch.n; ++l) {
    const McLayer L = ch.L[l];
    const int din = L.din, dout = L.dout, wp = din | 1;
    const float* h = hb[cur];
    float* hn = hb[cur ^ 1];
    const float* ws = wsm + L.woff;
    const bool last = l == ch.n - 1;
    const int din4 = (din + 3) & ~3;
    // item = (row block of MC_RB rows, column j), column fastest
    for (int i = tid; i < (MC_ROWS / MC_RB) * dout; i += MC_THREADS) {
      const int rb = i / dout, j = i - rb * dout;
      const float* w = ws + j * wp;
      float acc[MC_RB];
#pragma unroll
      for (int r = 0; r < MC_RB; ++r) acc[r] = 0.f;
      for (int k = 0; k < din4; k += 4) {
        float wv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) wv[e] = k + e < din ? w[k + e] : 0.f;
#pragma unroll
        for (int r = 0; r < MC_RB; ++r) {
          const float4 hv = *reinterpret_cast<const float4*>(h + (rb * MC_RB + r) * MC_HPITCH + k);
          acc[r] = fmaf(hv.x, wv[0], acc[r]);
          acc[r] = fmaf(hv.y, wv[1], acc[r]);
          acc[r] = fmaf(hv.z, wv[2], acc[r]);
          acc[r] = fmaf(hv.w, wv[3], acc[r]);
        }
      }
      const float bj = L.b ? __ldg(L.b + j) : 0.f;
#pragma unroll
      for (int r = 0; r < MC_RB; ++r) {
        const int row = rb * MC_RB + r;
        float v = cvad_act(acc[r] + bj, L.act);
        if (row < nr) {
          if (L.mask) v *= __ldg(L.mask + (r0 + row) * dout + j) * L.mask_scale;
          if (L.save) L.save[(r0 + row) * dout + j] = v;
          if (last) out[(r0 + row) * dout + j] = v;
        } else {
          v = 0.f;
        }
        hn[row * MC_HPITCH + j] = v;
      }
    }
    __syncthreads();
    // columns dout .. dout4 of the new activations must read as zero for the next layer's float4 loop
    if (!last) {
      const int dn4 = (dout + 3) & ~3;
      for (int i = tid; i < MC_ROWS * (dn4 - dout); i += MC_THREADS) hn[(i / (dn4 - dout)) * MC_HPITCH + dout + i % (dn4 - dout)] = 0.f;
      if (dn4 != dout) __syncthreads();
    }
    cur ^= 1;
  }
}

__global__ void __launch_bounds__(MC_THREADS) mlp_chain_bwd_kernel(const McChain ch, const float* __restrict__ dy, float* __restrict__ dx) {
  extern __shared__ __align__(16) float mc_smem[];
  float* gb[2] = {mc_smem, mc_smem + MC_ROWS * MC_HPITCH};
  float* wsm = mc_smem + 2 * MC_ROWS * MC_HPITCH;
  const int tid = threadIdx.x;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  for (int l = ch.n - 1; l >= (dx ? 0 : 1); --l) stage_weights(ch.L[l].W, ch.L[l].din, ch.L[l].dout, ch.L[l].din, wsm + ch.L[l].woff);
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const long long r0 = (long long)blockIdx.x * MC_ROWS;
  const int nr = (int)(ch.rows - r0 < MC_ROWS ? ch.rows - r0 : MC_ROWS);
  const int dl = ch.L[ch.n - 1].dout;
  for (int i = tid; i < MC_ROWS * MC_HPITCH; i += MC_THREADS) {
    const int r = i / MC_HPITCH, j = i - r * MC_HPITCH;
    gb[0][i] = (r < nr && j < dl) ? __ldg(dy + (r0 + r) * dl + j) : 0.f;
    gb[1][i] = 0.f;
  }
  __syncthreads();
  int cur = 0;
  for (int l = ch.n - 1; l >= 0; --l) {
    const McLayer L = ch.L[l];
    const int din = L.din, dout = L.dout;
    float* g = gb[cur];
    float* gp = gb[cur ^ 1];
    const float* ws = wsm + L.woff;
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
    if (l == 0 && dx == nullptr) break;
    // gp[r][k] = sum_j dz[r][j] W[j][k]: item = (row block, column k), k fastest; columns >= dout of g are zero (fill above / below)
    const int dout4 = (dout + 3) & ~3;
    for (int i = tid; i < (MC_ROWS / MC_RB) * din; i += MC_THREADS) {
      const int rb = i / din, k = i - rb * din;
      float acc[MC_RB];
#pragma unroll
      for (int r = 0; r < MC_RB; ++r) acc[r] = 0.f;
      for (int j = 0; j < dout4; j += 4) {
        float wv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) wv[e] = j + e < dout ? ws[(j + e) * din + k] : 0.f;
#pragma unroll
        for (int r = 0; r < MC_RB; ++r) {
          const float4 gv = *reinterpret_cast<const float4*>(g + (rb * MC_RB + r) * MC_HPITCH + j);
          acc[r] = fmaf(gv.x, wv[0], acc[r]);
          acc[r] = fmaf(gv.y, wv[1], acc[r]);
          acc[r] = fmaf(gv.z, wv[2], acc[r]);
          acc[r] = fmaf(gv.w, wv[3], acc[r]);
        }
      }
#pragma unroll
      for (int r = 0; r < MC_RB; ++r) gp[(rb * MC_RB + r) * MC_HPITCH + k] = acc[r];
    }
    // the next layer reads gp as its g with width din: keep the columns up to the next multiple of 4 zero
    const int din4 = (din + 3) & ~3;
    for (int i = tid; i < MC_ROWS * (din4 - din); i += MC_THREADS) gp[(i / (din4 - din)) * MC_HPITCH + din + i % (din4 - din)] = 0.f;
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

// returns 0 and the dynamic shared-memory size, or an error; `fwd` selects the padded (odd) weight pitch of the forward kernel
int fill_chain(McChain& ch, size_t& smem, bool fwd, long long rows, int n_layers, const int* dims, const int* acts, const void* const* weights,
               const void* const* biases, const void* const* masks, const float* mask_scales, void* const* saves, void* const* dzs) {
  if (n_layers < 1 || n_layers > MC_MAX_LAYERS || rows < 0) return (int)cudaErrorInvalidValue;
  ch.n = n_layers;
  ch.rows = rows;
  long long off = 0;
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
    L.woff = (int)off;
    off += (long long)((fwd ? (dims[l] | 1) : dims[l]) * dims[l + 1] + 3) & ~3LL;
  }
  if (off > MC_MAXW) return (int)cudaErrorInvalidValue;
  smem = (size_t)(2 * MC_ROWS * MC_HPITCH + off) * sizeof(float);
  return 0;
}

}  // namespace

CVAD_API int cvad_mlp_chain_fwd_f32(const float* x, long long rows, int n_layers, const int* dims, const int* acts, const void* const* weights,
                                    const void* const* biases, const void* const* masks, const float* mask_scales, void* const* saves, float* out,
                                    void* stream) {
  if (rows == 0) return 0;
  McChain ch;
  size_t smem = 0;
  int e = fill_chain(ch, smem, true, rows, n_layers, dims, acts, weights, biases, masks, mask_scales, saves, nullptr);
  if (e) return e;
  const unsigned grid = (unsigned)((rows + MC_ROWS - 1) / MC_ROWS);
  static size_t configured[CVAD_MAX_DEVICES] = {};
  cudaError_t ce = cvad_ensure_dyn_smem(mlp_chain_fwd_kernel, smem, configured);
  if (ce != cudaSuccess) return (int)ce;
  ce = cvad_launch_pdl(mlp_chain_fwd_kernel, dim3(grid), dim3(MC_THREADS), smem, (cudaStream_t)stream, ch, x, out);
  return (int)ce;
}

CVAD_API int cvad_mlp_chain_bwd_f32(const float* dy, long long rows, int n_layers, const int* dims, const int* acts, const void* const* weights,
                                    const void* const* masks, const float* mask_scales, const void* const* saves, void* const* dzs, float* dx,
                                    void* stream) {
  if (rows == 0) return 0;
  McChain ch;
  size_t smem = 0;
  int e = fill_chain(ch, smem, false, rows, n_layers, dims, acts, weights, nullptr, masks, mask_scales, (void* const*)saves, dzs);
  if (e) return e;
  for (int l = 0; l < n_layers; ++l)
    if (!ch.L[l].dz || (ch.L[l].act != ACT_NONE && !ch.L[l].save)) return (int)cudaErrorInvalidValue;
  const unsigned grid = (unsigned)((rows + MC_ROWS - 1) / MC_ROWS);
  static size_t configured[CVAD_MAX_DEVICES] = {};
  cudaError_t ce = cvad_ensure_dyn_smem(mlp_chain_bwd_kernel, smem, configured);
  if (ce != cudaSuccess) return (int)ce;
  ce = cvad_launch_pdl(mlp_chain_bwd_kernel, dim3(grid), dim3(MC_THREADS), smem, (cudaStream_t)stream, ch, dy, dx);
  return (int)ce;
}
