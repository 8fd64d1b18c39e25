// Fused loss kernels (K11) and the fused clip + Adam/AdamW step over a flat parameter arena (K12).
//
//   cvad_mb_loss_f32   s2:135-205  focal BCE on pseudo-labels + acyclicity + sparsity + consistency (O(B^2) pairs)
//                                  + structure hinge; forward value, the 7 reported components and d/dscores, d/dadj
//   cvad_bce_loss_f32  mc3:240,287 nn.BCELoss(mean) forward + gradient, non-finite flag (mc3:282-292)
//   cvad_ma_loss_f32   cad:649-662 0.4*CE(softmax probs) + 0.3*MSE(final) + 0.2*MSE(causal) + 0.1*KL, with gradients
//   cvad_sumsq_f32 / cvad_adam_flat_f32 / cvad_opt_finish   torch clip_grad_norm_ + AdamW/Adam semantics
//                                  (s2:115-119,236-238; cad:615-618,665-667; mc3:229-234,298-311), no host sync:
//                                  the norm, the non-finite test and the skip decision all stay on the device.
#include "common.cuh"
#include "cvad_b200.h"

namespace {

__device__ __forceinline__ float sgnf(float v) { return (v > 0.f) - (v < 0.f); }

// ------------------------------------------------------------------------------------------------ M-B loss
// ws layout (floats): [0,256) sum_b adj ; [256,256+B) focal ; [..+B) dfocal/ds ; [..+B) edges ; [..+B) rowdist
__global__ void mb_loss_partials_kernel(const float* __restrict__ scores, const float* __restrict__ adj, const float* __restrict__ pseudo,
                                        int B, float* __restrict__ ws, float* __restrict__ dadj) {
  __shared__ float sh[32];
  const int b = blockIdx.x, e = threadIdx.x;   // 256 threads = 16x16 adjacency entries
  float* abar = ws;
  float* focal = ws + 256;
  float* dfocal = focal + B;
  float* edges = dfocal + B;
  float* rowdist = edges + B;
  const float a = adj[(long long)b * 256 + e];
  atomicAdd(abar + e, a);
  float ecount = block_sum(a > 0.1f ? 1.f : 0.f, sh);
  float dist = 0.f, sg = 0.f;
  const bool normal = pseudo[b] == 0.f;
  if (normal) {
    for (int j = 0; j < B; ++j) {
      if (j == b || pseudo[j] != 0.f) continue;
      float d = a - __ldg(adj + (long long)j * 256 + e);
      dist += fabsf(d);
      sg += sgnf(d);
    }
  }
  dist = block_sum(dist, sh);
  dadj[(long long)b * 256 + e] = sg;
  if (e == 0) {
    edges[b] = ecount;
    rowdist[b] = dist * (1.f / 256.f);
    // focal BCE (alpha 0.25, gamma 2), log terms clamped at -100 like F.binary_cross_entropy
    const float s = scores[b], y = pseudo[b];
    float ls = logf(s), l1s = logf(1.f - s);
    float dls = 1.f / s, dl1s = -1.f / (1.f - s);
    if (ls < -100.f) { ls = -100.f; dls = 0.f; }
    if (l1s < -100.f) { l1s = -100.f; dl1s = 0.f; }
    const float ce = -(y * ls + (1.f - y) * l1s);
    const float dce = -(y * dls + (1.f - y) * dl1s);
    const float pt = expf(-ce);
    const float om = 1.f - pt;
    focal[b] = 0.25f * om * om * ce;
    dfocal[b] = 0.25f * (om * om + 2.f * om * pt * ce) * dce;
  }
}

__global__ void mb_loss_finish_kernel(const float* __restrict__ pseudo, int B, const float* __restrict__ ws, float w_anom, float w_causal,
                                      float w_sparse, float w_cons, float* __restrict__ out, float* __restrict__ dscores,
                                      float* __restrict__ dadj, float* __restrict__ flag) {
  __shared__ float sh[32];
  const int b = blockIdx.x, e = threadIdx.x;
  const float* abar = ws;
  const float* focal = ws + 256;
  const float* dfocal = focal + B;
  const float* edges = dfocal + B;
  const float* rowdist = edges + B;
  float f = 0.f, ed = 0.f, rd = 0.f, nn = 0.f;
  for (int j = e; j < B; j += blockDim.x) {
    f += focal[j];
    ed += edges[j];
    rd += rowdist[j];
    nn += pseudo[j] == 0.f ? 1.f : 0.f;
  }
  f = block_sum(f, sh);
  ed = block_sum(ed, sh);
  rd = block_sum(rd, sh);
  nn = block_sum(nn, sh);
  const float invB = 1.f / (float)B;
  const int i = e >> 4, j = e & 15;
  const float aij = abar[e] * invB, aji = abar[j * 16 + i] * invB;
  const float acyc = block_sum(aij * aji, sh);   // trace(Abar @ Abar)
  const float npairs = nn * (nn - 1.f) * 0.5f;
  float cons = 0.f, csign = 0.f;
  if (nn > 1.f) {
    float avg = (rd * 0.5f) / npairs;
    cons = fabsf(avg - 0.1f);
    csign = sgnf(avg - 0.1f);
  }
  float g = w_causal * 2.f * aji * invB;
  if (nn > 1.f && pseudo[b] == 0.f) g += w_cons * csign * dadj[(long long)b * 256 + e] / (npairs * 256.f);
  dadj[(long long)b * 256 + e] = g;
  if (e == 0) dscores[b] = w_anom * dfocal[b] * invB;
  if (b == 0 && e == 0) {
    const float anomaly = f * invB;
    const float ratio = ed / ((float)B * 256.f);
    const float spars = fabsf(ratio - 0.3f);
    float st = 0.f;
    if (ed < 10.f) st = (10.f - ed) * 0.01f;
    else if (ed > 40.f) st = (ed - 40.f) * 0.01f;
    const float total = w_anom * anomaly + w_causal * acyc + w_sparse * spars + w_cons * cons + 0.01f * st;
    out[0] = total; out[1] = anomaly; out[2] = acyc; out[3] = spars; out[4] = cons; out[5] = st; out[6] = ed; out[7] = ratio;
    if (flag && !(fabsf(total) <= 3.0e38f)) *flag = 1.f;   // NaN/Inf loss -> the step is skipped (s2:230-232)
  }
}

// ------------------------------------------------------------------------------------------------ BCE (M-C)
__global__ void bce_loss_kernel(const float* __restrict__ s, const float* __restrict__ y, int B, float* __restrict__ out,
                                float* __restrict__ ds, float* __restrict__ flag) {
  __shared__ float sh[32];
  float acc = 0.f, bad = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    float sv = s[b], yv = y[b];
    if (!(fabsf(sv) <= 3.0e38f)) bad = 1.f;
    float ls = logf(sv), l1s = logf(1.f - sv);
    float dls = 1.f / sv, dl1s = -1.f / (1.f - sv);
    if (ls < -100.f) { ls = -100.f; dls = 0.f; }
    if (l1s < -100.f) { l1s = -100.f; dl1s = 0.f; }
    acc += -(yv * ls + (1.f - yv) * l1s);
    if (ds) ds[b] = -(yv * dls + (1.f - yv) * dl1s) / (float)B;
  }
  acc = block_sum(acc, sh);
  bad = block_sum(bad, sh);
  if (threadIdx.x == 0) {
    float loss = acc / (float)B;
    out[0] = loss;
    if (flag && (bad > 0.f || !(fabsf(loss) <= 3.0e38f))) *flag = 1.f;
  }
}

// ------------------------------------------------------------------------------------------------ M-A 4-term loss
__global__ void ma_loss_kernel(const float* __restrict__ probs, const float* __restrict__ fin, const float* __restrict__ causal,
                               const float* __restrict__ kl, const long long* __restrict__ labels, int B, float* __restrict__ out,
                               float* __restrict__ dprobs, float* __restrict__ dfin, float* __restrict__ dcausal,
                               float* __restrict__ dkl, float* __restrict__ flag) {
  cvad_pdl_enter();
  __shared__ float sh[32];
  float ce = 0.f, mf = 0.f, mc = 0.f, ks = 0.f;
  const float invB = 1.f / (float)B;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float p0 = probs[2 * b], p1 = probs[2 * b + 1];
    const int lab = (int)labels[b];
    const float y = (float)lab;
    const float mx = fmaxf(p0, p1);
    const float e0 = expf(p0 - mx), e1 = expf(p1 - mx);
    const float lse = mx + logf(e0 + e1);
    ce += lse - (lab ? p1 : p0);                      // CrossEntropy applied to probabilities (cad:537, 649)
    const float q0 = e0 / (e0 + e1), q1 = e1 / (e0 + e1);
    const float df = fin[b] - y, dc = causal[b] - y;
    mf += df * df;
    mc += dc * dc;
    const float k = kl[b];
    const bool fin_k = fabsf(k) <= 3.0e38f;           // cad:653 keeps finite terms only
    if (fin_k) ks += k;
    if (dprobs) {
      dprobs[2 * b] = 0.4f * (q0 - (lab == 0 ? 1.f : 0.f)) * invB;
      dprobs[2 * b + 1] = 0.4f * (q1 - (lab == 1 ? 1.f : 0.f)) * invB;
      dfin[b] = 0.3f * 2.f * df * invB;
      dcausal[b] = 0.2f * 2.f * dc * invB;
      dkl[b] = fin_k ? 0.1f * invB : 0.f;
    }
  }
  ce = block_sum(ce, sh);
  mf = block_sum(mf, sh);
  mc = block_sum(mc, sh);
  ks = block_sum(ks, sh);
  if (threadIdx.x == 0) {
    ce *= invB; mf *= invB; mc *= invB; ks *= invB;
    const float total = 0.4f * ce + 0.3f * mf + 0.2f * mc + 0.1f * ks;
    out[0] = total; out[1] = ce; out[2] = mf; out[3] = mc; out[4] = ks;
    if (flag && !(fabsf(total) <= 3.0e38f)) *flag = 1.f;
  }
}

// ------------------------------------------------------------------------------------------------ optimizer
__global__ void sumsq_kernel(const float* __restrict__ g, long long n, float scale, cvad_opt_state* __restrict__ st) {
  __shared__ double shd[32];
  double s = 0.0, bad = 0.0;
  const long long n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = __ldg(g4 + i);
    float a = v.x * scale, b = v.y * scale, c = v.z * scale, d = v.w * scale;
    float q = a * a + b * b + c * c + d * d;
    if (!(q <= 3.0e38f)) bad += 1.0;
    s += q;
  }
  s = block_sum_d(s, shd);
  bad = block_sum_d(bad, shd);
  if (threadIdx.x == 0) {
    atomicAdd(&st->gradsq, s);
    if (bad > 0.0) atomicAdd(&st->nonfinite, bad);
  }
}

// one block = 1024 consecutive arena elements (all of one tensor; tensors are padded to 1024)
__global__ void __launch_bounds__(256) adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, const int* __restrict__ block_slot,
                                                        const cvad_opt_state* __restrict__ st, float grad_scale, float lr, float beta1,
                                                        float beta2, float eps, float wd, int decoupled, int clip_mode, float max_norm,
                                                        float clip_threshold, int nan_mode) {
  const int slot = block_slot[blockIdx.x];
  if (slot < 0) return;                                   // header / padding block
  const float* header = g;                                // first 16 floats of the gradient arena
  const bool nonfinite = st->nonfinite > 0.0 || header[0] != 0.f;
  if (nan_mode == 1 && nonfinite) return;                 // whole step skipped
  if (slot > 0 && !(header[slot] > 0.f)) return;          // this group received no gradient (torch: grad is None)
  const float norm = (float)sqrt(st->gradsq);
  float coef = 1.f;
  if (clip_mode == 1 || (clip_mode == 2 && norm > clip_threshold)) coef = fminf(1.f, max_norm / (norm + 1e-6f));
  coef *= grad_scale;
  const double t = (double)(st->step[slot] + 1);
  const float bc1 = (float)(1.0 - pow((double)beta1, t));
  const float bc2s = (float)sqrt(1.0 - pow((double)beta2, t));
  if (st->lr_device > 0.0) lr = (float)st->lr_device;
  const float step_size = lr / bc1;
  const long long base = ((long long)blockIdx.x * 256 + threadIdx.x) * 4;
  float4 pv = *reinterpret_cast<float4*>(p + base);
  const float4 gv = *reinterpret_cast<const float4*>(g + base);
  float4 mv = *reinterpret_cast<float4*>(m + base);
  float4 vv = *reinterpret_cast<float4*>(v + base);
  float pp[4] = {pv.x, pv.y, pv.z, pv.w}, gg[4] = {gv.x, gv.y, gv.z, gv.w}, mm[4] = {mv.x, mv.y, mv.z, mv.w},
        vq[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float gk = gg[k] * coef;
    if (decoupled) pp[k] *= (1.f - lr * wd);
    else gk = fmaf(wd, pp[k], gk);
    mm[k] = beta1 * mm[k] + (1.f - beta1) * gk;
    vq[k] = beta2 * vq[k] + (1.f - beta2) * gk * gk;
    const float denom = sqrtf(vq[k]) / bc2s + eps;
    pp[k] -= step_size * (mm[k] / denom);
  }
  *reinterpret_cast<float4*>(p + base) = make_float4(pp[0], pp[1], pp[2], pp[3]);
  *reinterpret_cast<float4*>(m + base) = make_float4(mm[0], mm[1], mm[2], mm[3]);
  *reinterpret_cast<float4*>(v + base) = make_float4(vq[0], vq[1], vq[2], vq[3]);
}

__global__ void opt_finish_kernel(cvad_opt_state* __restrict__ st, const float* __restrict__ header, int nan_mode) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const bool nonfinite = st->nonfinite > 0.0 || header[0] != 0.f;
    st->last_gradnorm = sqrt(st->gradsq);
    if (nan_mode == 1 && nonfinite) {
      st->skipped += 1;
    } else {
      st->step[0] += 1;
      for (int k = 1; k < 8; ++k)
        if (header[k] > 0.f) st->step[k] += 1;
    }
    st->gradsq = 0.0;
    st->nonfinite = 0.0;
  }
}

// 16-byte stores over the aligned body, scalars over the (at most 3 + 3) elements around it
__global__ void fill_kernel(float* __restrict__ x, long long n, float v) {
  const long long head = (4 - (((uintptr_t)x >> 2) & 3)) & 3;          // elements before the first 16-byte boundary
  const long long h = head < n ? head : n;
  const long long n4 = (n - h) >> 2;
  float4* x4 = reinterpret_cast<float4*>(x + h);
  const float4 v4 = make_float4(v, v, v, v);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) x4[i] = v4;
  if (blockIdx.x == 0) {
    for (long long i = threadIdx.x; i < h; i += blockDim.x) x[i] = v;
    for (long long i = h + 4 * n4 + threadIdx.x; i < n; i += blockDim.x) x[i] = v;
  }
}

// out[i] = a * x[i*xs] + b * y[i*ys]
__global__ void lincomb2_kernel(float* __restrict__ out, const float* __restrict__ x, long long xs, float a, const float* __restrict__ y,
                                long long ys, float b, long long n) {
  cvad_pdl_enter();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = a * x[i * xs] + b * y[i * ys];
}

}  // namespace

CVAD_API long long cvad_mb_loss_ws_floats(int B) { return 256 + 4LL * B; }

CVAD_API int cvad_mb_loss_f32(const float* scores, const float* adj, const float* pseudo, int B, float w_anom, float w_causal,
                              float w_sparse, float w_cons, float* ws, float* out8, float* dscores, float* dadj, float* nonfinite_flag,
                              void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (B <= 0) return 0;
  cudaError_t e = cudaMemsetAsync(ws, 0, 256 * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  mb_loss_partials_kernel<<<B, 256, 0, st>>>(scores, adj, pseudo, B, ws, dadj);
  CVAD_LAUNCH_CHECK();
  mb_loss_finish_kernel<<<B, 256, 0, st>>>(pseudo, B, ws, w_anom, w_causal, w_sparse, w_cons, out8, dscores, dadj, nonfinite_flag);
  CVAD_LAUNCH_CHECK();
  return 0;
}

CVAD_API int cvad_bce_loss_f32(const float* scores, const float* targets, int B, float* out1, float* dscores, float* nonfinite_flag,
                               void* stream) {
  if (B <= 0) return 0;
  bce_loss_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(scores, targets, B, out1, dscores, nonfinite_flag);
  CVAD_LAUNCH_CHECK();
  return 0;
}

CVAD_API int cvad_ma_loss_f32(const float* probs, const float* final_scores, const float* causal_scores, const float* kl,
                              const long long* labels, int B, float* out5, float* dprobs, float* dfinal, float* dcausal, float* dkl,
                              float* nonfinite_flag, void* stream) {
  if (B <= 0) return 0;
  cvad_launch_pdl(ma_loss_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, probs, final_scores, causal_scores, kl, labels, B, out5, dprobs, dfinal, dcausal,
                                                       dkl, nonfinite_flag);
  CVAD_LAUNCH_CHECK();
  return 0;
}

CVAD_API int cvad_sumsq_f32(const float* g, long long n, float scale, cvad_opt_state* state, void* stream) {
  if (n <= 0) return 0;
  long long n4 = n / 4;
  int blocks = (int)((n4 + 255) / 256);
  int cap = 4 * cvad_num_sms();
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  sumsq_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(g, n, scale, state);
  CVAD_LAUNCH_CHECK();
  return 0;
}

CVAD_API int cvad_adam_flat_f32(float* p, const float* g, float* m, float* v, long long n, const int* block_slot, cvad_opt_state* state,
                                float grad_scale, float lr, float beta1, float beta2, float eps, float weight_decay, int decoupled,
                                int clip_mode, float max_norm, float clip_threshold, int nan_mode, void* stream) {
  if (n <= 0 || (n % 1024) != 0) return (int)cudaErrorInvalidValue;
  cudaStream_t st = (cudaStream_t)stream;
  adam_flat_kernel<<<(unsigned)(n / 1024), 256, 0, st>>>(p, g, m, v, block_slot, state, grad_scale, lr, beta1, beta2, eps, weight_decay,
                                                         decoupled, clip_mode, max_norm, clip_threshold, nan_mode);
  CVAD_LAUNCH_CHECK();
  opt_finish_kernel<<<1, 32, 0, st>>>(state, g, nan_mode);
  CVAD_LAUNCH_CHECK();
  return 0;
}

CVAD_API int cvad_fill_f32(float* x, long long n, float value, void* stream) {
  if (n <= 0) return 0;
  int blocks = (int)((n / 4 + 255) / 256);
  if (blocks > 8 * cvad_num_sms()) blocks = 8 * cvad_num_sms();
  if (blocks < 1) blocks = 1;
  fill_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, n, value);
  CVAD_LAUNCH_CHECK();
  return 0;
}

CVAD_API int cvad_lincomb2_f32(float* out, const float* x, long long xs, float a, const float* y, long long ys, float b, long long n,
                               void* stream) {
  if (n <= 0) return 0;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 8 * cvad_num_sms()) blocks = 8 * cvad_num_sms();
  cvad_launch_pdl(lincomb2_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, out, x, xs, a, y, ys, b, n);
  CVAD_LAUNCH_CHECK();
  return 0;
}
