// Dense, masked, batched kernels for the causal branch of M-A (causal_anomaly_detection.py:160-502): the reference
// walks ragged Python lists (B*T*5 scalar comparisons with a host sync each, batch-1 MLP / GRU calls per track);
// here every clip carries 5 padded track slots plus a per-clip track count and each stage is one launch.
//
//   det_decode      cad:198-228  sigmoid-scale, validity window, in-order compaction, fallback box
//   traj_assemble   cad:251-269  [box | reid] rows, zero padding rows, permute to (B,5,T,68); tracks/clip = max_t count
//   gru_fwd/bwd     cad:284,298  nn.GRU(68->64) recurrence (input projection is a GEMM done by the caller), last state
//   reparam_kl      cad:328-347  z = mu + eps*exp(0.5*logvar), KL per track, mean over the clip's tracks
//   pair_concat     cad:385      [node_i | node_j] rows for the edge MLP
//   adj_assemble    cad:380-390  6x6 adjacency from the 5x5 edge probabilities, i != j, tracks < count
//   structured      cad:420      (adj @ z^T)^T
//   scorer_inputs   cad:469-494  masked means over tracks, |cur - pred|, the three concatenations
//   softmax_rows    cad:537      nn.Softmax(dim=-1)
#include "common.cuh"
#include "cvad_b200.h"

namespace {

constexpr int MAXDET = 5;
constexpr int HID = 64;       // GRU hidden size (cad:279)
constexpr int NF = 6;         // causal factors (cad:511)

__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }

// ------------------------------------------------------------------------------------------------ detector decode
__global__ void det_decode_kernel(const float* __restrict__ raw, long long R, float* __restrict__ box, int* __restrict__ cnt,
                                  int* __restrict__ src, float* __restrict__ flag) {
  cvad_pdl_enter();
  const float scale[4] = {360.f, 240.f, 80.f, 120.f}, off[4] = {0.f, 0.f, 15.f, 25.f};
  const float lo[4] = {10.f, 10.f, 10.f, 20.f}, hi[4] = {350.f, 230.f, 100.f, 150.f};
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < R; r += (long long)gridDim.x * blockDim.x) {
    int n = 0;
    for (int k = 0; k < MAXDET; ++k) {
      float v[4];
      bool ok = true;
      for (int c = 0; c < 4; ++c) {
        v[c] = sigmoidf_(raw[(r * MAXDET + k) * 4 + c]) * scale[c] + off[c];
        ok = ok && v[c] >= lo[c] && v[c] <= hi[c];
      }
      if (ok) {
        for (int c = 0; c < 4; ++c) box[(r * MAXDET + n) * 4 + c] = v[c];
        src[r * MAXDET + n] = k;
        ++n;
      }
    }
    if (n == 0) {   // fallback detection, cad:225 (a constant: no gradient)
      box[(r * MAXDET) * 4 + 0] = 180.f; box[(r * MAXDET) * 4 + 1] = 120.f;
      box[(r * MAXDET) * 4 + 2] = 30.f;  box[(r * MAXDET) * 4 + 3] = 60.f;
      src[r * MAXDET] = -1;
      n = 1;
    } else if (flag) {
      *flag = 1.f;   // the detector group receives a gradient this step
    }
    for (int k = n; k < MAXDET; ++k) {
      for (int c = 0; c < 4; ++c) box[(r * MAXDET + k) * 4 + c] = 0.f;
      src[r * MAXDET + k] = -1;
    }
    cnt[r] = n;
  }
}

__global__ void det_decode_bwd_kernel(const float* __restrict__ raw, const float* __restrict__ dbox, const int* __restrict__ src,
                                      long long R, float* __restrict__ draw) {
  cvad_pdl_enter();
  const float scale[4] = {360.f, 240.f, 80.f, 120.f};
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < R; r += (long long)gridDim.x * blockDim.x) {
    for (int k = 0; k < MAXDET * 4; ++k) draw[r * MAXDET * 4 + k] = 0.f;
    for (int k = 0; k < MAXDET; ++k) {
      int j = src[r * MAXDET + k];
      if (j < 0) continue;
      for (int c = 0; c < 4; ++c) {
        float s = sigmoidf_(raw[(r * MAXDET + j) * 4 + c]);
        draw[(r * MAXDET + j) * 4 + c] = dbox[(r * MAXDET + k) * 4 + c] * scale[c] * s * (1.f - s);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ trajectories
// one block per clip; box (B,T,5,4), reid (B,T,5,64) -> traj (B,5,T,68); ntr[b] = max_t cnt[b,t]
__global__ void traj_assemble_kernel(const float* __restrict__ box, const float* __restrict__ reid, const int* __restrict__ cnt, int T,
                                     int reid_dim, float* __restrict__ traj, int* __restrict__ ntr, float* __restrict__ multi_flag) {
  cvad_pdl_enter();
  const int b = blockIdx.x;
  const int F = 4 + reid_dim;
  __shared__ int smax;
  if (threadIdx.x == 0) {
    int m = 0;
    for (int t = 0; t < T; ++t) m = max(m, cnt[b * T + t]);
    smax = m;
    ntr[b] = m;
    if (m >= 2 && multi_flag) *multi_flag = 1.f;   // >= 2 tracks: the edge MLP runs and receives gradients (cad:382-387)
  }
  __syncthreads();
  const int total = MAXDET * T * F;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    int f = i % F, t = (i / F) % T, k = i / (F * T);
    long long r = (long long)b * T + t;
    float v = 0.f;
    if (k < cnt[r]) v = f < 4 ? box[(r * MAXDET + k) * 4 + f] : reid[(r * MAXDET + k) * reid_dim + (f - 4)];
    traj[(((long long)b * MAXDET + k) * T + t) * F + f] = v;
  }
}

__global__ void traj_assemble_bwd_kernel(const float* __restrict__ dtraj, const int* __restrict__ cnt, int T, int reid_dim,
                                         float* __restrict__ dbox, float* __restrict__ dreid) {
  cvad_pdl_enter();
  const int b = blockIdx.x;
  const int F = 4 + reid_dim;
  const int total = MAXDET * T * F;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    int f = i % F, t = (i / F) % T, k = i / (F * T);
    long long r = (long long)b * T + t;
    float v = k < cnt[r] ? dtraj[(((long long)b * MAXDET + k) * T + t) * F + f] : 0.f;
    if (f < 4) dbox[(r * MAXDET + k) * 4 + f] = v;
    else dreid[(r * MAXDET + k) * reid_dim + (f - 4)] = v;
  }
}

// ------------------------------------------------------------------------------------------------ GRU recurrence
// gi (N,T,192) = x W_ih^T + b_ih precomputed.  One block (192 threads) per sequence n; inactive tracks emit zeros.
// saved (N,T,5,64): r, z, n, gh_n, h_prev
__global__ void __launch_bounds__(192) gru_fwd_kernel(const float* __restrict__ gi, const float* __restrict__ w_hh,
                                                      const float* __restrict__ b_hh, const int* __restrict__ ntr, int T,
                                                      float* __restrict__ hT, float* __restrict__ saved) {
  cvad_pdl_enter();
  const int n = blockIdx.x, g = threadIdx.x;
  const int b = n / MAXDET, k = n % MAXDET;
  __shared__ float h[HID];
  __shared__ float gh[3 * HID];
  if (k >= ntr[b]) {
    if (g < HID) hT[(long long)n * HID + g] = 0.f;
    return;
  }
  float w[HID];
#pragma unroll
  for (int i = 0; i < HID; ++i) w[i] = w_hh[g * HID + i];
  const float bias = b_hh[g];
  if (g < HID) h[g] = 0.f;
  __syncthreads();
  for (int t = 0; t < T; ++t) {
    float acc = bias;
#pragma unroll
    for (int i = 0; i < HID; ++i) acc = fmaf(w[i], h[i], acc);
    gh[g] = acc;
    __syncthreads();
    if (g < HID) {
      const float* gir = gi + ((long long)n * T + t) * 3 * HID;
      float r = sigmoidf_(gir[g] + gh[g]);
      float z = sigmoidf_(gir[HID + g] + gh[HID + g]);
      float ghn = gh[2 * HID + g];
      float nn = tanhf(gir[2 * HID + g] + r * ghn);
      float hp = h[g];
      float hn = (1.f - z) * nn + z * hp;
      if (saved) {
        float* s = saved + ((long long)n * T + t) * 5 * HID;
        s[g] = r; s[HID + g] = z; s[2 * HID + g] = nn; s[3 * HID + g] = ghn; s[4 * HID + g] = hp;
      }
      h[g] = hn;
    }
    __syncthreads();
  }
  if (g < HID) hT[(long long)n * HID + g] = h[g];
}

__global__ void __launch_bounds__(192) gru_bwd_kernel(const float* __restrict__ dhT, const float* __restrict__ saved,
                                                      const float* __restrict__ w_hh, const int* __restrict__ ntr, int T,
                                                      float* __restrict__ dgi, float* __restrict__ dw_hh, float* __restrict__ db_hh) {
  cvad_pdl_enter();
  const int n = blockIdx.x, g = threadIdx.x;
  const int b = n / MAXDET, k = n % MAXDET;
  extern __shared__ float sm[];
  float* wT = sm;                      // [HID][3*HID+1]  transposed W_hh: wT[i][g] = w_hh[g][i]
  float* dh = wT + HID * (3 * HID + 1);
  float* dgh = dh + HID;
  float* hp = dgh + 3 * HID;
  if (k >= ntr[b]) {
    for (int t = 0; t < T; ++t) dgi[((long long)n * T + t) * 3 * HID + g] = 0.f;
    return;
  }
  for (int i = 0; i < HID; ++i) wT[i * (3 * HID + 1) + g] = w_hh[g * HID + i];
  if (g < HID) dh[g] = dhT[(long long)n * HID + g];
  float dw[HID];
#pragma unroll
  for (int i = 0; i < HID; ++i) dw[i] = 0.f;
  float dbias = 0.f;
  __syncthreads();
  for (int t = T - 1; t >= 0; --t) {
    const float* s = saved + ((long long)n * T + t) * 5 * HID;
    float* dgir = dgi + ((long long)n * T + t) * 3 * HID;
    if (g < HID) {
      float r = s[g], z = s[HID + g], nn = s[2 * HID + g], ghn = s[3 * HID + g], hprev = s[4 * HID + g];
      float d = dh[g];
      float dn = d * (1.f - z);
      float dz = d * (hprev - nn);
      float dnpre = dn * (1.f - nn * nn);
      float dr = dnpre * ghn;
      float dzpre = dz * z * (1.f - z);
      float drpre = dr * r * (1.f - r);
      dgir[g] = drpre; dgir[HID + g] = dzpre; dgir[2 * HID + g] = dnpre;
      dgh[g] = drpre; dgh[HID + g] = dzpre; dgh[2 * HID + g] = dnpre * r;
      hp[g] = hprev;
      dh[g] = d * z;      // direct path; the W_hh^T dgh term is added below
    }
    __syncthreads();
    {
      float my = dgh[g];
      dbias += my;
#pragma unroll
      for (int i = 0; i < HID; ++i) dw[i] = fmaf(my, hp[i], dw[i]);
    }
    if (g < HID) {
      float acc = 0.f;
      const float* col = wT + g * (3 * HID + 1);
      for (int j = 0; j < 3 * HID; ++j) acc = fmaf(col[j], dgh[j], acc);
      dh[g] += acc;
    }
    __syncthreads();
  }
  if (dw_hh)
    for (int i = 0; i < HID; ++i) atomicAdd(dw_hh + g * HID + i, dw[i]);
  if (db_hh) atomicAdd(db_hh + g, dbias);
}

// ------------------------------------------------------------------------------------------------ VAE head
// one block per clip, threads over (track, factor)
__global__ void reparam_kl_kernel(const float* __restrict__ mu, const float* __restrict__ lv, const float* __restrict__ eps,
                                  const int* __restrict__ ntr, float* __restrict__ z, float* __restrict__ kl) {
  cvad_pdl_enter();
  const int b = blockIdx.x, i = threadIdx.x;   // 32 threads, 30 used
  const int k = i / NF;
  const int nt = ntr[b];
  float term = 0.f;
  if (i < MAXDET * NF) {
    long long o = (long long)b * MAXDET * NF + i;
    if (k < nt) {
      float m = mu[o], l = lv[o];
      z[o] = m + eps[o] * expf(0.5f * l);
      term = -0.5f * (1.f + l - m * m - expf(l));
    } else {
      z[o] = 0.f;
    }
  }
  term = warp_sum(term);
  if (i == 0) kl[b] = term / (float)nt;
}

__global__ void reparam_kl_bwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv, const float* __restrict__ eps,
                                      const int* __restrict__ ntr, const float* __restrict__ dz, const float* __restrict__ dkl,
                                      float* __restrict__ dmu, float* __restrict__ dlv) {
  cvad_pdl_enter();
  const int b = blockIdx.x, i = threadIdx.x;
  if (i >= MAXDET * NF) return;
  const int k = i / NF;
  const int nt = ntr[b];
  long long o = (long long)b * MAXDET * NF + i;
  if (k < nt) {
    float m = mu[o], l = lv[o];
    float g = dz[o];
    float gk = dkl[b] / (float)nt;
    dmu[o] = g + gk * m;
    dlv[o] = g * eps[o] * 0.5f * expf(0.5f * l) + gk * (-0.5f) * (1.f - expf(l));
  } else {
    dmu[o] = 0.f;
    dlv[o] = 0.f;
  }
}

// ------------------------------------------------------------------------------------------------ structure learner
__global__ void pair_concat_kernel(const float* __restrict__ node, int B, int Hn, float* __restrict__ pair) {
  cvad_pdl_enter();
  long long total = (long long)B * MAXDET * MAXDET * 2 * Hn;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    int f = (int)(t % (2 * Hn));
    long long q = t / (2 * Hn);
    int j = (int)(q % MAXDET), i = (int)((q / MAXDET) % MAXDET);
    long long b = q / (MAXDET * MAXDET);
    pair[t] = f < Hn ? node[(b * MAXDET + i) * Hn + f] : node[(b * MAXDET + j) * Hn + (f - Hn)];
  }
}
__global__ void pair_concat_bwd_kernel(const float* __restrict__ dpair, int B, int Hn, float* __restrict__ dnode) {
  cvad_pdl_enter();
  long long total = (long long)B * MAXDET * Hn;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    int f = (int)(t % Hn);
    int i = (int)((t / Hn) % MAXDET);
    long long b = t / ((long long)Hn * MAXDET);
    float s = 0.f;
    for (int j = 0; j < MAXDET; ++j) {
      s += dpair[((b * MAXDET + i) * MAXDET + j) * 2 * Hn + f];            // node_i as the first half of pair (i,j)
      s += dpair[((b * MAXDET + j) * MAXDET + i) * 2 * Hn + Hn + f];       // node_i as the second half of pair (j,i)
    }
    dnode[t] = s;
  }
}

// e (B,5,5) -> adj (B,6,6) (bwd: dadj -> de), masked by i != j and i,j < ntr[b]
__global__ void adj_assemble_kernel(const float* __restrict__ src, const int* __restrict__ ntr, int B, float* __restrict__ dst,
                                    int backward) {
  cvad_pdl_enter();
  long long total = (long long)B * NF * NF;
  if (!backward) {
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
      int j = (int)(t % NF), i = (int)((t / NF) % NF);
      long long b = t / (NF * NF);
      int nt = ntr[b];
      float v = 0.f;
      if (i < MAXDET && j < MAXDET && i != j && i < nt && j < nt) v = src[(b * MAXDET + i) * MAXDET + j];
      dst[t] = v;
    }
  } else {
    long long tot5 = (long long)B * MAXDET * MAXDET;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < tot5; t += (long long)gridDim.x * blockDim.x) {
      int j = (int)(t % MAXDET), i = (int)((t / MAXDET) % MAXDET);
      long long b = t / (MAXDET * MAXDET);
      int nt = ntr[b];
      dst[t] = (i != j && i < nt && j < nt) ? src[(b * NF + i) * NF + j] : 0.f;
    }
  }
}

// structured[b,k,i] = sum_j adj[b,i,j] z[b,k,j]
__global__ void structured_kernel(const float* __restrict__ adj, const float* __restrict__ z, int B, float* __restrict__ out) {
  cvad_pdl_enter();
  long long total = (long long)B * MAXDET * NF;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    int i = (int)(t % NF), k = (int)((t / NF) % MAXDET);
    long long b = t / (NF * MAXDET);
    float s = 0.f;
    for (int j = 0; j < NF; ++j) s = fmaf(adj[(b * NF + i) * NF + j], z[(b * MAXDET + k) * NF + j], s);
    out[t] = s;
  }
}
__global__ void structured_bwd_kernel(const float* __restrict__ adj, const float* __restrict__ z, const float* __restrict__ dout, int B,
                                      float* __restrict__ dadj, float* __restrict__ dz) {
  cvad_pdl_enter();
  long long na = (long long)B * NF * NF, nz = (long long)B * MAXDET * NF;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < na + nz; t += (long long)gridDim.x * blockDim.x) {
    if (t < na) {
      int j = (int)(t % NF), i = (int)((t / NF) % NF);
      long long b = t / (NF * NF);
      float s = 0.f;
      for (int k = 0; k < MAXDET; ++k) s = fmaf(dout[(b * MAXDET + k) * NF + i], z[(b * MAXDET + k) * NF + j], s);
      dadj[t] = s;
    } else {
      long long u = t - na;
      int j = (int)(u % NF), k = (int)((u / NF) % MAXDET);
      long long b = u / (NF * MAXDET);
      float s = 0.f;
      for (int i = 0; i < NF; ++i) s = fmaf(dout[(b * MAXDET + k) * NF + i], adj[(b * NF + i) * NF + j], s);
      dz[u] = s;
    }
  }
}

// ------------------------------------------------------------------------------------------------ scorer inputs
// cin (B,18) = [cur | prd | |cur-prd|], min (B,12) = [cur | prd], tin (B,6) = cur   (means over the clip's tracks)
__global__ void scorer_inputs_kernel(const float* __restrict__ z, const float* __restrict__ pred, const int* __restrict__ ntr, int B,
                                     float* __restrict__ cin, float* __restrict__ min_, float* __restrict__ tin) {
  cvad_pdl_enter();
  long long total = (long long)B * NF;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    int f = (int)(t % NF);
    long long b = t / NF;
    int nt = ntr[b];
    float c = 0.f, p = 0.f;
    for (int k = 0; k < nt; ++k) {
      c += z[(b * MAXDET + k) * NF + f];
      p += pred[(b * MAXDET + k) * NF + f];
    }
    c /= (float)nt;
    p /= (float)nt;
    cin[b * 18 + f] = c; cin[b * 18 + 6 + f] = p; cin[b * 18 + 12 + f] = fabsf(c - p);
    min_[b * 12 + f] = c; min_[b * 12 + 6 + f] = p;
    tin[b * 6 + f] = c;
  }
}
__global__ void scorer_inputs_bwd_kernel(const float* __restrict__ cin, const int* __restrict__ ntr, int B, const float* __restrict__ dcin,
                                         const float* __restrict__ dmin, const float* __restrict__ dtin, float* __restrict__ dz,
                                         float* __restrict__ dpred) {
  cvad_pdl_enter();
  long long total = (long long)B * NF;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    int f = (int)(t % NF);
    long long b = t / NF;
    int nt = ntr[b];
    float d = cin[b * 18 + f] - cin[b * 18 + 6 + f];
    float sg = (d > 0.f) - (d < 0.f);
    float gd = dcin[b * 18 + 12 + f] * sg;
    float gc = dcin[b * 18 + f] + dmin[b * 12 + f] + dtin[b * 6 + f] + gd;
    float gp = dcin[b * 18 + 6 + f] + dmin[b * 12 + 6 + f] - gd;
    for (int k = 0; k < MAXDET; ++k) {
      dz[(b * MAXDET + k) * NF + f] = k < nt ? gc / (float)nt : 0.f;
      dpred[(b * MAXDET + k) * NF + f] = k < nt ? gp / (float)nt : 0.f;
    }
  }
}

__global__ void lincomb3_kernel(float* __restrict__ out, const float* __restrict__ x, float a, const float* __restrict__ y, float b,
                                const float* __restrict__ z, float c, long long n) {
  cvad_pdl_enter();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = a * x[i] + b * y[i] + c * z[i];
}

// ------------------------------------------------------------------------------------------------ softmax over small rows
__global__ void softmax_rows_kernel(const float* __restrict__ x, long long rows, int C, float* __restrict__ y) {
  cvad_pdl_enter();
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
    float m = -INFINITY;
    for (int c = 0; c < C; ++c) m = fmaxf(m, x[r * C + c]);
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += expf(x[r * C + c] - m);
    for (int c = 0; c < C; ++c) y[r * C + c] = expf(x[r * C + c] - m) / s;
  }
}
__global__ void softmax_rows_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy, long long rows, int C,
                                        float* __restrict__ dx) {
  cvad_pdl_enter();
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
    float dot = 0.f;
    for (int c = 0; c < C; ++c) dot += y[r * C + c] * dy[r * C + c];
    for (int c = 0; c < C; ++c) dx[r * C + c] = y[r * C + c] * (dy[r * C + c] - dot);
  }
}

inline int nb(long long n) {
  long long b = (n + 127) / 128;
  long long cap = 8LL * cvad_num_sms();
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

CVAD_API int cvad_det_decode_f32(const float* raw, long long rows, float* box, int* cnt, int* src, float* active_flag, void* stream) {
  if (rows <= 0) return 0;
  cvad_launch_pdl(det_decode_kernel, dim3(nb(rows)), dim3(128), 0, (cudaStream_t)stream, raw, rows, box, cnt, src, active_flag);
  CVAD_LAUNCH_CHECK();
  return 0;
}
CVAD_API int cvad_det_decode_bwd_f32(const float* raw, const float* dbox, const int* src, long long rows, float* draw, void* stream) {
  if (rows <= 0) return 0;
  cvad_launch_pdl(det_decode_bwd_kernel, dim3(nb(rows)), dim3(128), 0, (cudaStream_t)stream, raw, dbox, src, rows, draw);
  CVAD_LAUNCH_CHECK();
  return 0;
}
CVAD_API int cvad_traj_assemble_f32(const float* box, const float* reid, const int* cnt, int B, int T, int reid_dim, float* traj, int* ntr,
                                    float* multi_flag, void* stream) {
  if (B <= 0) return 0;
  cvad_launch_pdl(traj_assemble_kernel, dim3(B), dim3(256), 0, (cudaStream_t)stream, box, reid, cnt, T, reid_dim, traj, ntr, multi_flag);
  CVAD_LAUNCH_CHECK();
  return 0;
}
CVAD_API int cvad_traj_assemble_bwd_f32(const float* dtraj, const int* cnt, int B, int T, int reid_dim, float* dbox, float* dreid,
                                        void* stream) {
  if (B <= 0) return 0;
  cvad_launch_pdl(traj_assemble_bwd_kernel, dim3(B), dim3(256), 0, (cudaStream_t)stream, dtraj, cnt, T, reid_dim, dbox, dreid);
  CVAD_LAUNCH_CHECK();
  return 0;
}
CVAD_API int cvad_gru_fwd_f32(const float* gi, const float* w_hh, const float* b_hh, const int* ntr, int B, int T, float* hT, float* saved,
                              void* stream) {
  if (B <= 0) return 0;
  cvad_launch_pdl(gru_fwd_kernel, dim3(B * MAXDET), dim3(192), 0, (cudaStream_t)stream, gi, w_hh, b_hh, ntr, T, hT, saved);
  CVAD_LAUNCH_CHECK();
  return 0;
}
CVAD_API int cvad_gru_bwd_f32(const float* dhT, const float* saved, const float* w_hh, const int* ntr, int B, int T, float* dgi,
                              float* dw_hh, float* db_hh, void* stream) {
  if (B <= 0) return 0;
  size_t smem = (size_t)(HID * (3 * HID + 1) + HID + 3 * HID + HID) * sizeof(float);
  static size_t configured[CVAD_MAX_DEVICES] = {};
  const cudaError_t ce = cvad_ensure_dyn_smem(gru_bwd_kernel, smem, configured);
  if (ce != cudaSuccess) return (int)ce;
  cvad_launch_pdl(gru_bwd_kernel, dim3(B * MAXDET), dim3(192), smem, (cudaStream_t)stream, dhT, saved, w_hh, ntr, T, dgi, dw_hh, db_hh);
  CVAD_LAUNCH_CHECK();
  return 0;
}
CVAD_API int cvad_reparam_kl_f32(const float* mu, const float* logvar, const float* eps, const int* ntr, int B, float* z, float* kl,
                                 void* stream) {
  if (B <= 0) return 0;
  cvad_launch_pdl(reparam_kl_kernel, dim3(B), dim3(32), 0, (cudaStream_t)stream, mu, logvar, eps, ntr, z, kl);
  CVAD_LAUNCH_CHECK();
  return 0;
}
CVAD_API int cvad_reparam_kl_bwd_f32(const float* mu, const float* logvar, const float* eps, const int* ntr, int B, const float* dz,
                                     const float* dkl, float* dmu, float* dlogvar, void* stream) {
  if (B <= 0) return 0;
  cvad_launch_pdl(reparam_kl_bwd_kernel, dim3(B), dim3(32), 0, (cudaStream_t)stream, mu, logvar, eps, ntr, dz, dkl, dmu, dlogvar);
  CVAD_LAUNCH_CHECK();
  return 0;
}
CVAD_API int cvad_pair_concat_f32(const float* node, int B, int node_dim, float* pair, void* stream) {
  if (B <= 0) return 0;
  cvad_launch_pdl(pair_concat_kernel, dim3(nb((long long)B * 25 * 2 * node_dim)), dim3(128), 0, (cudaStream_t)stream, node, B, node_dim, pair);
  CVAD_LAUNCH_CHECK();
  return 0;
}
CVAD_API int cvad_pair_concat_bwd_f32(const float* dpair, int B, int node_dim, float* dnode, void* stream) {
  if (B <= 0) return 0;
  cvad_launch_pdl(pair_concat_bwd_kernel, dim3(nb((long long)B * 5 * node_dim)), dim3(128), 0, (cudaStream_t)stream, dpair, B, node_dim, dnode);
  CVAD_LAUNCH_CHECK();
  return 0;
}
CVAD_API int cvad_adj_assemble_f32(const float* src, const int* ntr, int B, float* dst, int backward, void* stream) {
  if (B <= 0) return 0;
  cvad_launch_pdl(adj_assemble_kernel, dim3(nb((long long)B * 36)), dim3(128), 0, (cudaStream_t)stream, src, ntr, B, dst, backward);
  CVAD_LAUNCH_CHECK();
  return 0;
}
CVAD_API int cvad_structured_f32(const float* adj, const float* z, int B, float* out, void* stream) {
  if (B <= 0) return 0;
  cvad_launch_pdl(structured_kernel, dim3(nb((long long)B * 30)), dim3(128), 0, (cudaStream_t)stream, adj, z, B, out);
  CVAD_LAUNCH_CHECK();
  return 0;
}
CVAD_API int cvad_structured_bwd_f32(const float* adj, const float* z, const float* dout, int B, float* dadj, float* dz, void* stream) {
  if (B <= 0) return 0;
  cvad_launch_pdl(structured_bwd_kernel, dim3(nb((long long)B * 66)), dim3(128), 0, (cudaStream_t)stream, adj, z, dout, B, dadj, dz);
  CVAD_LAUNCH_CHECK();
  return 0;
}
CVAD_API int cvad_scorer_inputs_f32(const float* z, const float* pred, const int* ntr, int B, float* cin, float* min_, float* tin,
                                    void* stream) {
  if (B <= 0) return 0;
  cvad_launch_pdl(scorer_inputs_kernel, dim3(nb((long long)B * 6)), dim3(128), 0, (cudaStream_t)stream, z, pred, ntr, B, cin, min_, tin);
  CVAD_LAUNCH_CHECK();
  return 0;
}
CVAD_API int cvad_scorer_inputs_bwd_f32(const float* cin, const int* ntr, int B, const float* dcin, const float* dmin, const float* dtin,
                                        float* dz, float* dpred, void* stream) {
  if (B <= 0) return 0;
  cvad_launch_pdl(scorer_inputs_bwd_kernel, dim3(nb((long long)B * 6)), dim3(128), 0, (cudaStream_t)stream, cin, ntr, B, dcin, dmin, dtin, dz, dpred);
  CVAD_LAUNCH_CHECK();
  return 0;
}
CVAD_API int cvad_lincomb3_f32(float* out, const float* x, float a, const float* y, float b, const float* z, float c, long long n,
                               void* stream) {
  if (n <= 0) return 0;
  cvad_launch_pdl(lincomb3_kernel, dim3(nb(n)), dim3(128), 0, (cudaStream_t)stream, out, x, a, y, b, z, c, n);
  CVAD_LAUNCH_CHECK();
  return 0;
}
CVAD_API int cvad_softmax_rows_f32(const float* x, long long rows, int C, float* y, void* stream) {
  if (rows <= 0) return 0;
  cvad_launch_pdl(softmax_rows_kernel, dim3(nb(rows)), dim3(128), 0, (cudaStream_t)stream, x, rows, C, y);
  CVAD_LAUNCH_CHECK();
  return 0;
}
CVAD_API int cvad_softmax_rows_bwd_f32(const float* y, const float* dy, long long rows, int C, float* dx, void* stream) {
  if (rows <= 0) return 0;
  cvad_launch_pdl(softmax_rows_bwd_kernel, dim3(nb(rows)), dim3(128), 0, (cudaStream_t)stream, y, dy, rows, C, dx);
  CVAD_LAUNCH_CHECK();
  return 0;
}
