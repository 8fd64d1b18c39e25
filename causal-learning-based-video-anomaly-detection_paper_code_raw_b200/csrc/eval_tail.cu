// Evaluation tail on the device: what the reference does on the host with numpy / sklearn right after the hot path --
//   * np.percentile(scores, q) thresholds and the pseudo-labels above them        s1:60, cad1:609, cad1:709
//   * sklearn.metrics.roc_auc_score (rank statistic with tie handling)            mc3:388, cad:1233-1248
//   * the 8 evaluation metrics incl. len(np.unique(graphs, axis=0))               s2:286-295
//   * np.convolve(scores, ones(w)/w, 'valid') moving average                      cad:1085-1087, vad:833-835
// so that a whole evaluation pass reads back a handful of scalars instead of every score and every (16,16) graph.
//
// One building block: an in-place bitonic sort of 16-byte {key, value} records in global memory (shared-memory passes for strides
// below 1024, one launch per larger stride).  Clip counts are 10^2..10^5, so this is latency-bound work: a few launches of a few
// microseconds.  Float keys are mapped to order-preserving unsigned integers (NaNs sort last, as numpy does).
// No allocation inside: the caller passes the record workspace (cvad_eval_workspace_bytes).
#include "common.cuh"
#include "cvad_b200.h"

namespace {

struct KV {
  unsigned long long k;
  unsigned int v, pad;
};

__device__ __forceinline__ bool kv_less(const KV& a, const KV& b) { return a.k < b.k || (a.k == b.k && a.v < b.v); }

__device__ __forceinline__ unsigned int float_order_key(float x) {
  if (x != x) return 0xFFFFFFFFu;                       // NaN: after +inf
  unsigned int u = __float_as_uint(x + 0.0f);           // -0.0 + 0.0 = +0.0: the two zeros compare equal, like numpy
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

constexpr int SORT_TILE = 2048;                          // records per CTA in the shared-memory passes (32 KB)
constexpr int SORT_THREADS = 1024;

__device__ __forceinline__ void cmp_swap(KV& a, KV& b, bool ascending) {
  if (kv_less(b, a) == ascending) { KV t = a; a = b; b = t; }
}

// all stages k = 2 .. SORT_TILE of the network on one tile held in shared memory
__global__ void __launch_bounds__(SORT_THREADS) bitonic_tile_sort_kernel(KV* __restrict__ d) {
  __shared__ KV s[SORT_TILE];
  const long long base = (long long)blockIdx.x * SORT_TILE;
  for (int i = threadIdx.x; i < SORT_TILE; i += SORT_THREADS) s[i] = d[base + i];
  __syncthreads();
  for (int k = 2; k <= SORT_TILE; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      const int t = threadIdx.x;
      const int i = ((t / j) * 2 * j) + (t % j);        // lower index of the pair
      const bool asc = (((base + i) & k) == 0);
      cmp_swap(s[i], s[i + j], asc);
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < SORT_TILE; i += SORT_THREADS) d[base + i] = s[i];
}

// one (k, j) step with j >= SORT_TILE over global memory
__global__ void bitonic_global_step_kernel(KV* __restrict__ d, long long n2, long long k, long long j) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n2 / 2; t += (long long)gridDim.x * blockDim.x) {
    const long long i = ((t / j) * 2 * j) + (t % j);
    KV a = d[i], b = d[i + j];
    const bool asc = ((i & k) == 0);
    if (kv_less(b, a) == asc) { d[i] = b; d[i + j] = a; }
  }
}

// the steps j = SORT_TILE/2 .. 1 of stage k (k > SORT_TILE) on one tile in shared memory
__global__ void __launch_bounds__(SORT_THREADS) bitonic_tile_merge_kernel(KV* __restrict__ d, long long k) {
  __shared__ KV s[SORT_TILE];
  const long long base = (long long)blockIdx.x * SORT_TILE;
  for (int i = threadIdx.x; i < SORT_TILE; i += SORT_THREADS) s[i] = d[base + i];
  __syncthreads();
  const bool asc = ((base & k) == 0);                    // constant over the tile because k > SORT_TILE
  for (int j = SORT_TILE >> 1; j > 0; j >>= 1) {
    const int t = threadIdx.x;
    const int i = ((t / j) * 2 * j) + (t % j);
    cmp_swap(s[i], s[i + j], asc);
    __syncthreads();
  }
  for (int i = threadIdx.x; i < SORT_TILE; i += SORT_THREADS) d[base + i] = s[i];
}

long long pow2_at_least(long long n) {
  long long p = SORT_TILE;
  while (p < n) p <<= 1;
  return p;
}

int sort_records(KV* d, long long n2, cudaStream_t st) {
  bitonic_tile_sort_kernel<<<(unsigned)(n2 / SORT_TILE), SORT_THREADS, 0, st>>>(d);
  for (long long k = 2LL * SORT_TILE; k <= n2; k <<= 1) {
    for (long long j = k >> 1; j >= SORT_TILE; j >>= 1) {
      long long blocks = (n2 / 2 + 255) / 256;
      if (blocks > 8LL * cvad_num_sms()) blocks = 8LL * cvad_num_sms();
      bitonic_global_step_kernel<<<(unsigned)blocks, 256, 0, st>>>(d, n2, k, j);
    }
    bitonic_tile_merge_kernel<<<(unsigned)(n2 / SORT_TILE), SORT_THREADS, 0, st>>>(d, k);
  }
  CVAD_LAUNCH_CHECK();
  return 0;
}

__global__ void fill_score_records_kernel(const float* __restrict__ x, long long n, long long n2, KV* __restrict__ d) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
    KV r;
    r.k = i < n ? (unsigned long long)float_order_key(x[i]) : ~0ULL;      // padding sorts behind every real record (NaN key is 2^32-1)
    r.v = (unsigned int)i;
    r.pad = 0;
    d[i] = r;
  }
}

__global__ void unpack_sorted_kernel(const KV* __restrict__ d, const float* __restrict__ x, long long n, float* __restrict__ sorted,
                                     int* __restrict__ order) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned int src = d[i].v;
    if (sorted) sorted[i] = x[src];
    if (order) order[i] = (int)src;
  }
}

// np.percentile(a, q) for float32 a, default 'linear' method -- numpy does ALL of it in float32 for a float32 array (the quantile
// q / 100, the virtual index (n - 1) * quantile, gamma and the two-sided lerp of numpy/lib/_function_base_impl.py:_lerp); mirrored
// operation by operation so the threshold is bit-identical.
__global__ void percentile_kernel(const float* __restrict__ sorted, long long n, float q, float* __restrict__ out) {
  if (threadIdx.x || blockIdx.x) return;
  if (n <= 0) { out[0] = 0.f; return; }
  const float quant = __fdiv_rn(q, 100.f);
  const float vi = __fmul_rn((float)(n - 1), quant);
  long long lo = (long long)floorf(vi);
  if (lo < 0) lo = 0;
  if (lo > n - 1) lo = n - 1;
  const long long hi = lo + 1 < n ? lo + 1 : n - 1;
  const float g = __fsub_rn(vi, (float)lo);
  const float a = sorted[lo], b = sorted[hi];
  const float d = __fsub_rn(b, a);
  out[0] = g >= 0.5f ? __fsub_rn(b, __fmul_rn(d, __fsub_rn(1.f, g))) : __fadd_rn(a, __fmul_rn(d, g));
}

__global__ void threshold_labels_kernel(const float* __restrict__ x, long long n, const float* __restrict__ thr, float* __restrict__ labels) {
  const float t = thr[0];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    labels[i] = x[i] > t ? 1.f : 0.f;
}

// Rank statistic of the positives with average ranks for ties (== the area under sklearn's ROC curve): for every sorted position the
// bounds of its group of equal keys come from two binary searches; acc[0] += sum of ranks of positives, acc[1] += positives.
__global__ void auc_rank_kernel(const KV* __restrict__ d, const float* __restrict__ targets, long long n, double* __restrict__ acc) {
  __shared__ double sh[32];
  double rs = 0.0, np_ = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const KV r = d[i];
    if (targets[r.v] > 0.5f) {
      long long lo = 0, hi = i;                          // first position with the same key
      while (lo < hi) { const long long m = (lo + hi) >> 1; if (d[m].k < r.k) lo = m + 1; else hi = m; }
      const long long first = lo;
      lo = i; hi = n;                                    // one past the last position with the same key
      while (lo < hi) { const long long m = (lo + hi) >> 1; if (d[m].k <= r.k) lo = m + 1; else hi = m; }
      rs += 0.5 * (double)(first + (lo - 1)) + 1.0;
      np_ += 1.0;
    }
  }
  rs = block_sum_d(rs, sh);
  np_ = block_sum_d(np_, sh);
  if (threadIdx.x == 0 && np_ > 0.0) { atomicAdd(acc, rs); atomicAdd(acc + 1, np_); }
}

__global__ void auc_finalize_kernel(const double* __restrict__ acc, long long n, double* __restrict__ auc) {
  const double npos = acc[1], nneg = (double)n - acc[1];
  auc[0] = (npos == 0.0 || nneg == 0.0) ? 0.0 : (acc[0] - npos * (npos + 1.0) * 0.5) / (npos * nneg);
}

__global__ void init_acc_kernel(double* __restrict__ acc, unsigned int* __restrict__ mm) {
  if (threadIdx.x < 16) acc[threadIdx.x] = 0.0;
  if (threadIdx.x == 0) { mm[0] = 0xFFFFFFFFu; mm[1] = 0u; }
}

// ---- s2:286-295: mean / std / min / max of the scores, mean edge count (adj > 0.1), number of distinct graphs
__device__ __forceinline__ unsigned long long mix64(unsigned long long h, unsigned long long v) {
  h ^= v + 0x9E3779B97F4A7C15ULL + (h << 6) + (h >> 2);
  h *= 0xFF51AFD7ED558CCDULL;
  return h ^ (h >> 33);
}

// one warp per graph row: edge count and a 64-bit hash of the row's float bits (-0.0 canonicalised)
__global__ void graph_rows_kernel(const float* __restrict__ g, long long n, int row_len, float edge_thr, long long n2, KV* __restrict__ d,
                                  double* __restrict__ acc /* [5] += edges */) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  double edges = 0.0;
  for (long long r = warp; r < n2; r += nwarps) {
    if (r >= n) {
      if (lane == 0) { KV p; p.k = ~0ULL; p.v = (unsigned int)r; p.pad = 0; d[r] = p; }
      continue;
    }
    const float* row = g + r * row_len;
    unsigned long long h = 0x243F6A8885A308D3ULL + lane;
    int e = 0;
    for (int c = lane; c < row_len; c += 32) {
      const float v = row[c];
      e += v > edge_thr;
      h = mix64(h, (unsigned long long)__float_as_uint(v + 0.0f) | ((unsigned long long)c << 32));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      e += __shfl_xor_sync(0xffffffffu, e, o);
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, h, o);
      h = mix64(h, other) + mix64(other, h);            // symmetric: both lanes of the pair end up with the same value
    }
    if (lane == 0) {
      KV p;
      p.k = h == ~0ULL ? h - 1 : h;
      p.v = (unsigned int)r;
      p.pad = 0;
      d[r] = p;
      edges += (double)e;
    }
  }
  if (lane == 0 && edges != 0.0) atomicAdd(acc + 5, edges);
}

// after the sort: a record opens a new distinct graph when its hash differs from its predecessor's or -- same hash -- the rows differ
__global__ void count_distinct_kernel(const KV* __restrict__ d, const float* __restrict__ g, long long n, int row_len, double* __restrict__ acc) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  double cnt = 0.0;
  for (long long i = warp; i < n; i += nwarps) {
    bool fresh = i == 0 || d[i].k != d[i - 1].k;
    if (!fresh) {
      const float* a = g + (long long)d[i].v * row_len;
      const float* b = g + (long long)d[i - 1].v * row_len;
      int diff = 0;
      for (int c = lane; c < row_len; c += 32) diff |= !(a[c] == b[c]) && !(a[c] != a[c] && b[c] != b[c]);
      fresh = __any_sync(0xffffffffu, diff);
    }
    cnt += fresh ? 1.0 : 0.0;
  }
  if (lane == 0 && cnt != 0.0) atomicAdd(acc + 6, cnt);
}

// acc[0] sum, acc[1] sum of squares (about the first score, for a stable variance), acc[2] min key, acc[3] max key (order-preserving ints)
__global__ void score_moments_kernel(const float* __restrict__ x, long long n, double* __restrict__ acc, unsigned int* __restrict__ mm) {
  __shared__ double sh[32];
  const double pivot = n > 0 ? (double)x[0] : 0.0;
  double s = 0.0, q = 0.0;
  unsigned int lo = 0xFFFFFFFFu, hi = 0u;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    const double dv = (double)v - pivot;
    s += dv;
    q += dv * dv;
    const unsigned int k = float_order_key(v);
    lo = k < lo ? k : lo;
    hi = k > hi ? k : hi;
  }
  s = block_sum_d(s, sh);
  q = block_sum_d(q, sh);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { atomicMin(mm, lo); atomicMax(mm + 1, hi); }
  if (threadIdx.x == 0) { atomicAdd(acc, s); atomicAdd(acc + 1, q); }
}

__device__ __forceinline__ float key_to_float(unsigned int k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k); }

// out8 = [mean_score, std_score, min_score, max_score, score_range, avg_edges, avg_sparsity, unique_graphs]
__global__ void mb_metrics_finalize_kernel(const float* __restrict__ x, long long n, int row_len, const double* __restrict__ acc, const unsigned int* __restrict__ mm,
                                           double* __restrict__ out8) {
  const double nn = n > 0 ? (double)n : 1.0, pivot = n > 0 ? (double)x[0] : 0.0;
  const double mean_d = acc[0] / nn;
  double var = acc[1] / nn - mean_d * mean_d;
  var = var > 0.0 ? var : 0.0;
  const double mn = n > 0 ? (double)key_to_float(mm[0]) : 0.0, mx = n > 0 ? (double)key_to_float(mm[1]) : 0.0;
  out8[0] = pivot + mean_d;
  out8[1] = sqrt(var);
  out8[2] = mn;
  out8[3] = mx;
  out8[4] = (double)((float)mx - (float)mn);            // the reference subtracts the two float32 values
  out8[5] = acc[5] / nn;
  out8[6] = acc[5] / nn / (double)row_len;
  out8[7] = acc[6];
}

__global__ void moving_average_kernel(const float* __restrict__ x, long long n, int w, double* __restrict__ out) {
  const double inv = 1.0 / (double)w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n - w + 1; i += (long long)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int k = 0; k < w; ++k) s += (double)x[i + k] * inv;
    out[i] = s;
  }
}

inline unsigned grid_for(long long n, int threads) {
  long long b = (n + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > 8LL * cvad_num_sms()) b = 8LL * cvad_num_sms();
  return (unsigned)b;
}

}  // namespace

CVAD_API long long cvad_eval_workspace_bytes(long long n) {
  if (n < 0) return -1;
  return pow2_at_least(n) * (long long)sizeof(KV) + 256;     // records + 16 doubles of accumulators + min/max keys
}

CVAD_API int cvad_sort_scores_f32(const float* scores, long long n, void* workspace, float* sorted, int* order, void* stream) {
  if (n <= 0) return 0;
  if (n >= (1LL << 31)) return (int)cudaErrorInvalidValue;
  cudaStream_t st = (cudaStream_t)stream;
  KV* d = (KV*)workspace;
  const long long n2 = pow2_at_least(n);
  fill_score_records_kernel<<<grid_for(n2, 256), 256, 0, st>>>(scores, n, n2, d);
  int e = sort_records(d, n2, st);
  if (e) return e;
  if (sorted || order) unpack_sorted_kernel<<<grid_for(n, 256), 256, 0, st>>>(d, scores, n, sorted, order);
  CVAD_LAUNCH_CHECK();
  return 0;
}

CVAD_API int cvad_percentile_sorted_f32(const float* sorted, long long n, float q, float* out, void* stream) {
  percentile_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(sorted, n, q, out);
  CVAD_LAUNCH_CHECK();
  return 0;
}

CVAD_API int cvad_threshold_labels_f32(const float* scores, long long n, const float* threshold, float* labels, void* stream) {
  if (n <= 0) return 0;
  threshold_labels_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(scores, n, threshold, labels);
  CVAD_LAUNCH_CHECK();
  return 0;
}

CVAD_API int cvad_roc_auc_f32(const float* scores, const float* targets, long long n, void* workspace, double* auc, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const long long n2 = pow2_at_least(n > 0 ? n : 1);
  KV* d = (KV*)workspace;
  double* acc = (double*)((char*)workspace + n2 * sizeof(KV));
  init_acc_kernel<<<1, 32, 0, st>>>(acc, (unsigned int*)(acc + 16));
  if (n > 0) {
    int e = cvad_sort_scores_f32(scores, n, workspace, nullptr, nullptr, stream);
    if (e) return e;
    auc_rank_kernel<<<grid_for(n, 256), 256, 0, st>>>(d, targets, n, acc);
  }
  auc_finalize_kernel<<<1, 1, 0, st>>>(acc, n, auc);
  CVAD_LAUNCH_CHECK();
  return 0;
}

CVAD_API int cvad_mb_eval_metrics_f32(const float* scores, const float* graphs, long long n, int row_len, float edge_threshold, void* workspace,
                                      double* out8, void* stream) {
  if (n >= (1LL << 31) || row_len <= 0) return (int)cudaErrorInvalidValue;
  cudaStream_t st = (cudaStream_t)stream;
  const long long n2 = pow2_at_least(n > 0 ? n : 1);
  KV* d = (KV*)workspace;
  double* acc = (double*)((char*)workspace + n2 * sizeof(KV));
  unsigned int* mm = (unsigned int*)(acc + 16);
  init_acc_kernel<<<1, 32, 0, st>>>(acc, mm);
  if (n > 0) {
    score_moments_kernel<<<grid_for(n, 256), 256, 0, st>>>(scores, n, acc, mm);
    graph_rows_kernel<<<grid_for(n2 * 32, 256), 256, 0, st>>>(graphs, n, row_len, edge_threshold, n2, d, acc);
    int e = sort_records(d, n2, st);
    if (e) return e;
    count_distinct_kernel<<<grid_for(n * 32, 256), 256, 0, st>>>(d, graphs, n, row_len, acc);
  }
  mb_metrics_finalize_kernel<<<1, 1, 0, st>>>(scores, n, row_len, acc, mm, out8);
  CVAD_LAUNCH_CHECK();
  return 0;
}

CVAD_API int cvad_moving_average_f32(const float* x, long long n, int w, double* out, void* stream) {
  if (w <= 0) return (int)cudaErrorInvalidValue;
  if (n < w) return 0;
  moving_average_kernel<<<grid_for(n - w + 1, 256), 256, 0, (cudaStream_t)stream>>>(x, n, w, out);
  CVAD_LAUNCH_CHECK();
  return 0;
}
