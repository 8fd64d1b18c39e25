// Stem of the M-A backbone on the tensor cores: conv 7x7 stride 2 pad 3, 1 -> 32 channels (cad:115, 145) as a
// tcgen05 kind::tf32 implicit GEMM, fused with the BatchNorm batch statistics (pass 1) or with BN + ReLU (pass 2);
// MaxPool2d(3,2,1) (cad:118,148) follows as a 16-byte-vector bandwidth kernel that writes the padded-flat bf16 layout.
//
// GEMM view: M = output positions (N*Ho*Wo, 128 per tile), N = 32 channels, K = 49 taps padded to 56 (7 MMAs of K=8).
// With a single input channel an im2col row has no contiguous 16-byte piece, so TMA cannot build the A tile; instead the
// four "builder" warps gather it: thread r loads the 49 fp32 pixels of its output position straight from global memory
// (neighbouring threads' windows overlap, so L1 serves ~12 of every 13 reads) and writes one 256-byte K-major row into the
// SWIZZLE_128B layout the MMA descriptor expects.  A fifth warp issues the MMAs; the builder warps then read the 128x32
// fp32 accumulator back from TMEM (they own TMEM lanes 32w..32w+31) and either accumulate per-channel sum / sum of
// squares in registers across all tiles of the CTA (pass 1: no conv output is ever written), or apply BN + ReLU and
// store bf16 NHWC rows (pass 2).  A tiles and accumulators are double buffered so gather, MMA and epilogue overlap.
// The frozen stem needs no gradient (cad:596-598).  fp32 inputs are rounded to tf32 (10-bit mantissa) by the MMA.
#include "common.cuh"
#include "cvad_b200.h"
#include "tc_common.cuh"

namespace {

using namespace cvad_tc;

constexpr int ST_C = 32;             // output channels
constexpr int ST_K = 7;              // kernel size
constexpr int ST_A_BYTES = 2 * 128 * 128;   // two 32-float K slabs of 128 rows
constexpr int ST_W_BYTES = 2 * ST_C * 128;

struct StemGeo {
  int N, H, W, Ho, Wo;
  long long total;        // N*Ho*Wo
};

__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// MODE 0: statistics (ws[c] += sum y, ws[32+c] += sum y^2, y = conv + bias).  MODE 1: out = relu(y*sc + sh) bf16 NHWC.
template <int MODE>
__global__ void __launch_bounds__(160) stem_tf32_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                        StemGeo g, const float* __restrict__ mean, const float* __restrict__ invstd,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        double* __restrict__ ws, __nv_bfloat16* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_a_full[2], bar_a_empty[2], bar_acc_full[2], bar_acc_empty[2];
  __shared__ uint32_t tmem_base_sh;
  __shared__ float s_sc[ST_C], s_sh[ST_C];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t s_a = smem_base;                    // 2 stages x 32 KiB
  const uint32_t s_w = smem_base + 2 * ST_A_BYTES;   // 8 KiB
  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
  const long long n_tiles = (g.total + 127) / 128;

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_a_full[i], 128);
      mbar_init(&bar_a_empty[i], 1);
      mbar_init(&bar_acc_full[i], 1);
      mbar_init(&bar_acc_empty[i], 128);
    }
    fence_barrier_init();
  }
  if (tid < ST_C) {
    if (MODE == 1) {
      const float sc = invstd[tid] * gamma[tid];
      s_sc[tid] = sc;
      s_sh[tid] = (bias[tid] - mean[tid]) * sc + beta[tid];
    } else {
      s_sc[tid] = 1.f;
      s_sh[tid] = bias[tid];
    }
  }
  // weights -> K-major SWIZZLE_128B tile [32 rows][64 floats] (taps 49..63 zero)
  for (int i = tid; i < ST_C * 64; i += blockDim.x) {
    const int n = i >> 6, k = i & 63;
    const float v = k < ST_K * ST_K ? w[n * ST_K * ST_K + k] : 0.f;
    const int slab = k >> 5, chunk = (k & 31) >> 2, e = k & 3;
    *reinterpret_cast<float*>(smem_gen + 2 * ST_A_BYTES + slab * (ST_C * 128) + n * 128 + ((chunk ^ (n & 7)) << 4) + e * 4) = v;
  }
  if (warp == 4) tmem_alloc<64>(&tmem_base_sh);
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_sh;

  if (warp < 4) {
    // ---------------------------------------------------------------- builders + epilogue: thread <-> one tile row / TMEM lane
    const int r = tid;
    float s[ST_C], q[ST_C];
    if (MODE == 0) {
#pragma unroll
      for (int c = 0; c < ST_C; ++c) { s[c] = 0.f; q[c] = 0.f; }
    }
    const int HoWo = g.Ho * g.Wo;
    int it = 0;
    long long prev_m = -1;
    for (long long t = blockIdx.x; ; t += gridDim.x, ++it) {
      const bool have = t < n_tiles;
      if (have) {
        const int st = it & 1;
        mbar_wait(&bar_a_empty[st], ((it >> 1) & 1) ^ 1);
        const long long m = t * 128 + r;
        float v[56];
#pragma unroll
        for (int i = 0; i < 56; ++i) v[i] = 0.f;
        if (m < g.total) {
          const int n = (int)(m / HoWo);
          const int rem = (int)(m - (long long)n * HoWo);
          const int oh = rem / g.Wo, ow = rem - oh * g.Wo;
          const float* xn = x + (long long)n * g.H * g.W;
          const int iw0 = 2 * ow - 3;
#pragma unroll
          for (int kh = 0; kh < ST_K; ++kh) {
            const int ih = 2 * oh + kh - 3;
            if ((unsigned)ih < (unsigned)g.H) {
              const float* xr = xn + (long long)ih * g.W;
#pragma unroll
              for (int kw = 0; kw < ST_K; ++kw) {
                const int iw = iw0 + kw;
                if ((unsigned)iw < (unsigned)g.W) v[kh * ST_K + kw] = __ldg(xr + iw);
              }
            }
          }
        }
        uint8_t* arow = smem_gen + st * ST_A_BYTES + r * 128;
#pragma unroll
        for (int j = 0; j < 14; ++j) {
          const int slab = j >> 3, chunk = j & 7;
          *reinterpret_cast<float4*>(arow + slab * (128 * 128) + ((chunk ^ (r & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        mbar_arrive(&bar_a_full[st]);
      }
      // epilogue of the previous tile
      if (it > 0) {
        const int pst = (it - 1) & 1;
        mbar_wait(&bar_acc_full[pst], ((it - 1) >> 1) & 1);
        tc_fence_after();
        uint32_t a0[16], a1[16];
        const uint32_t taddr = tmem_base + pst * ST_C + ((uint32_t)(warp * 32) << 16);
        tmem_ld16(taddr, a0);
        tmem_ld16(taddr + 16, a1);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&bar_acc_empty[pst]);
        if (prev_m < g.total) {
          if (MODE == 0) {
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              const float y0 = __uint_as_float(a0[c]) + s_sh[c], y1 = __uint_as_float(a1[c]) + s_sh[16 + c];
              s[c] += y0; q[c] = fmaf(y0, y0, q[c]);
              s[16 + c] += y1; q[16 + c] = fmaf(y1, y1, q[16 + c]);
            }
          } else {
            uint32_t pk[16];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const float y0 = fmaxf(fmaf(__uint_as_float(a0[2 * c]), s_sc[2 * c], s_sh[2 * c]), 0.f);
              const float y1 = fmaxf(fmaf(__uint_as_float(a0[2 * c + 1]), s_sc[2 * c + 1], s_sh[2 * c + 1]), 0.f);
              const float z0 = fmaxf(fmaf(__uint_as_float(a1[2 * c]), s_sc[16 + 2 * c], s_sh[16 + 2 * c]), 0.f);
              const float z1 = fmaxf(fmaf(__uint_as_float(a1[2 * c + 1]), s_sc[16 + 2 * c + 1], s_sh[16 + 2 * c + 1]), 0.f);
              __nv_bfloat162 h0 = __floats2bfloat162_rn(y0, y1), h1 = __floats2bfloat162_rn(z0, z1);
              pk[c] = *reinterpret_cast<uint32_t*>(&h0);
              pk[8 + c] = *reinterpret_cast<uint32_t*>(&h1);
            }
            uint4* o = reinterpret_cast<uint4*>(out + prev_m * ST_C);
            o[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            o[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            o[2] = make_uint4(pk[8], pk[9], pk[10], pk[11]);
            o[3] = make_uint4(pk[12], pk[13], pk[14], pk[15]);
          }
        }
      }
      if (!have) break;
      prev_m = t * 128 + r;
    }
    if (MODE == 0) {
      // CTA reduction over the 128 row-threads (reuses A stage 0: every MMA has retired, see the last acc_full wait)
      float* red = reinterpret_cast<float*>(smem_gen);          // [128][65]
      asm volatile("bar.sync 1, 128;\n" ::: "memory");
#pragma unroll
      for (int c = 0; c < ST_C; ++c) { red[r * 65 + c] = s[c]; red[r * 65 + 32 + c] = q[c]; }
      asm volatile("bar.sync 1, 128;\n" ::: "memory");
      if (r < 64) {
        double acc = 0.0;
        for (int i = 0; i < 128; ++i) acc += (double)red[i * 65 + r];
        atomicAdd(ws + r, acc);
      }
    }
  } else {
    // ---------------------------------------------------------------- MMA issuer
    uint32_t idesc = 0;
    idesc |= 1u << 4;                 // D = f32
    idesc |= 2u << 7;                 // A = tf32
    idesc |= 2u << 10;                // B = tf32
    idesc |= (uint32_t)(ST_C >> 3) << 17;
    idesc |= (uint32_t)(128 >> 4) << 24;
    const uint64_t desc_hi = make_smem_desc(0, 16, 1024, UMMA_SW128);
    int it = 0;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const int st = it & 1;
      mbar_wait(&bar_a_full[st], (it >> 1) & 1);
      mbar_wait(&bar_acc_empty[st], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 7; ++ks) {
          const uint32_t a_addr = s_a + st * ST_A_BYTES + (ks >> 2) * (128 * 128) + (ks & 3) * 32;
          const uint32_t b_addr = s_w + (ks >> 2) * (ST_C * 128) + (ks & 3) * 32;
          const uint64_t da = desc_hi | (uint64_t)((a_addr >> 4) & 0x3FFF);
          const uint64_t db = desc_hi | (uint64_t)((b_addr >> 4) & 0x3FFF);
          tc_mma_tf32(tmem_base + st * ST_C, da, db, idesc, ks != 0);
        }
        tc_commit(&bar_a_empty[st]);
        tc_commit(&bar_acc_full[st]);
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<64>(tmem_base);
}

__global__ void stem_finalize_kernel(double* __restrict__ ws, double count, float eps, float momentum, float* __restrict__ mean,
                                     float* __restrict__ invstd, float* __restrict__ running_mean, float* __restrict__ running_var,
                                     long long* __restrict__ nbt) {
  const int c = threadIdx.x;
  if (c < ST_C) {
    const double m = ws[c] / count;
    double var = ws[ST_C + c] / count - m * m;
    if (var < 0.0) var = 0.0;
    mean[c] = (float)m;
    invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) {
      const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
    ws[c] = 0.0;
    ws[ST_C + c] = 0.0;
  }
  if (c == 0 && nbt) *nbt += 1;
}

// MaxPool2d(3,2,1) over relu'd bf16 NHWC (N,H,W,C) -> padded-flat (N,PH+2,PW+2,C) with zero border; 8 channels per thread
__global__ void maxpool3x3s2_pad_kernel(const __nv_bfloat16* __restrict__ y, int N, int H, int W, int C, int PH, int PW,
                                        __nv_bfloat16* __restrict__ out) {
  const int groups = C >> 3;
  const int rowlen = (PW + 2) * groups;
  const int rows = N * (PH + 2);
  for (int rr = blockIdx.x; rr < rows; rr += gridDim.x) {
    const int php = rr % (PH + 2), n = rr / (PH + 2);
    uint4* dst = reinterpret_cast<uint4*>(out) + (long long)rr * rowlen;
    const bool row_ok = php >= 1 && php <= PH;
    for (int v = threadIdx.x; v < rowlen; v += blockDim.x) {
      const int pwp = v / groups, cg = v - pwp * groups;
      uint4 o = make_uint4(0, 0, 0, 0);
      if (row_ok && pwp >= 1 && pwp <= PW) {
        const int ph = php - 1, pw = pwp - 1;
        __nv_bfloat162 best[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) best[i] = __floats2bfloat162_rn(0.f, 0.f);     // inputs are post-ReLU (>= 0)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const int h = 2 * ph - 1 + a;
          if ((unsigned)h >= (unsigned)H) continue;
#pragma unroll
          for (int b = 0; b < 3; ++b) {
            const int ww = 2 * pw - 1 + b;
            if ((unsigned)ww >= (unsigned)W) continue;
            const uint4 t = __ldg(reinterpret_cast<const uint4*>(y) + (((long long)n * H + h) * W + ww) * groups + cg);
            const __nv_bfloat162* tp = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
            for (int i = 0; i < 4; ++i) best[i] = __hmax2(best[i], tp[i]);
          }
        }
        o = *reinterpret_cast<uint4*>(best);
      }
      dst[v] = o;
    }
  }
}

int stem_launch(int mode, const float* x, const float* w, const float* bias, int N, int H, int W, const float* mean, const float* invstd,
                const float* gamma, const float* beta, double* ws, void* out, cudaStream_t st) {
  StemGeo g;
  g.N = N; g.H = H; g.W = W;
  g.Ho = (H + 6 - 7) / 2 + 1;
  g.Wo = (W + 6 - 7) / 2 + 1;
  g.total = (long long)N * g.Ho * g.Wo;
  const size_t smem = 2 * ST_A_BYTES + ST_W_BYTES + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(stem_tf32_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(stem_tf32_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const long long n_tiles = (g.total + 127) / 128;
  long long grid = 3LL * cvad_num_sms();
  if (grid > n_tiles) grid = n_tiles;
  if (mode == 0)
    stem_tf32_kernel<0><<<(unsigned)grid, 160, smem, st>>>(x, w, bias, g, mean, invstd, gamma, beta, ws, (__nv_bfloat16*)out);
  else
    stem_tf32_kernel<1><<<(unsigned)grid, 160, smem, st>>>(x, w, bias, g, mean, invstd, gamma, beta, ws, (__nv_bfloat16*)out);
  CVAD_LAUNCH_CHECK();
  return 0;
}

}  // namespace

CVAD_API int cvad_stem_tf32_stats(const float* x, const float* w, const float* bias, int N, int H, int W, double* ws, float eps, float momentum,
                                  float* mean, float* invstd, float* running_mean, float* running_var, long long* num_batches_tracked,
                                  void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  int e = stem_launch(0, x, w, bias, N, H, W, nullptr, nullptr, nullptr, nullptr, ws, nullptr, st);
  if (e) return e;
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  stem_finalize_kernel<<<1, 32, 0, st>>>(ws, (double)N * Ho * Wo, eps, momentum, mean, invstd, running_mean, running_var, num_batches_tracked);
  CVAD_LAUNCH_CHECK();
  return 0;
}

CVAD_API int cvad_stem_tf32_bn_relu(const float* x, const float* w, const float* bias, int N, int H, int W, const float* mean,
                                    const float* invstd, const float* gamma, const float* beta, void* y, void* stream) {
  return stem_launch(1, x, w, bias, N, H, W, mean, invstd, gamma, beta, nullptr, y, (cudaStream_t)stream);
}

CVAD_API int cvad_pad_maxpool3x3s2_bf16(const void* y, int N, int H, int W, int C, void* out, void* stream) {
  if (C % 8) return (int)cudaErrorInvalidValue;
  const int PH = (H - 1) / 2 + 1, PW = (W - 1) / 2 + 1;
  const int rows = N * (PH + 2);
  const int blocks = rows < 16 * cvad_num_sms() ? rows : 16 * cvad_num_sms();
  maxpool3x3s2_pad_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)y, N, H, W, C, PH, PW, (__nv_bfloat16*)out);
  CVAD_LAUNCH_CHECK();
  return 0;
}
