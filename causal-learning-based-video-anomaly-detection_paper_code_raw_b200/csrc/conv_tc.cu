// bf16 tensor-core implicit-GEMM 3x3 convolution for the M-A backbone (cad:128-139, 150-153): forward, data-gradient
// and weight-gradient as tcgen05.mma (UMMA) kernels with fp32 accumulators in TMEM.
//
// Layout in HBM: activations NHWC bf16 (channels innermost -> an im2col row chunk of 8 channels is one 16-byte
// vector), weights re-packed per step to bf16 K-major matrices ([Cout][tap][Cin] for forward, [Cin][tap][Cout] for
// dgrad); weight gradients are accumulated in fp32 straight into the OIHW slice of the flat gradient arena.
//
// One CTA (128 threads) computes a 128-row output tile:
//   * operands are gathered global->shared with 16-byte cp.async (zero-fill implements the conv padding, the
//     stride-divisibility test of dgrad and all tile tails), NSTAGE-deep ring;
//   * smem tiles use the UMMA canonical no-swizzle ("interleaved") layouts: 8x16B core matrices, K-major for
//     fwd/dgrad (LBO = 128 B between the two K-chunks of an MMA, SBO = 1 KiB between 8-row groups) and MN-major for
//     wgrad, where the reduction runs over pixels and both operands are channel-contiguous;
//   * one thread issues tcgen05.mma (M=128, N=BN, K=16) per 16-wide K step and tcgen05.commit's the stage's
//     mbarrier, which is what frees the smem slot for the next gather;
//   * the epilogue reads the accumulator with tcgen05.ld (32x32b: warp w owns TMEM lanes 32w..32w+31), adds the
//     bias and stores bf16 NHWC rows (fwd/dgrad) or atomically adds fp32 into dW (wgrad).
// SASS evidence: UTCHMMA / LDTM / LDGSTS (see profiles/).
#include "common.cuh"
#include "cvad_b200.h"

namespace {

constexpr int BM = 128;       // UMMA M (TMEM lanes)
constexpr int BK = 64;        // K elements per pipeline stage (4 MMAs of K=16)
constexpr int NSTAGE = 3;
constexpr int NT = 128;

enum { TC_FWD = 0, TC_DGRAD = 1, TC_WGRAD = 2 };

struct TcGeo {
  int N, H, W, C;        // "source" activation tensor that is gathered (x for fwd/wgrad, dy for dgrad)
  int Ho, Wo, Co;        // the other side: output of the GEMM rows (y for fwd, dx for dgrad); for wgrad: dy dims
  int stride, pad;       // conv stride / padding (3x3 kernel)
  long long rows;        // GEMM rows: fwd N*Ho*Wo, dgrad N*Hdx*Wdx, wgrad 9*Cin
  int K;                 // reduction length: fwd 9*Cin, dgrad 9*Cout, wgrad pixels-per-split handled separately
};

// ---------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// UMMA shared-memory descriptor, no swizzle, version 1 (Blackwell).  Offsets are given in bytes.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version
  return d;                 // base_offset 0, lbo_mode 0, layout_type SWIZZLE_NONE (0)
}
// instruction descriptor: bf16 x bf16 -> fp32, M=128, N=n
__device__ __forceinline__ uint32_t make_idesc(int n, int a_mn_major, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;                       // c_format = F32
  d |= 1u << 7;                       // a_format = BF16
  d |= 1u << 10;                      // b_format = BF16
  d |= (uint32_t)a_mn_major << 15;
  d |= (uint32_t)b_mn_major << 16;
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(BM >> 4) << 24;
  return d;
}

// ---------------------------------------------------------------------------------------------- fwd / dgrad kernel
// K-major tiles: element (row r, k) of a [rows x 64] stage lives at  (r%8)*16 + (k/8)*128 + (r/8)*1024  bytes.
struct RowInfo {
  int pix_base;     // n * H * W of the gathered tensor
  short c0h, c0w;   // fwd: oh*s-pad, ow*s-pad ; dgrad: ih+pad, iw+pad ; -30000 marks a tail row
};

template <int MODE, int BN>
__global__ void __launch_bounds__(NT) conv3x3_tc_kernel(TcGeo g, const __nv_bfloat16* __restrict__ src, const __nv_bfloat16* __restrict__ wpk,
                                                        const float* __restrict__ bias, __nv_bfloat16* __restrict__ dst, int ldd) {
  // src: NHWC bf16 (N,H,W,C) ; wpk: [Ntotal][K] bf16, K = 9*C ordered (tap, c) ; dst: rows x ldd bf16 (NHWC of the result)
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int A_BYTES = BM * BK * 2;    // 16 KiB
  constexpr int B_BYTES = BN * BK * 2;
  uint8_t* sA = smem;
  uint8_t* sB = smem + NSTAGE * A_BYTES;
  __shared__ uint64_t mma_bar[NSTAGE];
  __shared__ uint32_t tmem_base_sh;
  __shared__ RowInfo rinfo[BM];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int C = g.C;
  const int K = g.K;
  const int KB = (K + BK - 1) / BK;

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) mbar_init(&mma_bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  {   // row table
    long long m = m0 + tid;
    RowInfo ri;
    if (m < g.rows) {
      int Wr = MODE == TC_FWD ? g.Wo : g.Wo;   // rows are laid over (n, Ho, Wo) of the result tensor in both modes
      int Hr = g.Ho;
      int w = (int)(m % Wr);
      long long t = m / Wr;
      int h = (int)(t % Hr);
      int n = (int)(t / Hr);
      ri.pix_base = n * g.H * g.W;
      if (MODE == TC_FWD) { ri.c0h = (short)(h * g.stride - g.pad); ri.c0w = (short)(w * g.stride - g.pad); }
      else { ri.c0h = (short)(h + g.pad); ri.c0w = (short)(w + g.pad); }
    } else {
      ri.pix_base = 0; ri.c0h = -30000; ri.c0w = -30000;
    }
    rinfo[tid] = ri;
  }
  if (warp == 0) {   // TMEM allocation: BN fp32 columns (power of two >= 32)
    constexpr uint32_t COLS = BN < 32 ? 32 : BN;
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_sh)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_sh;

  // gather one stage: A = 128 rows x 8 chunks, B = BN rows x 8 chunks ; warp-instruction = 8 rows x 4 chunks
  auto gather = [&](int kb, int stage) {
    const uint32_t a_base = smem_u32(sA + stage * A_BYTES);
    const uint32_t b_base = smem_u32(sB + stage * B_BYTES);
    const int k0 = kb * BK;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int grp = j * 4 + warp;                       // 32 groups: 16 row-groups x 2 chunk-halves
      const int row = (grp >> 1) * 8 + (lane & 7);
      const int chunk = (grp & 1) * 4 + (lane >> 3);
      const int k = k0 + chunk * 8;
      const RowInfo ri = rinfo[row];
      bool ok = k < K;
      const __nv_bfloat16* p = src;
      if (ok) {
        const int tap = k / C, c = k - tap * C;
        const int kh = tap / 3, kw = tap - kh * 3;
        int hh, ww;
        if (MODE == TC_FWD) {
          hh = ri.c0h + kh; ww = ri.c0w + kw;
          ok = (unsigned)hh < (unsigned)g.H && (unsigned)ww < (unsigned)g.W;
        } else {
          int th = ri.c0h - kh, tw = ri.c0w - kw;
          hh = th / g.stride; ww = tw / g.stride;
          ok = th >= 0 && tw >= 0 && hh * g.stride == th && ww * g.stride == tw && hh < g.H && ww < g.W;
        }
        if (ok) p = src + ((long long)(ri.pix_base + hh * g.W + ww)) * C + c;
      }
      cp_async16(a_base + (row & 7) * 16 + chunk * 128 + (row >> 3) * 1024, p, ok);
    }
#pragma unroll
    for (int j = 0; j < BN / 16; ++j) {
      const int grp = j * 4 + warp;
      const int row = (grp >> 1) * 8 + (lane & 7);
      const int chunk = (grp & 1) * 4 + (lane >> 3);
      const int k = k0 + chunk * 8;
      const bool ok = k < K;
      const __nv_bfloat16* p = ok ? wpk + (long long)(n0 + row) * K + k : wpk;
      cp_async16(b_base + (row & 7) * 16 + chunk * 128 + (row >> 3) * 1024, p, ok);
    }
  };

  const uint32_t idesc = make_idesc(BN, 0, 0);
#pragma unroll
  for (int s = 0; s < NSTAGE - 1; ++s) {
    if (s < KB) gather(s, s);
    cp_async_commit();
  }
  for (int kb = 0; kb < KB; ++kb) {
    const int stage = kb % NSTAGE;
    cp_async_wait<NSTAGE - 2>();
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t a_base = smem_u32(sA + stage * A_BYTES);
      const uint32_t b_base = smem_u32(sB + stage * B_BYTES);
#pragma unroll
      for (int kk = 0; kk < BK / 16; ++kk) {
        const uint64_t da = make_desc(a_base + kk * 256, 128, 1024);
        const uint64_t db = make_desc(b_base + kk * 256, 128, 1024);
        tc_mma_bf16(tmem_d, da, db, idesc, (kb | kk) != 0);
      }
      tc_commit(&mma_bar[stage]);
    }
    // refill the slot used by k-block kb-1 with k-block kb+NSTAGE-1 once its MMAs have retired
    const int nk = kb + NSTAGE - 1;
    if (nk < KB) {
      if (kb >= 1) mbar_wait(&mma_bar[(kb - 1) % NSTAGE], ((kb - 1) / NSTAGE) & 1);
      gather(nk, nk % NSTAGE);
    }
    cp_async_commit();
  }
  mbar_wait(&mma_bar[(KB - 1) % NSTAGE], ((KB - 1) / NSTAGE) & 1);
  tc_fence_after();

  // ---- epilogue: warp w <-> TMEM lanes 32w.., thread <-> one output row
  {
    const long long m = m0 + warp * 32 + lane;
    const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
    __nv_bfloat16* out = dst + m * (long long)ldd + n0;
#pragma unroll
    for (int c0 = 0; c0 < BN; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(taddr + c0, v);
      tmem_ld_wait();
      if (m < g.rows) {
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float a = __uint_as_float(v[2 * i]), b = __uint_as_float(v[2 * i + 1]);
          if (bias) { a += __ldg(bias + n0 + c0 + 2 * i); b += __ldg(bias + n0 + c0 + 2 * i + 1); }
          __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
          pk[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(out + c0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(out + c0 + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    constexpr uint32_t COLS = BN < 32 ? 32 : BN;
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "n"(COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------- wgrad kernel
// D[(tap,ci) x co] += sum_pixels  X_im2col[pixel][(tap,ci)] * dY[pixel][co].  Both operands are MN-major:
// element (mn, k=pixel) of a stage lives at  (k%8)*16 + (mn/8)*128 + (k/8)*(ROWS/8*128)  bytes
// (core matrix = 8 pixels x 8 channels; SBO = 128 B between channel groups, LBO = ROWS*16 B between pixel groups).
template <int BN>
__global__ void __launch_bounds__(NT) conv3x3_wgrad_tc_kernel(TcGeo g, const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                                              float* __restrict__ dw, long long pix_per_split) {
  // x: NHWC (N,H,W,C) ; dy: NHWC (N,Ho,Wo,Co) ; dw: OIHW fp32 (Co, C, 3, 3), accumulated atomically
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int A_BYTES = BM * BK * 2;
  constexpr int B_BYTES = BN * BK * 2;
  uint8_t* sA = smem;
  uint8_t* sB = smem + NSTAGE * A_BYTES;
  __shared__ uint64_t mma_bar[NSTAGE];
  __shared__ uint32_t tmem_base_sh;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r0 = blockIdx.x * BM;          // first (tap,ci) row of this tile
  const int n0 = blockIdx.y * BN;          // first output channel
  const int C = g.C, Co = g.Co;
  const int KR = 9 * C;
  const long long P = (long long)g.N * g.Ho * g.Wo;
  const long long p_begin = (long long)blockIdx.z * pix_per_split;
  const long long p_end = p_begin + pix_per_split < P ? p_begin + pix_per_split : P;
  const int KB = p_begin < p_end ? (int)((p_end - p_begin + BK - 1) / BK) : 0;

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) mbar_init(&mma_bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    constexpr uint32_t COLS = BN < 32 ? 32 : BN;
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_sh)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_sh;

  // one stage = 64 pixels.  A: 16 channel-groups x 64 pixels ; B: BN/8 channel-groups x 64 pixels.
  // warp-instruction = 8 consecutive pixels x 4 consecutive channel groups.
  auto gather = [&](int kb, int stage) {
    const uint32_t a_base = smem_u32(sA + stage * A_BYTES);
    const uint32_t b_base = smem_u32(sB + stage * B_BYTES);
    const long long pb = p_begin + (long long)kb * BK;
    // this lane's pixel within each 8-pixel group is fixed (lane & 7); 8 pixel groups per stage
#pragma unroll
    for (int pg = 0; pg < 8; ++pg) {
      const long long pix = pb + pg * 8 + (lane & 7);
      const bool pok = pix < p_end;
      int ow = 0, oh = 0, n = 0;
      if (pok) {
        ow = (int)(pix % g.Wo);
        long long t = pix / g.Wo;
        oh = (int)(t % g.Ho);
        n = (int)(t / g.Ho);
      }
      // A: channel groups cg = warp*4 + (lane>>3) ... 16 groups per 128 rows -> one pass
      {
        const int cgp = warp * 4 + (lane >> 3);
        const int r = r0 + cgp * 8;
        bool ok = pok && r < KR;
        const __nv_bfloat16* p = x;
        if (ok) {
          const int tap = r / C, c = r - tap * C;
          const int kh = tap / 3, kw = tap - kh * 3;
          const int hh = oh * g.stride - g.pad + kh, ww = ow * g.stride - g.pad + kw;
          ok = (unsigned)hh < (unsigned)g.H && (unsigned)ww < (unsigned)g.W;
          if (ok) p = x + ((long long)(n * g.H + hh) * g.W + ww) * C + c;
        }
        cp_async16(a_base + (lane & 7) * 16 + cgp * 128 + pg * (BM / 8 * 128), p, ok);
      }
#pragma unroll
      for (int j = 0; j < BN / 128 + (BN % 128 ? 1 : 0); ++j) {
        const int cgp = j * 16 + warp * 4 + (lane >> 3);
        if (cgp < BN / 8) {
          const int co = n0 + cgp * 8;
          const bool ok = pok && co < Co;
          const __nv_bfloat16* p = ok ? dy + pix * Co + co : dy;
          cp_async16(b_base + (lane & 7) * 16 + cgp * 128 + pg * (BN / 8 * 128), p, ok);
        }
      }
    }
  };

  const uint32_t idesc = make_idesc(BN, 1, 1);
#pragma unroll
  for (int s = 0; s < NSTAGE - 1; ++s) {
    if (s < KB) gather(s, s);
    cp_async_commit();
  }
  for (int kb = 0; kb < KB; ++kb) {
    const int stage = kb % NSTAGE;
    cp_async_wait<NSTAGE - 2>();
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t a_base = smem_u32(sA + stage * A_BYTES);
      const uint32_t b_base = smem_u32(sB + stage * B_BYTES);
#pragma unroll
      for (int kk = 0; kk < BK / 16; ++kk) {
        // K = 16 pixels = 2 pixel groups; LBO = distance between pixel groups, SBO = between channel groups
        const uint64_t da = make_desc(a_base + kk * 2 * (BM / 8 * 128), BM / 8 * 128, 128);
        const uint64_t db = make_desc(b_base + kk * 2 * (BN / 8 * 128), BN / 8 * 128, 128);
        tc_mma_bf16(tmem_d, da, db, idesc, (kb | kk) != 0);
      }
      tc_commit(&mma_bar[stage]);
    }
    const int nk = kb + NSTAGE - 1;
    if (nk < KB) {
      if (kb >= 1) mbar_wait(&mma_bar[(kb - 1) % NSTAGE], ((kb - 1) / NSTAGE) & 1);
      gather(nk, nk % NSTAGE);
    }
    cp_async_commit();
  }
  if (KB > 0) {
    mbar_wait(&mma_bar[(KB - 1) % NSTAGE], ((KB - 1) / NSTAGE) & 1);
    tc_fence_after();
    const int r = r0 + warp * 32 + lane;       // (tap, ci) row owned by this thread
    const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
    const int tap = r / C, ci = r - tap * C;
#pragma unroll
    for (int c0 = 0; c0 < BN; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(taddr + c0, v);
      tmem_ld_wait();
      if (r < KR) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int co = n0 + c0 + i;
          if (co < Co) atomicAdd(dw + ((long long)co * C + ci) * 9 + tap, __uint_as_float(v[i]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    constexpr uint32_t COLS = BN < 32 ? 32 : BN;
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "n"(COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------- weight packing
// OIHW fp32 (Co, Ci, 3, 3) -> fwd [Co][tap][Ci] bf16 and dgrad [Ci][tap][Co] bf16
__global__ void pack_w3x3_kernel(const float* __restrict__ w, int Co, int Ci, __nv_bfloat16* __restrict__ wf, __nv_bfloat16* __restrict__ wd) {
  const int total = Co * Ci * 9;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int tap = i % 9, ci = (i / 9) % Ci, co = i / (9 * Ci);
    const __nv_bfloat16 v = __float2bfloat16(w[i]);
    if (wf) wf[((long long)co * 9 + tap) * Ci + ci] = v;
    if (wd) wd[((long long)ci * 9 + tap) * Co + co] = v;
  }
}

template <int MODE>
int launch_tc(const TcGeo& g, int ntotal, const __nv_bfloat16* src, const __nv_bfloat16* wpk, const float* bias, __nv_bfloat16* dst, int ldd,
              cudaStream_t st) {
  const unsigned gx = (unsigned)((g.rows + BM - 1) / BM);
#define TC_GO(BNV)                                                                                                          \
  {                                                                                                                         \
    constexpr int smem = NSTAGE * (BM * BK * 2 + BNV * BK * 2);                                                             \
    static bool set = false;                                                                                                \
    if (!set) {                                                                                                             \
      cudaFuncSetAttribute(conv3x3_tc_kernel<MODE, BNV>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);               \
      set = true;                                                                                                           \
    }                                                                                                                       \
    conv3x3_tc_kernel<MODE, BNV><<<dim3(gx, ntotal / BNV), NT, smem, st>>>(g, src, wpk, bias, dst, ldd);                    \
  }
  if (ntotal % 128 == 0) TC_GO(128)
  else if (ntotal % 64 == 0) TC_GO(64)
  else if (ntotal % 32 == 0) TC_GO(32)
  else return (int)cudaErrorInvalidValue;
#undef TC_GO
  CVAD_LAUNCH_CHECK();
  return 0;
}

}  // namespace

CVAD_API int cvad_pack_w3x3_bf16(const float* w, int Cout, int Cin, void* w_fwd, void* w_dgrad, void* stream) {
  int total = Cout * Cin * 9;
  pack_w3x3_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, Cout, Cin, (__nv_bfloat16*)w_fwd, (__nv_bfloat16*)w_dgrad);
  CVAD_LAUNCH_CHECK();
  return 0;
}

CVAD_API int cvad_conv3x3_fwd_bf16(const void* x, const void* w_fwd, const float* bias, void* y, int N, int H, int W, int Cin, int Cout,
                                   int stride, void* stream) {
  if (Cin % 8 || Cout % 32) return (int)cudaErrorInvalidValue;
  TcGeo g;
  g.N = N; g.H = H; g.W = W; g.C = Cin; g.stride = stride; g.pad = 1;
  g.Ho = (H + 2 - 3) / stride + 1; g.Wo = (W + 2 - 3) / stride + 1; g.Co = Cout;
  g.rows = (long long)N * g.Ho * g.Wo;
  g.K = 9 * Cin;
  return launch_tc<TC_FWD>(g, Cout, (const __nv_bfloat16*)x, (const __nv_bfloat16*)w_fwd, bias, (__nv_bfloat16*)y, Cout, (cudaStream_t)stream);
}

CVAD_API int cvad_conv3x3_dgrad_bf16(const void* dy, const void* w_dgrad, void* dx, int N, int H, int W, int Cin, int Cout, int stride,
                                     void* stream) {
  // (N,H,W,Cin) is the conv INPUT geometry (= dx); dy is (N,Ho,Wo,Cout)
  if (Cout % 8 || Cin % 32) return (int)cudaErrorInvalidValue;
  TcGeo g;
  const int Ho = (H + 2 - 3) / stride + 1, Wo = (W + 2 - 3) / stride + 1;
  g.N = N; g.H = Ho; g.W = Wo; g.C = Cout; g.stride = stride; g.pad = 1;   // gathered tensor = dy
  g.Ho = H; g.Wo = W; g.Co = Cin;                                           // GEMM rows run over dx pixels
  g.rows = (long long)N * H * W;
  g.K = 9 * Cout;
  return launch_tc<TC_DGRAD>(g, Cin, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)w_dgrad, nullptr, (__nv_bfloat16*)dx, Cin,
                             (cudaStream_t)stream);
}

CVAD_API int cvad_conv3x3_wgrad_bf16(const void* x, const void* dy, float* dw, int N, int H, int W, int Cin, int Cout, int stride,
                                     void* stream) {
  if (Cin % 8 || Cout % 32) return (int)cudaErrorInvalidValue;
  cudaStream_t st = (cudaStream_t)stream;
  TcGeo g;
  g.N = N; g.H = H; g.W = W; g.C = Cin; g.stride = stride; g.pad = 1;
  g.Ho = (H + 2 - 3) / stride + 1; g.Wo = (W + 2 - 3) / stride + 1; g.Co = Cout;
  g.rows = 9LL * Cin;
  g.K = 0;
  const long long P = (long long)N * g.Ho * g.Wo;
  const unsigned gx = (unsigned)((9 * Cin + BM - 1) / BM);
  const int bn = Cout % 128 == 0 ? 128 : (Cout % 64 == 0 ? 64 : 32);
  const unsigned gy = Cout / bn;
  long long want = (3LL * cvad_num_sms() + gx * gy - 1) / (gx * gy);
  long long maxs = (P + 4 * BK - 1) / (4 * BK);
  long long splits = want < 1 ? 1 : (want > maxs ? maxs : want);
  long long pps = ((P + splits - 1) / splits + BK - 1) / BK * BK;
  splits = (P + pps - 1) / pps;
#define WG_GO(BNV)                                                                                                          \
  {                                                                                                                         \
    constexpr int smem = NSTAGE * (BM * BK * 2 + BNV * BK * 2);                                                             \
    static bool set = false;                                                                                                \
    if (!set) {                                                                                                             \
      cudaFuncSetAttribute(conv3x3_wgrad_tc_kernel<BNV>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);                \
      set = true;                                                                                                           \
    }                                                                                                                       \
    conv3x3_wgrad_tc_kernel<BNV><<<dim3(gx, gy, (unsigned)splits), NT, smem, st>>>(g, (const __nv_bfloat16*)x,              \
                                                                                 (const __nv_bfloat16*)dy, dw, pps);       \
  }
  if (bn == 128) WG_GO(128)
  else if (bn == 64) WG_GO(64)
  else WG_GO(32)
#undef WG_GO
  CVAD_LAUNCH_CHECK();
  return 0;
}
