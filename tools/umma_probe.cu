// umma_probe: empirical check of tcgen05 shared-memory descriptor semantics on sm_100a (development tool, not product).
// Questions it answers (each config prints PASS/FAIL against a host model):
//   * TMA SWIZZLE_128B / SWIZZLE_64B tiles consumed by tcgen05.mma as K-major and MN-major operands
//   * descriptor start addresses shifted by whole rows (not multiples of the 8-row swizzle atom), with base_offset = 0
//     or base_offset = (addr >> 7) & 7
//   * MN-major atoms placed at a 64-byte LBO (overlapping "pixel shift" atoms), no-swizzle 16-byte shifts
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe tools/umma_probe.cu
// Run:   ./umma_probe <config>      (one config per process so a trap cannot hide the others)
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess) {                                                                   \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);          \
      exit(2);                                                                                 \
    }                                                                                          \
  } while (0)

struct MmaStep {
  uint32_t a_off, b_off;   // byte offsets added to the operand's smem base
};
struct ProbeParams {
  int use_tma;             // 1: TMA tiles ; 0: manual no-swizzle [chunk][row][16B] fill
  int a_boxes, b_boxes;    // number of TMA boxes per operand (column blocks)
  int a_box_cols, a_box_rows, b_box_cols, b_box_rows;
  int a_rows, a_cols, b_rows, b_cols;   // global tile dims (manual fill)
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;  // bytes
  int a_layout, b_layout;               // descriptor layout_type
  int a_mn, b_mn;                       // MN-major flags for the instruction descriptor
  int base_off_mode;                    // 1: base_offset = (addr >> 7) & 7
  int N;                                // UMMA N
  int nsteps;
  MmaStep steps[16];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t mk_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, int layout, int bo_mode) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  if (bo_mode) d |= (uint64_t)((saddr >> 7) & 7) << 49;
  d |= (uint64_t)(layout & 7) << 61;
  return d;
}

__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                                    const __nv_bfloat16* gA, const __nv_bfloat16* gB, ProbeParams p, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;              // 64 KiB
  uint8_t* sB = smem + 65536;      // 64 KiB
  __shared__ uint64_t bar_tma, bar_mma;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_tma)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_mma)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < 131072 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base;

  if (p.use_tma) {
    if (tid == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      uint32_t bytes = (uint32_t)(p.a_boxes * p.a_box_cols * p.a_box_rows * 2 + p.b_boxes * p.b_box_cols * p.b_box_rows * 2);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar_tma)), "r"(bytes) : "memory");
      for (int j = 0; j < p.a_boxes; ++j) {
        uint32_t dst = smem_u32(sA + (size_t)j * p.a_box_cols * p.a_box_rows * 2);
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                     "l"(&mapA), "r"(j * p.a_box_cols), "r"(0), "r"(smem_u32(&bar_tma))
                     : "memory");
      }
      for (int j = 0; j < p.b_boxes; ++j) {
        uint32_t dst = smem_u32(sB + (size_t)j * p.b_box_cols * p.b_box_rows * 2);
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                     "l"(&mapB), "r"(j * p.b_box_cols), "r"(0), "r"(smem_u32(&bar_tma))
                     : "memory");
      }
    }
    // wait
    {
      uint32_t done = 0;
      while (!done) {
        asm volatile("{ .reg .pred q; mbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0; selp.u32 %0, 1, 0, q; }" : "=r"(done)
                     : "r"(smem_u32(&bar_tma))
                     : "memory");
      }
    }
  } else {
    // manual [chunk][row][8 elem] fill
    for (int i = tid; i < p.a_rows * p.a_cols; i += 128) {
      int r = i / p.a_cols, c = i % p.a_cols;
      reinterpret_cast<__nv_bfloat16*>(sA)[(size_t)(c / 8) * p.a_rows * 8 + r * 8 + (c % 8)] = gA[i];
    }
    for (int i = tid; i < p.b_rows * p.b_cols; i += 128) {
      int r = i / p.b_cols, c = i % p.b_cols;
      reinterpret_cast<__nv_bfloat16*>(sB)[(size_t)(c / 8) * p.b_rows * 8 + r * 8 + (c % 8)] = gB[i];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0) {
    uint32_t idesc = 0;
    idesc |= 1u << 4;
    idesc |= 1u << 7;
    idesc |= 1u << 10;
    idesc |= (uint32_t)p.a_mn << 15;
    idesc |= (uint32_t)p.b_mn << 16;
    idesc |= (uint32_t)(p.N >> 3) << 17;
    idesc |= (uint32_t)(128 >> 4) << 24;
    for (int i = 0; i < p.nsteps; ++i) {
      uint64_t da = mk_desc(smem_u32(sA) + p.steps[i].a_off, p.a_lbo, p.a_sbo, p.a_layout, p.base_off_mode);
      uint64_t db = mk_desc(smem_u32(sB) + p.steps[i].b_off, p.b_lbo, p.b_sbo, p.b_layout, p.base_off_mode);
      uint32_t acc = i > 0;
      asm volatile("{ .reg .pred q; setp.ne.b32 q, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, q; }" ::"r"(tmem_d), "l"(da),
                   "l"(db), "r"(idesc), "r"(acc)
                   : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma)) : "memory");
  }
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile("{ .reg .pred q; mbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0; selp.u32 %0, 1, 0, q; }" : "=r"(done)
                   : "r"(smem_u32(&bar_mma))
                   : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
  for (int c0 = 0; c0 < p.N; c0 += 8) {
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr + c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 8; ++i) out[(warp * 32 + lane) * p.N + c0 + i] = __uint_as_float(v[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_d) : "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn get_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (!fn) { printf("no cuTensorMapEncodeTiled\n"); exit(2); }
  return (EncodeFn)fn;
}

static CUtensorMap make_map(EncodeFn enc, void* g, int rows, int cols, int box_rows, int box_cols, CUtensorMapSwizzle sw) {
  CUtensorMap m;
  memset(&m, 0, sizeof(m));
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(2); }
  return m;
}

int main(int argc, char** argv) {
  int cfg = argc > 1 ? atoi(argv[1]) : 0;
  EncodeFn enc = get_encode();
  ProbeParams p;
  memset(&p, 0, sizeof(p));
  int a_rows = 160, a_cols = 64, b_rows = 64, b_cols = 64;
  CUtensorMapSwizzle swA = CU_TENSOR_MAP_SWIZZLE_128B, swB = CU_TENSOR_MAP_SWIZZLE_128B;
  // logical model: kind 0: D[m][n] = sum_{step i} sum_{k<16} A(i,m,k) B(i,n,k)
  // A(i,m,k): K-major TA[ars + m][acs_i + k] ; MN-major TA[ars_i + k][acol(m)]
  int shiftA = 0, shiftB = 0;
  const char* name = "";
  enum { K_SW128, K_SW64, MN_SW128, MN_SW64_LBO64, MN_SW64_LBO0, K_NOSW, MN_NOSW, K_NOSW_STREAM } kind = K_SW128;
  switch (cfg) {
    case 0: kind = K_SW128; shiftA = 0; p.base_off_mode = 0; name = "K-major SW128 shift0"; break;
    case 1: kind = K_SW128; shiftA = 3; p.base_off_mode = 0; name = "K-major SW128 shiftA3 bo0"; break;
    case 2: kind = K_SW128; shiftA = 3; p.base_off_mode = 1; name = "K-major SW128 shiftA3 bo1"; break;
    case 3: kind = K_SW128; shiftA = 11; p.base_off_mode = 0; name = "K-major SW128 shiftA11 bo0"; break;
    case 4: kind = K_SW64; shiftA = 5; p.base_off_mode = 0; name = "K-major SW64 shiftA5 bo0"; break;
    case 5: kind = K_SW64; shiftA = 5; p.base_off_mode = 1; name = "K-major SW64 shiftA5 bo1"; break;
    case 6: kind = MN_SW128; shiftB = 0; p.base_off_mode = 0; name = "MN-major SW128 shift0"; break;
    case 7: kind = MN_SW128; shiftB = 3; p.base_off_mode = 0; name = "MN-major SW128 shiftB3 bo0"; break;
    case 8: kind = MN_SW128; shiftB = 3; p.base_off_mode = 1; name = "MN-major SW128 shiftB3 bo1"; break;
    case 9: kind = MN_SW64_LBO64; p.base_off_mode = 0; name = "MN-major SW64 A atoms at LBO=64B (4 pixel shifts x 32ch)"; break;
    case 10: kind = MN_SW64_LBO0; p.base_off_mode = 0; name = "MN-major SW64 A atoms at LBO=0"; break;
    case 11: kind = K_NOSW; shiftA = 3; name = "K-major no-swizzle shiftA3 (16B)"; break;
    case 12: kind = MN_NOSW; shiftB = 3; name = "MN-major no-swizzle shiftB3 (16B)"; break;
    case 13: kind = K_SW64; shiftA = 0; p.base_off_mode = 0; name = "K-major SW64 shift0"; break;
    case 14: kind = MN_SW128; shiftB = 11; p.base_off_mode = 0; name = "MN-major SW128 shiftB11 bo0"; break;
    case 15: kind = K_NOSW_STREAM; shiftA = 5; name = "K-major no-swizzle pixel stream: LBO=16B (overlapping K chunks), SBO=128B"; break;
    default: printf("unknown config\n"); return 1;
  }
  p.N = 32;
  p.use_tma = 1;
  switch (kind) {
    case K_SW128:
      a_rows = 160; a_cols = 64; b_rows = 32; b_cols = 64;
      p.a_boxes = 1; p.a_box_cols = 64; p.a_box_rows = 160; p.b_boxes = 1; p.b_box_cols = 64; p.b_box_rows = 32;
      p.a_lbo = 16; p.a_sbo = 1024; p.b_lbo = 16; p.b_sbo = 1024; p.a_layout = 2; p.b_layout = 2;
      p.nsteps = 4;
      for (int i = 0; i < 4; ++i) { p.steps[i].a_off = shiftA * 128 + i * 32; p.steps[i].b_off = i * 32; }
      break;
    case K_SW64:
      a_rows = 160; a_cols = 32; b_rows = 32; b_cols = 32; swA = swB = CU_TENSOR_MAP_SWIZZLE_64B;
      p.a_boxes = 1; p.a_box_cols = 32; p.a_box_rows = 160; p.b_boxes = 1; p.b_box_cols = 32; p.b_box_rows = 32;
      p.a_lbo = 16; p.a_sbo = 512; p.b_lbo = 16; p.b_sbo = 512; p.a_layout = 4; p.b_layout = 4;
      p.nsteps = 2;
      for (int i = 0; i < 2; ++i) { p.steps[i].a_off = shiftA * 64 + i * 32; p.steps[i].b_off = i * 32; }
      break;
    case MN_SW128:
      // A = TA[k][m], TA 48 rows x 128 cols as two boxes of 64 cols; B = TB[k + shiftB][n], TB 64 rows x 64 cols, N = 64
      a_rows = 48; a_cols = 128; b_rows = 64; b_cols = 64; p.N = 64;
      p.a_boxes = 2; p.a_box_cols = 64; p.a_box_rows = 48; p.b_boxes = 1; p.b_box_cols = 64; p.b_box_rows = 64;
      p.a_lbo = 48 * 128; p.a_sbo = 1024; p.b_lbo = 64 * 128; p.b_sbo = 1024; p.a_layout = 2; p.b_layout = 2; p.a_mn = 1; p.b_mn = 1;
      p.nsteps = 2;
      for (int i = 0; i < 2; ++i) { p.steps[i].a_off = i * 16 * 128; p.steps[i].b_off = (shiftB + i * 16) * 128; }
      break;
    case MN_SW64_LBO64:
    case MN_SW64_LBO0:
      // A = TA[k + j][ci], m = j*32 + ci (LBO = 64 B) or TA[k][ci] (LBO = 0); TA 48 rows x 32 cols; B = TB[k][n], TB 32 x 32
      a_rows = 48; a_cols = 32; b_rows = 32; b_cols = 32; p.N = 32; swA = swB = CU_TENSOR_MAP_SWIZZLE_64B;
      p.a_boxes = 1; p.a_box_cols = 32; p.a_box_rows = 48; p.b_boxes = 1; p.b_box_cols = 32; p.b_box_rows = 32;
      p.a_lbo = kind == MN_SW64_LBO64 ? 64 : 0; p.a_sbo = 512; p.b_lbo = 64; p.b_sbo = 512; p.a_layout = 4; p.b_layout = 4; p.a_mn = 1; p.b_mn = 1;
      p.nsteps = 2;
      for (int i = 0; i < 2; ++i) { p.steps[i].a_off = i * 16 * 64; p.steps[i].b_off = i * 16 * 64; }
      break;
    case K_NOSW:
      p.use_tma = 0;
      a_rows = 160; a_cols = 64; b_rows = 32; b_cols = 64;
      p.a_lbo = 160 * 16; p.a_sbo = 128; p.b_lbo = 32 * 16; p.b_sbo = 128; p.a_layout = 0; p.b_layout = 0;
      p.nsteps = 4;
      for (int i = 0; i < 4; ++i) { p.steps[i].a_off = shiftA * 16 + i * 2 * 160 * 16; p.steps[i].b_off = i * 2 * 32 * 16; }
      break;
    case K_NOSW_STREAM:
      // A operand = a linear stream of 16-byte pixels TA[p][8]; row m, K chunk c reads pixel (shift + m + c): LBO = 16 B, SBO = 128 B.
      // B = TB[n][k] (32 x 32) in the [chunk][row][16B] layout.  Two MMAs of K=16 (the second starts two pixels later).
      p.use_tma = 0;
      a_rows = 192; a_cols = 8; b_rows = 32; b_cols = 32;
      p.a_lbo = 16; p.a_sbo = 128; p.b_lbo = 32 * 16; p.b_sbo = 128; p.a_layout = 0; p.b_layout = 0;
      p.nsteps = 2;
      for (int i = 0; i < 2; ++i) { p.steps[i].a_off = (shiftA + 2 * i) * 16; p.steps[i].b_off = i * 2 * 32 * 16; }
      break;
    case MN_NOSW:
      // A = TA[k][m] (48 x 128), B = TB[k + shiftB][n] (64 x 64), [chunk][row][16B]: M-group stride (SBO) = rows*16, K-group stride (LBO) = 128
      p.use_tma = 0;
      a_rows = 48; a_cols = 128; b_rows = 64; b_cols = 64; p.N = 64;
      p.a_lbo = 128; p.a_sbo = 48 * 16; p.b_lbo = 128; p.b_sbo = 64 * 16; p.a_layout = 0; p.b_layout = 0; p.a_mn = 1; p.b_mn = 1;
      p.nsteps = 2;
      for (int i = 0; i < 2; ++i) { p.steps[i].a_off = i * 16 * 16; p.steps[i].b_off = (shiftB + i * 16) * 16; }
      break;
  }
  p.a_rows = a_rows; p.a_cols = a_cols; p.b_rows = b_rows; p.b_cols = b_cols;

  std::vector<__nv_bfloat16> hA((size_t)a_rows * a_cols), hB((size_t)b_rows * b_cols);
  std::vector<float> fA(hA.size()), fB(hB.size());
  srand(123 + cfg);
  for (size_t i = 0; i < hA.size(); ++i) { fA[i] = (float)((rand() % 15) - 7); hA[i] = __float2bfloat16(fA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { fB[i] = (float)((rand() % 15) - 7); hB[i] = __float2bfloat16(fB[i]); }
  __nv_bfloat16 *dA, *dB;
  float* dOut;
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMalloc(&dOut, 128 * 256 * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dOut, 0, 128 * 256 * 4));
  CUtensorMap mA, mB;
  memset(&mA, 0, sizeof(mA)); memset(&mB, 0, sizeof(mB));
  if (p.use_tma) {
    mA = make_map(enc, dA, a_rows, a_cols, p.a_box_rows, p.a_box_cols, swA);
    mB = make_map(enc, dB, b_rows, b_cols, p.b_box_rows, p.b_box_cols, swB);
  }
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
  probe_kernel<<<1, 128, 131072>>>(mA, mB, dA, dB, p, dOut);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> out(128 * p.N);
  CK(cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost));

  auto TA = [&](int r, int c) { return (r < a_rows && c < a_cols) ? fA[(size_t)r * a_cols + c] : 0.f; };
  auto TB = [&](int r, int c) { return (r < b_rows && c < b_cols) ? fB[(size_t)r * b_cols + c] : 0.f; };
  int bad = 0;
  double maxerr = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < p.N; ++n) {
      double ref = 0;
      switch (kind) {
        case K_SW128: case K_NOSW: for (int k = 0; k < 64; ++k) ref += TA(shiftA + m, k) * TB(n, k); break;
        case K_SW64: for (int k = 0; k < 32; ++k) ref += TA(shiftA + m, k) * TB(n, k); break;
        case MN_SW128: case MN_NOSW: for (int k = 0; k < 32; ++k) ref += TA(k, m) * TB(k + shiftB, n); break;
        case MN_SW64_LBO64: for (int k = 0; k < 32; ++k) ref += TA(k + m / 32, m % 32) * TB(k, n); break;
        case MN_SW64_LBO0: for (int k = 0; k < 32; ++k) ref += TA(k, m % 32) * TB(k, n); break;
        case K_NOSW_STREAM: for (int k = 0; k < 32; ++k) ref += TA(shiftA + m + k / 8, k % 8) * TB(n, k); break;
      }
      double e = fabs(ref - out[m * p.N + n]);
      if (e > maxerr) maxerr = e;
      if (e > 1e-3) ++bad;
    }
  printf("cfg %2d %-58s : %s (bad %d / %d, max err %.3g)\n", cfg, name, bad ? "FAIL" : "PASS", bad, 128 * p.N, maxerr);
  return 0;
}
