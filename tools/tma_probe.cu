// tma_probe: how fast can one SM pull [rows x 64/128 B] boxes of a row-major bf16 matrix into swizzled shared memory?
// (development tool).  Persistent CTAs, one elected thread issues TMA loads into a ring of `depth` slots and waits for
// them; no MMA.  Prints aggregate GB/s for a few box shapes / ring depths / L2-promotion settings.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/bin/tma_probe tools/tma_probe.cu
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                        \
    }                                                                                 \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap map, int box_rows, int row_bytes, int depth, long long total_rows,
                                             int rows_per_cta, int shift_stride, unsigned long long* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[16];
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  const int box_bytes = box_rows * row_bytes;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long r_begin = (long long)blockIdx.x * rows_per_cta;
    const int n = rows_per_cta / box_rows;
    int issued = 0, done = 0;
    while (done < n) {
      while (issued < n && issued - done < depth) {
        const int s = issued % depth;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(box_bytes) : "memory");
        long long row = (r_begin + (long long)issued * (shift_stride ? shift_stride : box_rows)) % (total_rows - box_rows);
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                         base + s * box_bytes),
                     "l"(&map), "r"(0), "r"((int)row), "r"(smem_u32(&bar[s]))
                     : "memory");
        ++issued;
      }
      const int s = done % depth;
      const uint32_t parity = (done / depth) & 1;
      uint32_t ok = 0;
      while (!ok) asm volatile("{ .reg .pred q; mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2; selp.u32 %0, 1, 0, q; }" : "=r"(ok) : "r"(smem_u32(&bar[s])), "r"(parity) : "memory");
      ++done;
    }
    if (sink) atomicAdd(sink, (unsigned long long)done);
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeFn enc = (EncodeFn)fn;
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const long long total_rows = 4LL << 20;     // 4M rows
  void* buf;
  CK(cudaMalloc(&buf, total_rows * 128));
  CK(cudaMemset(buf, 1, total_rows * 128));
  unsigned long long* sink;
  CK(cudaMalloc(&sink, 8));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  struct Cfg { int row_bytes, box_rows, depth, promo, stride; const char* note; };
  Cfg cfgs[] = {
      {128, 64, 2, 2, 0, "128B x 64 rows, depth 2, promo 256B"},  {128, 64, 4, 2, 0, "128B x 64 rows, depth 4"},
      {128, 64, 8, 2, 0, "128B x 64 rows, depth 8"},              {128, 256, 2, 2, 0, "128B x 256 rows, depth 2"},
      {128, 256, 4, 2, 0, "128B x 256 rows, depth 4"},            {128, 128, 8, 2, 0, "128B x 128 rows, depth 8"},
      {128, 256, 4, 0, 0, "128B x 256 rows, depth 4, promo none"}, {128, 256, 4, 1, 0, "128B x 256 rows, depth 4, promo 128B"},
      {64, 64, 4, 2, 0, "64B x 64 rows, depth 4"},                {64, 256, 4, 2, 0, "64B x 256 rows, depth 4"},
      {64, 256, 8, 2, 0, "64B x 256 rows, depth 8"},              {64, 32, 8, 2, 0, "64B x 32 rows, depth 8"},
      {128, 128, 4, 2, 0, "128B x 128 rows, depth 4"},            {128, 256, 4, 2, 64, "128B x 256 rows, depth 4, overlapping windows (stride 64 rows)"},
  };
  for (const Cfg& c : cfgs) {
    CUtensorMap m;
    memset(&m, 0, sizeof(m));
    const int cols = c.row_bytes / 2;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)(total_rows * 128 / c.row_bytes)};
    cuuint64_t strides[1] = {(cuuint64_t)c.row_bytes};
    cuuint32_t box[2] = {(cuuint32_t)cols, (cuuint32_t)c.box_rows};
    cuuint32_t es[2] = {1, 1};
    CUtensorMapL2promotion pr = c.promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : (c.promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE);
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     c.row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
    const long long rows_total = (long long)dims[1];
    const int rows_per_cta = 16384;
    const size_t smem = (size_t)c.depth * c.box_rows * c.row_bytes + 1024;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    probe<<<sms, 128, smem>>>(m, c.box_rows, c.row_bytes, c.depth, rows_total, rows_per_cta, c.stride, sink);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int it = 0; it < 5; ++it) probe<<<sms, 128, smem>>>(m, c.box_rows, c.row_bytes, c.depth, rows_total, rows_per_cta, c.stride, sink);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double bytes = 5.0 * sms * (double)(rows_per_cta / c.box_rows) * c.box_rows * c.row_bytes;
    printf("%-70s : %8.1f GB/s total, %6.2f B/cycle/SM @1.9GHz, %.1f cycles per box\n", c.note, bytes / ms / 1e6, bytes / ms / 1e6 / sms / 1.9,
           (double)ms * 1e-3 * 1.9e9 / (5.0 * (rows_per_cta / c.box_rows)));
  }
  return 0;
}
