// How long does the ISSUING warp spend per TMA request?  (design input for ulv_qr3's feed)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_issue tma_issue.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_tx(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ bool mbar_test(uint64_t* b, uint32_t ph) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(s32(b)), "r"(ph) : "memory");
  return ok;
}
__device__ __forceinline__ void bulk(double* d, const double* s, uint32_t n, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(d)), "l"(s), "r"(n), "r"(s32(b)) : "memory");
}
__device__ __forceinline__ void tile2d(double* d, const void* tm, int c0, int c1, uint64_t* b) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(s32(d)), "l"(tm), "r"(c0), "r"(c1), "r"(s32(b)) : "memory");
}

// mode 0: 1 lane issues one 2 KB bulk copy; 1: 16 lanes issue 512 B bulk copies; 2: 1 lane one 16x16 tile;
// 3: 4 lanes one tile each; 4: only expect_tx + test (no copy)
__global__ void k(const double* A, const unsigned char* tmap, int mode, int reps, long long* out) {
  extern __shared__ __align__(1024) double sm[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 8192);
  const int lane = threadIdx.x;
  if (lane == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncwarp();
  long long t_issue = 0, t_done = 0;
  for (int r = 0; r < reps; r++) {
    const long long t0 = clock64();
    if (mode == 0) { if (lane == 0) { mbar_tx(bar, 2048); bulk(sm, A + r * 256, 2048, bar); } }
    else if (mode == 1) { if (lane == 0) mbar_tx(bar, 16 * 512); __syncwarp(); if (lane < 16) bulk(sm + lane * 64, A + lane * 256 + r * 8, 512, bar); }
    else if (mode == 2) { if (lane == 0) { mbar_tx(bar, 2048); tile2d(sm, tmap, 16 * (r & 7), 16 * (r & 3), bar); } }
    else if (mode == 3) { if (lane == 0) mbar_tx(bar, 4 * 2048); __syncwarp(); if (lane < 4) tile2d(sm + lane * 256, tmap, 16 * lane + 64 * (r & 1), 16 * (r & 3), bar); }
    else { if (lane == 0) mbar_tx(bar, 0); }
    __syncwarp();
    const long long t1 = clock64();
    while (!mbar_test(bar, r & 1)) { }
    const long long t2 = clock64();
    t_issue += t1 - t0; t_done += t2 - t0;
  }
  if (lane == 0) { out[0] = t_issue / reps; out[1] = t_done / reps; }
}

int main() {
  const int m = 256, n = 281;
  double* A; cudaMalloc(&A, sizeof(double) * m * n * 4); cudaMemset(A, 0, sizeof(double) * m * n * 4);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  using Enc = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  CUtensorMap tm; memset(&tm, 0, sizeof(tm));
  cuuint64_t dims[2] = {m, n}, str[1] = {m * 8}; cuuint32_t box[2] = {16, 16}, es[2] = {1, 1};
  CUresult r = ((Enc)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, A, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode 2d rc=%d\n", (int)r);
  // rank-3 view {16 rows, naug cols, m/16 row blocks} with non-monotonic strides {m*8, 128}
  CUtensorMap tm3; memset(&tm3, 0, sizeof(tm3));
  cuuint64_t d3[3] = {16, n, m / 16}, s3[2] = {m * 8, 128}; cuuint32_t b3[3] = {16, 16, 4}, e3[3] = {1, 1, 1};
  r = ((Enc)fn)(&tm3, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, A, d3, s3, b3, e3, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode 3d (strides m*8, 128) rc=%d\n", (int)r);
  unsigned char* dtm; cudaMalloc(&dtm, 128); cudaMemcpy(dtm, &tm, 128, cudaMemcpyHostToDevice);
  long long* out; cudaMalloc(&out, 16); long long h[2];
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  const char* names[] = {"1 x bulk 2KB", "16 x bulk 512B (16 lanes)", "1 x tile 16x16", "4 x tile 16x16 (4 lanes)", "expect_tx only"};
  for (int mode = 0; mode < 5; mode++)
    for (int pass = 0; pass < 2; pass++) {
      k<<<1, 32, 70000>>>(A, dtm, mode, 64, out);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
      if (pass) printf("%-28s issue %5lld clk, issue->complete %5lld clk  (%s)\n", names[mode], h[0], h[1], cudaGetErrorString(e));
    }
  return 0;
}
