// strumpack_b200 -- common device/host helpers (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <stdexcept>
#include <string>

namespace sb200 {

#define SB200_CUDA(call)                                                      \
  do {                                                                        \
    cudaError_t e_ = (call);                                                  \
    if (e_ != cudaSuccess)                                                    \
      throw std::runtime_error(std::string("CUDA error: ") +                  \
                               cudaGetErrorString(e_) + " at " + __FILE__ +   \
                               ":" + std::to_string(__LINE__));               \
  } while (0)

// RAII device buffer
template <typename T> struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  explicit DevBuf(size_t n_) { alloc(n_); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
  void alloc(size_t n_) {
    release();
    n = n_;
    if (n) SB200_CUDA(cudaMalloc(&p, n * sizeof(T)));
  }
  void ensure(size_t n_) { if (n_ > n) alloc(n_); }
  void upload(const T* h, size_t cnt, cudaStream_t st = 0) {
    ensure(cnt);
    if (cnt) SB200_CUDA(cudaMemcpyAsync(p, h, cnt * sizeof(T),
                                        cudaMemcpyHostToDevice, st));
  }
};

#ifdef __CUDACC__

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// fp64 tensor-core MMA: D(8x8) += A(8x4, row) * B(4x8, col).
// Fragment ownership (lane = 4*g + t, g = lane/4, t = lane%4):
//   a = A[g][t], b = B[t][g], c0 = C[g][2t], c1 = C[g][2t+1].
// On B200 this pipe peaks at 37.1 TFLOP/s (profiles/microbench), the plain
// DFMA pipe at 33.9; tcgen05 has no f64 kind, so this IS the fp64 tensor path.
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile(
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, "
      "{%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

// 8-byte asynchronous global -> shared copy (LDGSTS); !pred zero-fills the
// destination without touching global memory (src-size 0).
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc, bool pred) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(sa), "l"(gsrc),
               "r"(pred ? 8 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}
// software pipelines: close a group of copies / wait until at most N groups are pending
__device__ __forceinline__ void cp_async_commit() {
  asm volatile("cp.async.commit_group;\n" ::: "memory");
}
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// pull one 128-byte line into L2 ahead of a dependent load
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p));
}

// Leading dimension for fp64 tiles in shared memory: ld % 16 == 4 makes both
// "k along rows" and "k along columns" DMMA fragment loads conflict-free per
// half-warp (an LDS.64 is two 128-byte wavefronts at best).
__host__ __device__ __forceinline__ int smem_ld(int rows) {
  int ld = (rows + 15) / 16 * 16 + 4;
  return ld;
}

// C(MxN, smem, ldc) = beta*C + alpha * op(A) * op(B), all operands in shared
// memory, column-major.  TA: A is given as K x M (use A^T); TB: B is N x K.
// Work is split over the calling warps in 16x16 output blocks (2x2 DMMA
// tiles); edges are zero-padded by predication, so any M, N, K is legal.
// All `nwarps` warps of the CTA must call it; no __syncthreads inside.
template <bool TA, bool TB>
__device__ __forceinline__ void smem_gemm(int M, int N, int K, double alpha,
                                          const double* __restrict__ A, int lda,
                                          const double* __restrict__ B, int ldb,
                                          double beta, double* __restrict__ C,
                                          int ldc, int warp, int nwarps,
                                          int lane) {
  const int g = lane >> 2, t = lane & 3;
  const int tm = (M + 15) >> 4, tn = (N + 15) >> 4;
  for (int tile = warp; tile < tm * tn; tile += nwarps) {
    const int i0 = (tile % tm) << 4, j0 = (tile / tm) << 4;
    double c[2][2][2] = {};
    const int ia0 = i0 + g, ia1 = i0 + 8 + g;
    const int jb0 = j0 + g, jb1 = j0 + 8 + g;
    for (int k0 = 0; k0 < K; k0 += 4) {
      const int k = k0 + t;
      const bool kin = k < K;
      double a0, a1, b0, b1;
      if (TA) {
        a0 = (kin && ia0 < M) ? A[k + (size_t)ia0 * lda] : 0.;
        a1 = (kin && ia1 < M) ? A[k + (size_t)ia1 * lda] : 0.;
      } else {
        a0 = (kin && ia0 < M) ? A[ia0 + (size_t)k * lda] : 0.;
        a1 = (kin && ia1 < M) ? A[ia1 + (size_t)k * lda] : 0.;
      }
      if (TB) {
        b0 = (kin && jb0 < N) ? B[jb0 + (size_t)k * ldb] : 0.;
        b1 = (kin && jb1 < N) ? B[jb1 + (size_t)k * ldb] : 0.;
      } else {
        b0 = (kin && jb0 < N) ? B[k + (size_t)jb0 * ldb] : 0.;
        b1 = (kin && jb1 < N) ? B[k + (size_t)jb1 * ldb] : 0.;
      }
      dmma(c[0][0][0], c[0][0][1], a0, b0);
      dmma(c[0][1][0], c[0][1][1], a0, b1);
      dmma(c[1][0][0], c[1][0][1], a1, b0);
      dmma(c[1][1][0], c[1][1][1], a1, b1);
    }
#pragma unroll
    for (int ti = 0; ti < 2; ti++)
#pragma unroll
      for (int tj = 0; tj < 2; tj++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int i = i0 + ti * 8 + g, j = j0 + tj * 8 + 2 * t + e;
          if (i < M && j < N) {
            double* p = C + i + (size_t)j * ldc;
            *p = (beta == 0. ? 0. : beta * *p) + alpha * c[ti][tj][e];
          }
        }
  }
}

#endif  // __CUDACC__

}  // namespace sb200
