// Microbenchmark: fp64 FMA-pipe vs DMMA (mma.sync.m8n8k4.f64) peak on sm_100a.
// Gives the roofline denominator for the fp64 ULV kernels (MEASURED_PEAKS.json
// only holds bf16 tensor + HBM copy).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double acc[16];
#pragma unroll
  for (int i=0; i<16; i++) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it=0; it<iters; it++) {
#pragma unroll
    for (int i=0; i<16; i++) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i=0; i<16; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template<int NACC> __global__ void dmma_kernel(double* out, int iters, double a, double b) {
  double c[NACC][2];
#pragma unroll
  for (int i=0; i<NACC; i++) { c[i][0] = threadIdx.x * 1e-3 + i; c[i][1] = i; }
  double av = a + threadIdx.x * 1e-9, bv = b;
  for (int it=0; it<iters; it++) {
#pragma unroll
    for (int i=0; i<NACC; i++) dmma(c[i][0], c[i][1], av, bv);
  }
  double s = 0;
#pragma unroll
  for (int i=0; i<NACC; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// smem-fed DMMA: warp tile 32x32 (4x4 mma tiles), A (32 x K) and B (K x 32) in smem
template<int WM, int WN> __global__ void dmma_smem_kernel(double* out, int iters) {
  extern __shared__ double sm[];
  const int K = 64;
  // per-warp private A (WM*8 x K) col-major-ish, B (K x WN*8)
  int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  double* As = sm + warp * (WM*8*K + WN*8*K);
  double* Bs = As + WM*8*K;
  for (int i=lane; i<WM*8*K + WN*8*K; i+=32) As[i] = 1e-3 * (i % 7);
  __syncwarp();
  double c[WM][WN][2];
#pragma unroll
  for (int i=0; i<WM; i++)
#pragma unroll
    for (int j=0; j<WN; j++) { c[i][j][0] = 0; c[i][j][1] = 0; }
  int g = lane / 4, t = lane % 4;
  for (int it=0; it<iters; it++) {
#pragma unroll 4
    for (int k=0; k<K; k+=4) {
      double af[WM], bf[WN];
      // A stored as [k][row] (row contiguous, padded) : a frag = A[row=g][k+t]
#pragma unroll
      for (int i=0; i<WM; i++) af[i] = As[(k+t)*(WM*8) + i*8 + g];
#pragma unroll
      for (int j=0; j<WN; j++) bf[j] = Bs[(k+t)*(WN*8) + j*8 + g];
#pragma unroll
      for (int i=0; i<WM; i++)
#pragma unroll
        for (int j=0; j<WN; j++) dmma(c[i][j][0], c[i][j][1], af[i], bf[j]);
    }
  }
  double s = 0;
#pragma unroll
  for (int i=0; i<WM; i++)
#pragma unroll
    for (int j=0; j<WN; j++) s += c[i][j][0] + c[i][j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template<typename F> float timeit(F f, int reps=5) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r=0; r<reps; r++) {
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("device %s SMs %d clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  int nsm = p.multiProcessorCount;
  double* out; cudaMalloc(&out, sizeof(double) * nsm * 8 * 1024);
  int iters = 20000;
  for (int bps : {1, 2, 4}) for (int th : {128, 256, 512, 1024}) {
    if (bps * th > 2048) continue;
    float ms = timeit([&]{ dfma_kernel<<<nsm*bps, th>>>(out, iters, 1.0000001, 1e-9); });
    double fl = 2.0 * 16 * iters * (double)nsm * bps * th;
    printf("DFMA  blocks/SM %d threads %4d : %8.3f ms  %7.2f TFLOP/s\n", bps, th, ms, fl / ms / 1e9);
  }
  for (int bps : {1, 2, 4}) for (int th : {128, 256, 512}) {
    float ms = timeit([&]{ dmma_kernel<8><<<nsm*bps, th>>>(out, iters, 1.0000001, 1e-9); });
    double fl = 2.0 * 256 * 8 * iters * (double)nsm * bps * (th/32);
    printf("DMMA8 blocks/SM %d threads %4d : %8.3f ms  %7.2f TFLOP/s\n", bps, th, ms, fl / ms / 1e9);
  }
  for (int th : {128, 256}) {
    float ms = timeit([&]{ dmma_kernel<16><<<nsm*2, th>>>(out, iters, 1.0000001, 1e-9); });
    double fl = 2.0 * 256 * 16 * iters * (double)nsm * 2 * (th/32);
    printf("DMMA16 blocks/SM 2 threads %4d : %8.3f ms  %7.2f TFLOP/s\n", th, ms, fl / ms / 1e9);
  }
  {
    int it2 = 2000;
    for (int th : {128, 256}) {
      size_t smem = (th/32) * (4*8*64 + 4*8*64) * sizeof(double);
      cudaFuncSetAttribute(dmma_smem_kernel<4,4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      for (int bps : {1, 2}) {
        float ms = timeit([&]{ dmma_smem_kernel<4,4><<<nsm*bps, th, smem>>>(out, it2); });
        double fl = 2.0 * 256 * 16 * (64/4) * it2 * (double)nsm * bps * (th/32);
        printf("DMMA smem 32x32 warp tile, blocks/SM %d threads %4d : %8.3f ms  %7.2f TFLOP/s\n", bps, th, ms, fl / ms / 1e9);
      }
      size_t smem2 = (th/32) * (2*8*64 + 2*8*64) * sizeof(double);
      for (int bps : {1, 2, 4}) {
        float ms = timeit([&]{ dmma_smem_kernel<2,2><<<nsm*bps, th, smem2>>>(out, it2); });
        double fl = 2.0 * 256 * 4 * (64/4) * it2 * (double)nsm * bps * (th/32);
        printf("DMMA smem 16x16 warp tile, blocks/SM %d threads %4d : %8.3f ms  %7.2f TFLOP/s\n", bps, th, ms, fl / ms / 1e9);
      }
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
