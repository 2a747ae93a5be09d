// Latency microbenchmarks (one warp / one CTA): dependent DFMA, DSQRT, DDIV, 64-bit SHFL
// reduction, LDS->use, bar.sync with 8 warps.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(double* out, long long* t, double seed) {
  __shared__ double sm[1024];
  const int tid = threadIdx.x, lane = tid & 31;
  for (int i = tid; i < 1024; i += blockDim.x) sm[i] = seed + i * 1e-3;
  __syncthreads();
  double a = seed + lane * 1e-6, b = 1.0000001, c = 1e-9;
  long long t0, t1;
  const int N = 256;
  // DFMA chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) a = fma(a, b, c);
  t1 = clock64(); if (tid == 0) t[0] = (t1 - t0);
  // DSQRT chain
  double s = a + 2.0;
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < 64; i++) s = sqrt(s + 1.5);
  t1 = clock64(); if (tid == 0) t[1] = (t1 - t0);
  // DDIV chain
  double d = s + 3.0;
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < 64; i++) d = 1.0 / (d + 0.5);
  t1 = clock64(); if (tid == 0) t[2] = (t1 - t0);
  // warp_sum (5 rounds of 64-bit shuffle + add), chained
  double r = d + lane;
  t0 = clock64();
  for (int i = 0; i < 32; i++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    r *= 1e-3;
  }
  t1 = clock64(); if (tid == 0) t[3] = (t1 - t0);
  // LDS dependent chain (pointer chasing on values)
  int idx = lane;
  double acc = r;
  t0 = clock64();
  for (int i = 0; i < 64; i++) { double v = sm[idx]; acc += v; idx = (idx + 33 + (v > 1e30)) & 1023; }
  t1 = clock64(); if (tid == 0) t[4] = (t1 - t0);
  // bar.sync
  t0 = clock64();
  for (int i = 0; i < 64; i++) __syncthreads();
  t1 = clock64(); if (tid == 0) t[5] = (t1 - t0);
  // DMMA dependent chain
  double c0 = acc, c1 = a;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < 128; i++)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(b), "d"(c));
  t1 = clock64(); if (tid == 0) t[6] = (t1 - t0);
  // rsqrt float seeded + 2 NR fp64
  double q = c0 + 5.0;
  t0 = clock64();
  for (int i = 0; i < 64; i++) {
    double y = (double)rsqrtf((float)q);
    y = y * (1.5 - 0.5 * q * y * y);
    y = y * (1.5 - 0.5 * q * y * y);
    q = q * y + 1.5;   // sqrt(q) + 1.5
  }
  t1 = clock64(); if (tid == 0) t[7] = (t1 - t0);
  out[tid] = a + s + d + r + acc + c0 + c1 + q;
}
int main() {
  double* out; long long* t; cudaMalloc(&out, 8 * 1024); cudaMalloc(&t, 64);
  for (int nt : {32, 256}) {
    lat<<<1, nt>>>(out, t, 1.25); cudaDeviceSynchronize();
    long long h[8]; cudaMemcpy(h, t, 64, cudaMemcpyDeviceToHost);
    printf("threads %d: DFMA %.1f clk/op, DSQRT %.1f, DDIV(rcp) %.1f, warp_sum(5 rounds) %.1f, LDS chain %.1f, bar.sync %.1f, DMMA dep %.1f, rsqrtf+2NR sqrt %.1f\n",
           nt, h[0] / 256.0, h[1] / 64.0, h[2] / 64.0, h[3] / 32.0, h[4] / 64.0, h[5] / 64.0, h[6] / 128.0, h[7] / 64.0);
  }
  return 0;
}
