// Isolated timing of the register-resident 8-column Householder sub-panel step
// (same code shape as ulv_qr_kernel<32,true>): clocks per column step.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256, 2) stepk(double* out, long long* t, int mp, int reps) {
  extern __shared__ double sm[];
  const int ldv = 260, LDW = 36;
  double* Vs = sm; double* Ws = Vs + ldv * 32; double* Ts = Ws + LDW * 32; double* zs = Ts + LDW * 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < ldv * 32; i += 256) Vs[i] = 1.0 + ((i * 7919) % 1000) * 1e-3;
  __syncthreads();
  long long t0 = clock64();
  for (int rep = 0; rep < reps; rep++) {
    const int cs = (rep & 3) * 8, sbw = 8;
    if (warp < 4) {
      double a[8][2];
      const int r0 = warp * 64 + lane;
#pragma unroll
      for (int q = 0; q < 8; q++)
#pragma unroll
        for (int rr = 0; rr < 2; rr++) { const int i = r0 + 32 * rr; a[q][rr] = i < mp ? Vs[i + (cs + q) * ldv] : 0.; }
      double tr[8];
#pragma unroll
      for (int q = 0; q < 8; q++) tr[q] = 0.;
      double* pair = zs; double* diag = Ws;
#pragma unroll
      for (int cq = 0; cq < 8; cq++) {
        if (cq < sbw) {
          const int c = cs + cq;
          double p[8];
#pragma unroll
          for (int q = 0; q < 8; q++) p[q] = 0.;
#pragma unroll
          for (int rr = 0; rr < 2; rr++) {
            const int i = r0 + 32 * rr;
            const double xv = i > c ? a[cq][rr] : 0.;
#pragma unroll
            for (int q = 0; q < 8; q++) p[q] += xv * a[q][rr];
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int q = 0; q < 8; q++) p[q] += __shfl_xor_sync(0xffffffffu, p[q], o);
          double* pw = pair + (cq & 1) * 32 + warp * 8;
          if (lane == 0) {
#pragma unroll
            for (int q = 0; q < 8; q++) pw[q] = p[q];
          }
          if (warp == 0 && lane == c) {
#pragma unroll
            for (int q = 0; q < 8; q++) diag[(cq & 1) * 8 + q] = a[q][0];
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          const double* pp = pair + (cq & 1) * 32;
          double sm_[8], dg[8];
#pragma unroll
          for (int q = 0; q < 8; q++) { sm_[q] = (pp[q] + pp[8 + q]) + (pp[16 + q] + pp[24 + q]); dg[q] = diag[(cq & 1) * 8 + q]; }
          const double pn = sm_[cq], alpha = dg[cq];
          double tc = 0., scal = 0., beta = alpha;
          if (pn > 0.) { beta = -copysign(sqrt(alpha * alpha + pn), alpha); const double d = alpha - beta; scal = 1. / d; tc = -d / beta; }
          const bool isdiag = (warp == 0 && lane == c);
#pragma unroll
          for (int q = 0; q < 8; q++) {
            if (q > cq) {
              const double w = tc * (dg[q] + scal * sm_[q]);
              const double ws = w * scal;
#pragma unroll
              for (int rr = 0; rr < 2; rr++) if (r0 + 32 * rr > c) a[q][rr] -= ws * a[cq][rr];
              if (isdiag) a[q][0] -= w;
            }
          }
#pragma unroll
          for (int rr = 0; rr < 2; rr++) if (r0 + 32 * rr > c) a[cq][rr] *= scal;
          if (isdiag) a[cq][0] = beta;
          {
            double val = (lane == cq) ? tc : 0.; double acc = 0.;
#pragma unroll
            for (int b = 0; b < 8; b++) if (b < cq) acc += (b >= lane ? tr[b] : 0.) * (dg[b] + scal * sm_[b]);
            if (lane < cq) val = -tc * acc;
            tr[cq] = val;
          }
        }
      }
#pragma unroll
      for (int q = 0; q < 8; q++)
#pragma unroll
        for (int rr = 0; rr < 2; rr++) { const int i = r0 + 32 * rr; if (i < mp) Vs[i + (cs + q) * ldv] = a[q][rr] * 1e-3 + 1.0; }
      if (warp == 0 && lane < 8) {
#pragma unroll
        for (int q = 0; q < 8; q++) Ts[(cs + lane) + (cs + q) * LDW] = (q >= lane) ? tr[q] : 0.;
      }
    }
    __syncthreads();
  }
  long long t1 = clock64();
  if (tid == 0 && blockIdx.x == 0) t[0] = t1 - t0;
  out[blockIdx.x * 256 + tid] = Vs[tid] + Ts[tid % 64];
}
int main() {
  double* out; long long* t; cudaMalloc(&out, 8 * 256 * 1024); cudaMalloc(&t, 64);
  size_t smem = (260 * 32 + 36 * 32 * 2 + 128) * 8;
  cudaFuncSetAttribute(stepk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int reps = 64;
  for (int grid : {1, 148, 296}) {
    stepk<<<grid, 256, smem>>>(out, t, 256, reps); cudaDeviceSynchronize();
    stepk<<<grid, 256, smem>>>(out, t, 256, reps); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, t, 8, cudaMemcpyDeviceToHost);
    printf("grid %d: %.1f clk per column step (sub-panel of 8: %.0f)  err=%s\n", grid, h / (double)(reps * 8), h / (double)reps, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
