// strumpack_b200 -- batched column-pivoted Gram-Schmidt QR with the reference's
// truncation rule (xGEQP3TOL, reference src/dense/lapack/dgeqp3tol.f:203-209:
// stop at the first |R_jj| with |R_jj|/|R_00| <= rtol or |R_jj| <= atol).
// Used for the interpolative decompositions of the HSS construction
// (DenseMatrix::ID_column_GEQP3, reference src/dense/DenseMatrix.cpp:764-790)
// and for the BLR tile compression (DenseMatrix::low_rank, :792-810).
#pragma once
#include "sb200_common.cuh"

namespace sb200 {
namespace {

constexpr int kCpqrThreads = 1024;   // one CTA per matrix: the steps are latency-bound, more warps = more columns in flight
constexpr int kCpqrWarps = kCpqrThreads / 32;

// ---- batched column-pivoted QR -> interpolative decomposition ----------------
struct IDTask {
  double* M;       // ns x nc, column-major, ld = ns (destroyed)
  double* R;       // rcap x nc workspace, ld = rcap
  int ns, nc, rcap;
  int* order;      // nc: pivot order (output): order[0:rank] = skeleton columns
  int* rank;       // 1
  double* E;       // (nc - rank) x rank column-major (output), capacity (nc x rcap);
                   // nullptr: keep R intact (low-rank factorisation Q R, no ID)
  int strict;      // 1: report rank -1 when the tolerance is not met within rcap
};

__global__ void __launch_bounds__(kCpqrThreads)
id_cpqr_kernel(const IDTask* __restrict__ tasks, double rtol, double atol,
               int max_rank) {
  const IDTask t = tasks[blockIdx.x];
  extern __shared__ double sm[];
  double* nrm2 = sm;                 // nc
  int* ord = (int*)(nrm2 + t.nc);    // nc
  __shared__ double redv[kCpqrWarps];
  __shared__ int redi[kCpqrWarps];
  __shared__ int s_piv;
  __shared__ double s_r00;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ns = t.ns, nc = t.nc;
  for (int c = tid; c < nc; c += kCpqrThreads) ord[c] = c;
  for (int c = warp; c < nc; c += kCpqrWarps) {
    const double* col = t.M + (size_t)c * ns;
    double a = 0.;
    for (int i = lane; i < ns; i += 32) a += col[i] * col[i];
    a = warp_sum(a);
    if (lane == 0) nrm2[c] = a;
  }
  __syncthreads();
  const int rmax = min(min(ns, nc), min(t.rcap, max_rank));
  int rank = 0;
  for (int j = 0; j < rmax; j++) {
    // pivot = remaining column (position >= j) with the largest norm
    double best = -1.;
    int bp = j;
    for (int p = j + tid; p < nc; p += kCpqrThreads) {
      double v = nrm2[ord[p]];
      if (v > best) { best = v; bp = p; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      double ov = __shfl_xor_sync(0xffffffffu, best, o);
      int op = __shfl_xor_sync(0xffffffffu, bp, o);
      if (ov > best || (ov == best && op < bp)) { best = ov; bp = op; }
    }
    if (lane == 0) { redv[warp] = best; redi[warp] = bp; }
    __syncthreads();
    if (tid == 0) {
      double b = redv[0]; int p = redi[0];
      for (int w = 1; w < kCpqrWarps; w++)
        if (redv[w] > b || (redv[w] == b && redi[w] < p)) { b = redv[w]; p = redi[w]; }
      int tmp = ord[j]; ord[j] = ord[p]; ord[p] = tmp;
      s_piv = ord[j];
      if (j == 0) s_r00 = sqrt(fmax(b, 0.));
    }
    __syncthreads();
    const int pc = s_piv;
    const double rjj = sqrt(fmax(nrm2[pc], 0.));
    // stopping rule of xGEQP3TOL: the new diagonal entry is tested first
    if (rjj / s_r00 <= rtol || rjj <= atol || !(rjj > 0.)) break;
    rank = j + 1;
    double* q = t.M + (size_t)pc * ns;
    const double inv = 1. / rjj;
    for (int i = tid; i < ns; i += kCpqrThreads) q[i] *= inv;
    if (tid == 0) t.R[j + (size_t)pc * t.rcap] = rjj;
    __syncthreads();
    // orthogonalise the remaining columns against q, refresh their norms.  The
    // step is bound by the latency of its global / L2 accesses, not by flops: a
    // warp takes kCG columns at a time so that kCG independent loads per lane are
    // in flight (q is read once for all of them); the sums of every column are
    // formed in the same order as one column at a time.
    constexpr int kCG = 4;
    for (int p0 = j + 1 + warp * kCG; p0 < nc; p0 += kCpqrWarps * kCG) {
      double* col[kCG];
      bool in[kCG];
      int cidx[kCG];
#pragma unroll
      for (int u = 0; u < kCG; u++) {
        in[u] = p0 + u < nc;
        cidx[u] = ord[in[u] ? p0 + u : j];
        col[u] = t.M + (size_t)cidx[u] * ns;
      }
      double r[kCG];
#pragma unroll
      for (int u = 0; u < kCG; u++) r[u] = 0.;
      for (int i = lane; i < ns; i += 32) {
        const double qi = q[i];
#pragma unroll
        for (int u = 0; u < kCG; u++) r[u] += qi * col[u][i];
      }
#pragma unroll
      for (int u = 0; u < kCG; u++) r[u] = warp_sum(r[u]);
      double a[kCG];
#pragma unroll
      for (int u = 0; u < kCG; u++) a[u] = 0.;
      for (int i = lane; i < ns; i += 32) {
        const double qi = q[i];
#pragma unroll
        for (int u = 0; u < kCG; u++)
          if (in[u]) {
            const double v = col[u][i] - r[u] * qi;
            col[u][i] = v;
            a[u] += v * v;
          }
      }
#pragma unroll
      for (int u = 0; u < kCG; u++) a[u] = warp_sum(a[u]);
      if (lane == 0) {
#pragma unroll
        for (int u = 0; u < kCG; u++)
          if (in[u]) { t.R[j + (size_t)cidx[u] * t.rcap] = r[u]; nrm2[cidx[u]] = a[u]; }
      }
    }
    __syncthreads();
  }
  __syncthreads();
  if (t.strict && rank == rmax && rmax < min(ns, nc)) {
    // not converged within the rank cap: is the next pivot still above the tolerance?
    double best = 0.;
    for (int p = rank + tid; p < nc; p += kCpqrThreads) best = fmax(best, nrm2[ord[p]]);
    for (int o = 16; o > 0; o >>= 1) best = fmax(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (lane == 0) redv[warp] = best;
    __syncthreads();
    if (tid == 0) {
      double b = 0.;
      for (int w = 0; w < kCpqrWarps; w++) b = fmax(b, redv[w]);
      const double rn = sqrt(b);
      if (!(rn / s_r00 <= rtol || rn <= atol)) s_piv = -1; else s_piv = 0;
    }
    __syncthreads();
    if (s_piv < 0) rank = -1;
  }
  if (tid == 0) *t.rank = rank;
  for (int c = tid; c < nc; c += kCpqrThreads) t.order[c] = ord[c];
  if (t.E == nullptr || rank < 0) return;
  // E^T = R11^{-1} R12 : back substitution, one thread per remaining column
  const int k = nc - rank;
  for (int p = tid; p < k; p += kCpqrThreads) {
    const int c = ord[rank + p];
    double* x = t.R + (size_t)c * t.rcap;   // in place in column c of R
    for (int a = rank - 1; a >= 0; a--) {
      double v = x[a];
      for (int b = a + 1; b < rank; b++) v -= t.R[a + (size_t)ord[b] * t.rcap] * x[b];
      x[a] = v / t.R[a + (size_t)ord[a] * t.rcap];
    }
    for (int a = 0; a < rank; a++) t.E[p + (size_t)a * k] = x[a];
  }
}


}  // namespace
}  // namespace sb200
