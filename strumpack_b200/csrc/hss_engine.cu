// strumpack_b200 -- HSS apply / ULV factor / ULV solve on sm_100a.
// See hss_engine.hpp for the reference routines each kernel family replaces.
#include "hss_engine.hpp"
#include "ulv_qr3.cuh"

#include <cuda.h>   // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#include <dlfcn.h>
#include <nccl.h>   // types only: the library is bound at run time (dlopen), single-GPU use needs no NCCL

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace sb200 {

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr size_t kMaxSmem = 200 * 1024;   // dynamic shared memory we ask for at most

// ===========================================================================
//                                  APPLY
// ===========================================================================
// Workspace convention: node block = base + w_off * s, column c of a block
// with `len` rows starts at + c * len.

// Up-sweep, one height class (root excluded):  t1 = top(P^T in) + E^H bot(P^T in)
//   reference apply_fwd   HSSMatrix.apply.hpp:55-82 (leaf :58-63, inner :76-80)
//   HSSBasisID::applyC    HSSBasisID.hpp:189-203
template <bool TRANS>
__global__ void __launch_bounds__(kThreads)
hss_up_kernel(const DNode* __restrict__ nodes, const int* __restrict__ list,
              const double* __restrict__ vals, const int* __restrict__ perms,
              const double* __restrict__ x, int ldx, double* __restrict__ t1,
              int s) {
  extern __shared__ __align__(16) double sm[];
  const DNode nd = nodes[list[blockIdx.x]];
  const int c = blockIdx.y, tid = threadIdx.x;
  const int r = TRANS ? nd.u_rank : nd.v_rank;
  const int n = TRANS ? nd.u_rows : nd.v_rows;
  if (r == 0) return;
  const int* P = perms + (TRANS ? nd.Pu : nd.Pv);
  double* pb = sm;
  if (nd.leaf) {
    const double* xin = x + (TRANS ? nd.row_off : nd.col_off) + (size_t)c * ldx;
    for (int i = tid; i < n; i += kThreads) pb[i] = xin[P[i]];
  } else {
    const DNode c0 = nodes[nd.ch0], c1 = nodes[nd.ch1];
    const int q0 = TRANS ? c0.u_rank : c0.v_rank;
    const int q1 = TRANS ? c1.u_rank : c1.v_rank;
    const double* a = t1 + (size_t)c0.w_off * s + (size_t)c * q0;
    const double* b = t1 + (size_t)c1.w_off * s + (size_t)c * q1;
    for (int i = tid; i < n; i += kThreads) {
      int p = P[i];
      pb[i] = p < q0 ? a[p] : b[p - q0];
    }
  }
  __syncthreads();
  const int k = n - r;
  const double* E = vals + (TRANS ? nd.Eu : nd.Ev);
  double* out = t1 + (size_t)nd.w_off * s + (size_t)c * r;
  const int warp = tid >> 5, lane = tid & 31;
  for (int j = warp; j < r; j += kWarps) {
    double acc = 0.;
    const double* Ej = E + (size_t)j * k;
    for (int i = lane; i < k; i += 32) acc += Ej[i] * pb[r + i];
    acc = warp_sum(acc);
    if (lane == 0) out[j] = pb[j] + acc;
  }
}

// Down-sweep over inner nodes of one height class (root included):
//   u = U t2 (= P [t2; E t2]) split over the children, plus B01/B10 coupling.
//   reference apply_bwd   HSSMatrix.apply.hpp:101-124, HSSBasisID::apply :155-169
template <bool TRANS>
__global__ void __launch_bounds__(kThreads)
hss_down_kernel(const DNode* __restrict__ nodes, const int* __restrict__ list,
                const double* __restrict__ vals, const int* __restrict__ perms,
                const double* __restrict__ t1, double* __restrict__ t2, int s) {
  extern __shared__ __align__(16) double sm[];
  const int id = list[blockIdx.x];
  const DNode nd = nodes[id];
  if (nd.leaf) return;
  const int c = blockIdx.y, tid = threadIdx.x;
  const DNode c0 = nodes[nd.ch0], c1 = nodes[nd.ch1];
  // ranks of the children on the way down (o*) and on the way up (q*)
  const int o0 = TRANS ? c0.v_rank : c0.u_rank, o1 = TRANS ? c1.v_rank : c1.u_rank;
  const int q0 = TRANS ? c0.u_rank : c0.v_rank, q1 = TRANS ? c1.u_rank : c1.v_rank;
  const int rout = TRANS ? nd.v_rank : nd.u_rank;
  const int nout = o0 + o1;
  double* u = sm;            // nout
  double* tin = sm + nout;   // rout : this node's t2
  // Latency-bound (the upper classes have a handful of nodes, their generators
  // come from DRAM every step): one warp per output row, the lanes split the
  // dot product, so that all loads of a stage are in flight at once -- three
  // dependent memory round trips per node instead of one per few terms.
  const int warp = tid >> 5, lane = tid & 31;
  const bool has_u = (nd.parent >= 0) && rout > 0;
  if (has_u) {
    const double* my = t2 + (size_t)nd.w_off * s + (size_t)c * rout;
    for (int i = tid; i < rout; i += kThreads) tin[i] = my[i];
    __syncthreads();
    const int* P = perms + (TRANS ? nd.Pv : nd.Pu);
    const double* E = vals + (TRANS ? nd.Ev : nd.Eu);
    const int k = nout - rout;
    for (int i = warp; i < nout; i += kWarps) {
      double v = 0.;
      if (i < rout) v = lane == 0 ? tin[i] : 0.;
      else
        for (int l = lane; l < rout; l += 32) v += E[(i - rout) + (size_t)l * k] * tin[l];
      v = warp_sum(v);
      if (lane == 0) u[P[i]] = v;
    }
  } else {
    for (int i = tid; i < nout; i += kThreads) u[i] = 0.;
  }
  __syncthreads();
  const double* ta = t1 + (size_t)c0.w_off * s + (size_t)c * q0;  // t1(c0)
  const double* tb = t1 + (size_t)c1.w_off * s + (size_t)c * q1;  // t1(c1)
  double* oa = t2 + (size_t)c0.w_off * s + (size_t)c * o0;
  double* ob = t2 + (size_t)c1.w_off * s + (size_t)c * o1;
  const double* B01 = vals + nd.B01;  // u_rank(c0) x v_rank(c1)
  const double* B10 = vals + nd.B10;  // u_rank(c1) x v_rank(c0)
  for (int i = warp; i < nout; i += kWarps) {
    double v = 0.;
    const bool first = i < o0;
    const int ii = first ? i : i - o0;
    if (!TRANS) {
      if (first) { for (int j = lane; j < q1; j += 32) v += B01[ii + (size_t)j * o0] * tb[j]; }
      else { for (int j = lane; j < q0; j += 32) v += B10[ii + (size_t)j * o1] * ta[j]; }
    } else {
      // t2(c0) += B10^H t1(c1) ; t2(c1) += B01^H t1(c0)   (apply.hpp:194-211)
      if (first) { for (int j = lane; j < q1; j += 32) v += B10[j + (size_t)ii * q1] * tb[j]; }
      else { for (int j = lane; j < q0; j += 32) v += B01[j + (size_t)ii * q0] * ta[j]; }
    }
    v = warp_sum(v);
    if (lane == 0) (first ? oa : ob)[ii] = v + u[i];
  }
}

// Leaves:  y = D x + U t2    (apply_bwd leaf, HSSMatrix.apply.hpp:87-99)
template <bool TRANS>
__global__ void __launch_bounds__(kThreads)
hss_leaf_kernel(const DNode* __restrict__ nodes, const int* __restrict__ list,
                const double* __restrict__ vals, const int* __restrict__ perms,
                const double* __restrict__ x, int ldx, const double* __restrict__ t2,
                double* __restrict__ y, int ldy, int s, double beta) {
  extern __shared__ __align__(16) double sm[];
  const DNode nd = nodes[list[blockIdx.x]];
  const int c = blockIdx.y, tid = threadIdx.x;
  const int nin = TRANS ? nd.rows : nd.cols;    // length of the x slice
  const int nout = TRANS ? nd.cols : nd.rows;   // length of the y slice
  const int rout = TRANS ? nd.v_rank : nd.u_rank;
  double* xs = sm;           // nin
  double* u = sm + nin;      // nout
  double* tin = u + nout;    // rout
  const double* xin = x + (TRANS ? nd.row_off : nd.col_off) + (size_t)c * ldx;
  for (int i = tid; i < nin; i += kThreads) xs[i] = xin[i];
  const bool has_u = (nd.parent >= 0) && rout > 0;
  if (has_u) {
    const double* my = t2 + (size_t)nd.w_off * s + (size_t)c * rout;
    for (int i = tid; i < rout; i += kThreads) tin[i] = my[i];
    __syncthreads();
    const int* P = perms + (TRANS ? nd.Pv : nd.Pu);
    const double* E = vals + (TRANS ? nd.Ev : nd.Eu);
    const int k = nout - rout;
    for (int i = tid; i < nout; i += kThreads) {
      double v;
      if (i < rout) v = tin[i];
      else {
        v = 0.;
        for (int l = 0; l < rout; l++) v += E[(i - rout) + (size_t)l * k] * tin[l];
      }
      u[P[i]] = v;
    }
  } else {
    for (int i = tid; i < nout; i += kThreads) u[i] = 0.;
  }
  __syncthreads();
  const double* D = vals + nd.D;  // rows x cols, ld = rows
  double* yo = y + (TRANS ? nd.col_off : nd.row_off) + (size_t)c * ldy;
  if (!TRANS) {
    const int m = nd.rows;
    for (int i = tid; i < m; i += kThreads) {
      double a0 = u[i], a1 = 0., a2 = 0., a3 = 0.;
      const double* Di = D + i;
      int j = 0;
      for (; j + 4 <= nin; j += 4) {
        a0 += Di[(size_t)j * m] * xs[j];
        a1 += Di[(size_t)(j + 1) * m] * xs[j + 1];
        a2 += Di[(size_t)(j + 2) * m] * xs[j + 2];
        a3 += Di[(size_t)(j + 3) * m] * xs[j + 3];
      }
      for (; j < nin; j++) a0 += Di[(size_t)j * m] * xs[j];
      const double v = (a0 + a1) + (a2 + a3);
      yo[i] = beta == 0. ? v : v + beta * yo[i];   // apply_bwd: C = beta C + ... (apply.hpp:87-99)
    }
  } else {
    const int m = nd.rows;
    const int warp = tid >> 5, lane = tid & 31;
    for (int j = warp; j < nout; j += kWarps) {
      const double* Dj = D + (size_t)j * m;
      double acc = 0.;
      for (int i = lane; i < m; i += 32) acc += Dj[i] * xs[i];
      acc = warp_sum(acc);
      if (lane == 0) yo[j] = beta == 0. ? u[j] + acc : u[j] + acc + beta * yo[j];
    }
  }
}

// ---------------------------------------------------------------------------
// Multi-rhs (GEMM-shaped) apply: one CTA per (node, tile of kMS right-hand
// sides).  The generators are read once per tile instead of once per column,
// and the two large products (E^H * bot in the up-sweep, D * x at the leaves)
// run on the fp64 tensor pipe (smem_gemm: A fragments straight from the
// generator arena, B = the kMS vectors in shared memory).  Same math and
// workspace layout as the single-column kernels above.
// ---------------------------------------------------------------------------
constexpr int kMS = 16;

template <bool TRANS>
__global__ void __launch_bounds__(kThreads)
hss_up_mm_kernel(const DNode* __restrict__ nodes, const int* __restrict__ list,
                 const double* __restrict__ vals, const int* __restrict__ perms,
                 const double* __restrict__ x, int ldx, double* t1, int s, int ldp) {
  extern __shared__ __align__(16) double sm[];
  const DNode nd = nodes[list[blockIdx.x]];
  const int c0 = blockIdx.y * kMS, ns = min(kMS, s - c0), tid = threadIdx.x;
  const int r = TRANS ? nd.u_rank : nd.v_rank;
  const int n = TRANS ? nd.u_rows : nd.v_rows;
  if (r == 0) return;
  const int* P = perms + (TRANS ? nd.Pu : nd.Pv);
  double* pb = sm;   // n x ns, ld = ldp
  if (nd.leaf) {
    const double* xin = x + (TRANS ? nd.row_off : nd.col_off) + (size_t)c0 * ldx;
    for (int idx = tid; idx < n * ns; idx += kThreads) {
      const int i = idx % n, c = idx / n;
      pb[i + c * ldp] = xin[P[i] + (size_t)c * ldx];
    }
  } else {
    const DNode ch0 = nodes[nd.ch0], ch1 = nodes[nd.ch1];
    const int q0 = TRANS ? ch0.u_rank : ch0.v_rank;
    const int q1 = TRANS ? ch1.u_rank : ch1.v_rank;
    const double* a = t1 + (size_t)ch0.w_off * s + (size_t)c0 * q0;
    const double* b = t1 + (size_t)ch1.w_off * s + (size_t)c0 * q1;
    for (int idx = tid; idx < n * ns; idx += kThreads) {
      const int i = idx % n, c = idx / n, p = P[i];
      pb[i + c * ldp] = p < q0 ? a[p + (size_t)c * q0] : b[(p - q0) + (size_t)c * q1];
    }
  }
  __syncthreads();
  double* out = t1 + (size_t)nd.w_off * s + (size_t)c0 * r;   // r x ns, ld = r
  for (int idx = tid; idx < r * ns; idx += kThreads) out[idx] = pb[(idx % r) + (idx / r) * ldp];
  __syncthreads();
  const int k = n - r;
  if (k > 0)   // out += E^H * bot      (E is k x r)
    smem_gemm<true, false>(r, ns, k, 1., vals + (TRANS ? nd.Eu : nd.Ev), k, pb + r, ldp, 1.,
                           out, r, tid >> 5, kWarps, tid & 31);
}

// u(P[i], :) = [t2; E t2](i, :) for the kMS columns of this tile; `dst` is nout x ns
// (ld = ldd).  One thread per row, the columns in registers: E is read once.
template <bool TRANS>
__device__ __forceinline__ void basis_expand_mm(const DNode& nd, const double* __restrict__ vals,
                                                const int* __restrict__ perms,
                                                const double* tin, int ldt, int nout, int rout,
                                                int ns, double* dst, int ldd, int tid) {
  const int* P = perms + (TRANS ? nd.Pv : nd.Pu);
  const double* E = vals + (TRANS ? nd.Ev : nd.Eu);
  const int k = nout - rout;
  for (int i = tid; i < nout; i += kThreads) {
    double acc[kMS];
#pragma unroll
    for (int c = 0; c < kMS; c++) acc[c] = 0.;
    if (i < rout) {
#pragma unroll
      for (int c = 0; c < kMS; c++) if (c < ns) acc[c] = tin[i + c * ldt];
    } else {
      for (int l = 0; l < rout; l++) {
        const double e = E[(i - rout) + (size_t)l * k];
#pragma unroll
        for (int c = 0; c < kMS; c++) acc[c] += e * tin[l + c * ldt];
      }
    }
    const int p = P[i];
#pragma unroll
    for (int c = 0; c < kMS; c++) if (c < ns) dst[p + c * ldd] = acc[c];
  }
}

template <bool TRANS>
__global__ void __launch_bounds__(kThreads)
hss_down_mm_kernel(const DNode* __restrict__ nodes, const int* __restrict__ list,
                   const double* __restrict__ vals, const int* __restrict__ perms,
                   const double* __restrict__ t1, double* t2, int s, int ldp) {
  extern __shared__ __align__(16) double sm[];
  const int id = list[blockIdx.x];
  const DNode nd = nodes[id];
  if (nd.leaf) return;
  const int c0 = blockIdx.y * kMS, ns = min(kMS, s - c0), tid = threadIdx.x;
  const DNode ch0 = nodes[nd.ch0], ch1 = nodes[nd.ch1];
  const int o0 = TRANS ? ch0.v_rank : ch0.u_rank, o1 = TRANS ? ch1.v_rank : ch1.u_rank;
  const int q0 = TRANS ? ch0.u_rank : ch0.v_rank, q1 = TRANS ? ch1.u_rank : ch1.v_rank;
  const int rout = TRANS ? nd.v_rank : nd.u_rank;
  const int nout = o0 + o1;
  double* u = sm;                 // nout x kMS, ld = ldp
  double* tin = u + ldp * kMS;    // rout x kMS, ld = ldp
  double* ta = tin + ldp * kMS;   // t1(c0): q0 x kMS, ld = ldp
  double* tb = ta + ldp * kMS;    // t1(c1): q1 x kMS, ld = ldp
  const bool has_u = (nd.parent >= 0) && rout > 0;
  for (int idx = tid; idx < ldp * kMS * 4; idx += kThreads) sm[idx] = 0.;
  __syncthreads();
  if (has_u) {
    const double* my = t2 + (size_t)nd.w_off * s + (size_t)c0 * rout;
    for (int idx = tid; idx < rout * ns; idx += kThreads) tin[(idx % rout) + (idx / rout) * ldp] = my[idx];
  }
  {
    const double* a = t1 + (size_t)ch0.w_off * s + (size_t)c0 * q0;
    const double* b = t1 + (size_t)ch1.w_off * s + (size_t)c0 * q1;
    for (int idx = tid; idx < q0 * ns; idx += kThreads) ta[(idx % q0) + (idx / q0) * ldp] = a[idx];
    for (int idx = tid; idx < q1 * ns; idx += kThreads) tb[(idx % q1) + (idx / q1) * ldp] = b[idx];
  }
  __syncthreads();
  if (has_u) basis_expand_mm<TRANS>(nd, vals, perms, tin, ldp, nout, rout, ns, u, ldp, tid);
  __syncthreads();
  double* oa = t2 + (size_t)ch0.w_off * s + (size_t)c0 * o0;
  double* ob = t2 + (size_t)ch1.w_off * s + (size_t)c0 * o1;
  const double* B01 = vals + nd.B01;  // u_rank(c0) x v_rank(c1)
  const double* B10 = vals + nd.B10;  // u_rank(c1) x v_rank(c0)
  for (int i = tid; i < nout; i += kThreads) {
    double acc[kMS];
#pragma unroll
    for (int c = 0; c < kMS; c++) acc[c] = u[i + c * ldp];
    const bool first = i < o0;
    const int ii = first ? i : i - o0;
    const int nq = first ? q1 : q0;
    const double* tsrc = first ? tb : ta;
    for (int j = 0; j < nq; j++) {
      double bv;
      if (!TRANS) bv = first ? B01[ii + (size_t)j * o0] : B10[ii + (size_t)j * o1];
      else        bv = first ? B10[j + (size_t)ii * q1] : B01[j + (size_t)ii * q0];
#pragma unroll
      for (int c = 0; c < kMS; c++) acc[c] += bv * tsrc[j + c * ldp];
    }
    double* o = first ? oa : ob;
    const int lo = first ? o0 : o1;
#pragma unroll
    for (int c = 0; c < kMS; c++) if (c < ns) o[ii + (size_t)c * lo] = acc[c];
  }
}

template <bool TRANS>
__global__ void __launch_bounds__(kThreads)
hss_leaf_mm_kernel(const DNode* __restrict__ nodes, const int* __restrict__ list,
                   const double* __restrict__ vals, const int* __restrict__ perms,
                   const double* __restrict__ x, int ldx, const double* __restrict__ t2,
                   double* __restrict__ y, int ldy, int s, int ldp, double beta) {
  extern __shared__ __align__(16) double sm[];
  const DNode nd = nodes[list[blockIdx.x]];
  const int c0 = blockIdx.y * kMS, ns = min(kMS, s - c0), tid = threadIdx.x;
  const int nin = TRANS ? nd.rows : nd.cols;
  const int nout = TRANS ? nd.cols : nd.rows;
  const int rout = TRANS ? nd.v_rank : nd.u_rank;
  double* xs = sm;                 // nin x kMS, ld = ldp
  double* ys = xs + ldp * kMS;     // nout x kMS, ld = ldp
  double* tin = ys + ldp * kMS;    // rout x kMS, ld = ldp
  for (int idx = tid; idx < ldp * kMS * 3; idx += kThreads) sm[idx] = 0.;
  __syncthreads();
  const double* xin = x + (TRANS ? nd.row_off : nd.col_off) + (size_t)c0 * ldx;
  for (int idx = tid; idx < nin * ns; idx += kThreads) {
    const int i = idx % nin, c = idx / nin;
    xs[i + c * ldp] = xin[i + (size_t)c * ldx];
  }
  const bool has_u = (nd.parent >= 0) && rout > 0;
  if (has_u) {
    const double* my = t2 + (size_t)nd.w_off * s + (size_t)c0 * rout;
    for (int idx = tid; idx < rout * ns; idx += kThreads) tin[(idx % rout) + (idx / rout) * ldp] = my[idx];
  }
  __syncthreads();
  if (has_u) basis_expand_mm<TRANS>(nd, vals, perms, tin, ldp, nout, rout, ns, ys, ldp, tid);
  __syncthreads();
  // ys += op(D) xs      (D is rows x cols, ld = rows)
  smem_gemm<TRANS, false>(nout, ns, nin, 1., vals + nd.D, nd.rows, xs, ldp, 1., ys, ldp,
                          tid >> 5, kWarps, tid & 31);
  __syncthreads();
  double* yo = y + (TRANS ? nd.col_off : nd.row_off) + (size_t)c0 * ldy;
  for (int idx = tid; idx < nout * ns; idx += kThreads) {
    const int i = idx % nout, c = idx / nout;
    double* o = yo + i + (size_t)c * ldy;
    *o = beta == 0. ? ys[i + c * ldp] : ys[i + c * ldp] + beta * *o;
  }
}

__global__ void hss_shift_kernel(const DNode* __restrict__ nodes,
                                 const int* __restrict__ list, double* vals,
                                 double sigma) {
  const DNode nd = nodes[list[blockIdx.x]];
  if (!nd.leaf) return;
  double* D = vals + nd.D;
  const int n = min(nd.rows, nd.cols);
  for (int i = threadIdx.x; i < n; i += blockDim.x) D[i + (size_t)i * nd.rows] += sigma;
}

// ===========================================================================
//                         ELEMENT EXTRACTION
// ===========================================================================
// H(I, J) for arbitrary index sets  (reference HSSMatrix::extract / extract_add /
// get, HSSMatrix.extract.hpp:8-188, used by the sparse front assembly and
// checked by test/test_HSS_seq.cpp:204-233).  The reference walks the tree
// with per-node index sub-lists; here
//  1. one CTA per requested row (column) index climbs from its leaf to the
//     child of the root and records, for every node on the path, the row of
//     the node's nested basis  U_big(node)[i, :]  (V_big(node)[j, :]);
//  2. one thread per requested entry finds the level where the two paths part
//     and evaluates  u^T B v  with the coupling block of their common parent
//     (or reads D when both indices sit in the same leaf).
// paths: for every index `maxd` node ids, top-down (depth 1 first), -1 padded.
template <bool VSIDE>
__global__ void __launch_bounds__(128)
hss_basis_rows_kernel(const DNode* __restrict__ nodes, const double* __restrict__ vals,
                      const int* __restrict__ perms, const int* __restrict__ idx,
                      const int* __restrict__ paths, int maxd, int maxr,
                      double* __restrict__ chain) {
  extern __shared__ __align__(16) double sm[];
  double* ua = sm;                  // maxr
  double* ub = sm + maxr;           // maxr
  int* pos = reinterpret_cast<int*>(ub + maxr);   // maxr
  const int q = blockIdx.x, tid = threadIdx.x;
  const int* path = paths + (size_t)q * maxd;
  int dl = maxd - 1;
  while (dl >= 0 && path[dl] < 0) dl--;
  if (dl < 0) return;               // single-node tree: no basis
  double* out = chain + (size_t)q * maxd * maxr;
  // ---- leaf
  {
    const DNode nd = nodes[path[dl]];
    const int li = idx[q] - (VSIDE ? nd.col_off : nd.row_off);
    const int n = VSIDE ? nd.v_rows : nd.u_rows, r = VSIDE ? nd.v_rank : nd.u_rank;
    const int* P = perms + (VSIDE ? nd.Pv : nd.Pu);
    const double* E = vals + (VSIDE ? nd.Ev : nd.Eu);
    for (int j = tid; j < n; j += 128) if (P[j] == li) pos[0] = j;
    __syncthreads();
    const int p = pos[0], k = n - r;
    for (int c = tid; c < r; c += 128) ua[c] = p < r ? (p == c ? 1. : 0.) : E[(p - r) + (size_t)c * k];
    __syncthreads();
    for (int c = tid; c < r; c += 128) out[(size_t)dl * maxr + c] = ua[c];
  }
  // ---- ancestors
  for (int d = dl - 1; d >= 0; d--) {
    const DNode nd = nodes[path[d]];
    const int child = path[d + 1];
    const DNode c0 = nodes[nd.ch0];
    const int r0 = VSIDE ? c0.v_rank : c0.u_rank;
    const DNode cc = nodes[child];
    const int rc = VSIDE ? cc.v_rank : cc.u_rank;
    const int off = (child == nd.ch0) ? 0 : r0;
    const int n = VSIDE ? nd.v_rows : nd.u_rows, r = VSIDE ? nd.v_rank : nd.u_rank;
    const int* P = perms + (VSIDE ? nd.Pv : nd.Pu);
    const double* E = vals + (VSIDE ? nd.Ev : nd.Eu);
    __syncthreads();
    for (int j = tid; j < n; j += 128) {
      const int t = P[j] - off;
      if (t >= 0 && t < rc) pos[t] = j;
    }
    __syncthreads();
    const int k = n - r;
    for (int c = tid; c < r; c += 128) {
      double acc = 0.;
      for (int t = 0; t < rc; t++) {
        const int p = pos[t];
        acc += ua[t] * (p < r ? (p == c ? 1. : 0.) : E[(p - r) + (size_t)c * k]);
      }
      ub[c] = acc;
    }
    __syncthreads();
    for (int c = tid; c < r; c += 128) { ua[c] = ub[c]; out[(size_t)d * maxr + c] = ub[c]; }
  }
}

__global__ void __launch_bounds__(256)
hss_extract_kernel(const DNode* __restrict__ nodes, const double* __restrict__ vals,
                   const int* __restrict__ I, const int* __restrict__ J, int nI, int nJ,
                   const int* __restrict__ pI, const int* __restrict__ pJ, int maxd, int maxr,
                   const double* __restrict__ cu, const double* __restrict__ cv,
                   double* __restrict__ B, int ldb, int add) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= (long long)nI * nJ) return;
  const int a = (int)(e % nI), b = (int)(e / nI);
  const int* pa = pI + (size_t)a * maxd;
  const int* pb = pJ + (size_t)b * maxd;
  int d = 0;
  while (d < maxd && pa[d] >= 0 && pa[d] == pb[d]) d++;
  double v;
  if (d == maxd || (pa[d] < 0 && pb[d] < 0)) {   // same leaf (the last common node)
    const DNode nd = nodes[d == 0 ? 0 : pa[d - 1]];
    v = vals[nd.D + (I[a] - nd.row_off) + (size_t)(J[b] - nd.col_off) * nd.rows];
  } else {
    const DNode par = nodes[d == 0 ? 0 : pa[d - 1]];
    const DNode na = nodes[pa[d]], nb = nodes[pb[d]];
    const bool first = pa[d] == par.ch0;          // row index below child 0: block B01
    const double* Bk = vals + (first ? par.B01 : par.B10);
    const int ru = na.u_rank, rv = nb.v_rank;
    const double* u = cu + ((size_t)a * maxd + d) * maxr;
    const double* w = cv + ((size_t)b * maxd + d) * maxr;
    v = 0.;
    for (int q = 0; q < rv; q++) {
      double t = 0.;
      for (int p = 0; p < ru; p++) t += u[p] * Bk[p + (size_t)q * ru];
      v += t * w[q];
    }
  }
  double* o = B + a + (size_t)b * ldb;
  *o = add ? *o + v : v;
}

// ===========================================================================
//                               ULV FACTOR
// ===========================================================================
// Per non-root node the factor block F is m x naug (ld = m), columns
//   [0,k)            A = W0^T          -> R (upper) + Householder vectors V
//   [k,k+rv)         Vh                -> Q^T Vh  = [Vt0 ; Vt1]
//   [k+rv,k+rv+r)    W1^T              -> Q^T W1^T = [(W1 Q0^H)^T ; Dt^T]
// W0 = (P_u^T D)_bot - E_u (P_u^T D)_top, W1 = (P_u^T D)_top
// (reference factor.hpp:109-121).  The reference forms Q explicitly
// (DenseMatrix::LQ, DenseMatrix.cpp:693-719); here Q stays in compact-WY form.

// Vh = dense(V) = P_v [I; E_v] for leaves        (factor.hpp:100-103)
__global__ void __launch_bounds__(kThreads)
ulv_vh_leaf_kernel(const DNode* __restrict__ nodes, const int* __restrict__ list,
                   const double* __restrict__ vals, const int* __restrict__ perms,
                   double* __restrict__ fact) {
  const DNode nd = nodes[list[blockIdx.x]];
  if (!nd.leaf || nd.parent < 0) return;
  const int m = nd.m, rv = nd.v_rank, kv = m - rv;
  const int* P = perms + nd.Pv;
  const double* E = vals + nd.Ev;
  double* Vh = fact + nd.F + (size_t)nd.k * m;
  for (int idx = threadIdx.x; idx < m * rv; idx += kThreads) {
    int i = idx % m, c = idx / m;
    double v = i < rv ? (i == c ? 1. : 0.) : E[(i - rv) + (size_t)c * kv];
    Vh[P[i] + (size_t)c * m] = v;
  }
}

__device__ __forceinline__ double child_Dt(const double* f, const DNode& c, int a, int b) {
  // Dt[a,b] = F[(k+b) + (k+rv+a) m]
  return f[c.F + (c.k + b) + (size_t)(c.k + c.v_rank + a) * c.m];
}
// Inner nodes: Dfull = [Dt0, B01 Vt1(c1)^H; B10 Vt1(c0)^H, Dt1] into `dst`
// (ld = m) and, for non-root nodes, Vh = [Vt1(c0) Vd_top; Vt1(c1) Vd_bot]
// into the factor block                         (factor.hpp:68-98)
__global__ void __launch_bounds__(kThreads)
ulv_build_inner_kernel(const DNode* __restrict__ nodes, const int* __restrict__ list,
                       const double* __restrict__ vals, const int* __restrict__ perms,
                       double* __restrict__ fact, double* __restrict__ scratch,
                       const long long* __restrict__ scratch_off, int stage_vt1) {
  extern __shared__ __align__(16) double sm[];
  const int id = list[blockIdx.x];
  const DNode nd = nodes[id];
  if (nd.leaf) return;
  const DNode c0 = nodes[nd.ch0], c1 = nodes[nd.ch1];
  const int ru0 = c0.u_rank, ru1 = c1.u_rank, rv0 = c0.v_rank, rv1 = c1.v_rank;
  const int m = ru0 + ru1, tid = threadIdx.x;
  const int rv = nd.v_rank, nv = rv0 + rv1, kv = nv - rv;
  // the children's Vt1 blocks (ru x rv, strided inside their factor blocks) are
  // staged in shared memory once: every entry is used O(m) times below
  int* pinv = reinterpret_cast<int*>(sm);
  const double* v0 = fact + c0.F + c0.k + (size_t)c0.k * c0.m;   // Vt1(c0)[a,j] = v0[a + j*ld0]
  const double* v1 = fact + c1.F + c1.k + (size_t)c1.k * c1.m;
  int ld0 = c0.m, ld1 = c1.m;
  if (stage_vt1) {
    double* s0 = sm + ((nv + 1) / 2 + 1);
    double* s1 = s0 + ru0 * rv0;
    for (int idx = tid; idx < ru0 * rv0; idx += kThreads) s0[idx] = v0[(idx % ru0) + (size_t)(idx / ru0) * ld0];
    for (int idx = tid; idx < ru1 * rv1; idx += kThreads) s1[idx] = v1[(idx % ru1) + (size_t)(idx / ru1) * ld1];
    v0 = s0; v1 = s1; ld0 = ru0; ld1 = ru1;
  }
  if (nd.parent >= 0) {
    const int* P = perms + nd.Pv;
    for (int i = tid; i < nv; i += kThreads) pinv[P[i]] = i;
  }
  __syncthreads();
  double* Df = scratch + scratch_off[blockIdx.x];
  const double* B01 = vals + nd.B01;
  const double* B10 = vals + nd.B10;
  // gridDim.y CTAs share one node (the upper classes have few nodes and are
  // latency bound): each takes a strided slice of the entries
  const int step = kThreads * gridDim.y, first = tid + blockIdx.y * kThreads;
  for (int idx = first; idx < m * m; idx += step) {
    int i = idx % m, j = idx / m;
    double v;
    if (i < ru0 && j < ru0) v = child_Dt(fact, c0, i, j);
    else if (i >= ru0 && j >= ru0) v = child_Dt(fact, c1, i - ru0, j - ru0);
    else if (i < ru0) {  // B01 * Vt1(c1)^H
      int b = j - ru0;
      v = 0.;
      for (int q = 0; q < rv1; q++) v += B01[i + (size_t)q * ru0] * v1[b + q * ld1];
    } else {             // B10 * Vt1(c0)^H
      int a = i - ru0;
      v = 0.;
      for (int q = 0; q < rv0; q++) v += B10[a + (size_t)q * ru1] * v0[j + q * ld0];
    }
    Df[i + (size_t)j * m] = v;
  }
  if (nd.parent < 0) return;
  // Vd = dense(V) = P_v [I; E_v] is never formed: row j of Vd is row
  // pinv[j] of [I; E_v]; only the inverse permutation sits in smem.
  const double* E = vals + nd.Ev;
  double* Vh = fact + nd.F + (size_t)nd.k * m;
  for (int idx = first; idx < m * rv; idx += step) {
    int i = idx % m, c = idx / m;
    double v = 0.;
    const bool top = i < ru0;
    const double* vv = top ? v0 : v1;
    const int ldv_ = top ? ld0 : ld1;
    const int a = top ? i : i - ru0, qoff = top ? 0 : rv0, nq = top ? rv0 : rv1;
    for (int q = 0; q < nq; q++) {
      const int pi = pinv[qoff + q];
      const double vd = pi < rv ? (pi == c ? 1. : 0.) : E[(pi - rv) + (size_t)c * kv];
      v += vv[a + q * ldv_] * vd;
    }
    Vh[i + (size_t)c * m] = v;
  }
}

// Elimination step: writes A = W0^T and W1^T into the factor block.
//   W0^T[j, i] = Dp[r+i, j] - sum_l E[i, l] Dp[l, j],   W1^T[j, l] = Dp[l, j],
//   Dp = P_u^T D                                     (factor.hpp:109-121)
// Dsrc = leaf ? D generator : Dfull scratch.  Tiles of TJ columns of Dsrc are
// read coalesced into shared memory, so the row gather by P_u is a shared-
// memory gather; the rank-r correction runs on the fp64 tensor pipe: each
// warp owns 8-column slabs of the output, A fragments are gathered rows of the
// tile, B fragments (E^T) stay in registers across the tile's row blocks.
// HBM-bound: reads D once, writes the m x (k + r) block once.
// PIPE: the column tiles are prefetched one ahead with cp.async into a second
// buffer (the loads of tile t+1 overlap the products and stores of tile t).
template <int TJ, int MINB = 3, bool PIPE = false>
__global__ void __launch_bounds__(kThreads, MINB)
ulv_eliminate_kernel(const DNode* __restrict__ nodes, const int* __restrict__ list,
                     const double* __restrict__ vals, const int* __restrict__ perms,
                     double* __restrict__ fact, const double* __restrict__ scratch,
                     const long long* __restrict__ scratch_off) {
  extern __shared__ __align__(16) double sm[];
  const DNode nd = nodes[list[blockIdx.x]];
  if (nd.parent < 0) return;
  const int m = nd.m, r = nd.u_rank, k = nd.k, rv = nd.v_rank, tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const double* Dsrc = nd.leaf ? vals + nd.D : scratch + scratch_off[blockIdx.x];
  const int* P = perms + nd.Pu;
  const double* E = vals + nd.Eu;  // k x r
  double* A = fact + nd.F;
  double* W1t = A + (size_t)(k + rv) * m;
  constexpr int LD = TJ + 1;
  constexpr int RT = TJ / 8;       // row tiles of the output per column tile
  constexpr int KS = 8;            // k-steps held in registers (r <= 32 per pass)
  double* Dn0 = sm;                // m x LD : Dn[i*LD + jj] = Dsrc[i, j0+jj]  (x 2 buffers if PIPE)
  int* Ps = reinterpret_cast<int*>(Dn0 + (size_t)m * LD * (PIPE ? 2 : 1));   // m
  for (int i = tid; i < m; i += kThreads) Ps[i] = P[i];
  auto issue = [&](int j0, int buf) {
    const int tj = min(TJ, m - j0);
    double* D = Dn0 + (size_t)buf * m * LD;
    for (int jj = warp; jj < TJ; jj += kWarps) {
      const bool jin = jj < tj;
      const double* src = Dsrc + (size_t)(jin ? j0 + jj : 0) * m;
      for (int i = lane; i < m; i += 32) cp_async8(D + i * LD + jj, src + i, jin);
    }
    cp_async_commit();
  };
  auto load_e = [&](double (&bn)[KS], int i0, int l0) {
#pragma unroll
    for (int s = 0; s < KS; s++) {
      const int l = l0 + 4 * s + t;
      // B[k=l][n=i] = -E[i, l]
      bn[s] = (l < r && i0 + g < k) ? E[(i0 + g) + (size_t)l * k] : 0.;
    }
  };
  const bool onepass = r <= 4 * KS;
  double bnext[KS];
  if (PIPE) issue(0, 0);
  int buf = 0;
  for (int j0 = 0; j0 < m; j0 += TJ, buf ^= (PIPE ? 1 : 0)) {
    const int tj = min(TJ, m - j0);
    __syncthreads();
    if (PIPE) {
      if (onepass && warp * 8 < k) load_e(bnext, warp * 8, 0);
      if (j0 + TJ < m) { issue(j0 + TJ, buf ^ 1); cp_async_wait<1>(); }
      else cp_async_wait<0>();
    } else {
      // one buffer, but the whole tile in flight at once (LDGSTS): loads through registers kept 4 per thread
      // in flight and exposed 8 HBM round trips per tile (ncu source page: 39 % of the samples on their STS)
      issue(j0, 0);
      if (onepass && warp * 8 < k) load_e(bnext, warp * 8, 0);   // under the tile's HBM round trip
      cp_async_wait<0>();
    }
    __syncthreads();
    const double* Dn = Dn0 + (size_t)buf * m * LD;
    // W1^T[j, l] = Dp[l, j]
    for (int l = warp; l < r; l += kWarps) {
      const double* src = Dn + Ps[l] * LD;
      for (int jj = lane; jj < tj; jj += 32) W1t[(j0 + jj) + (size_t)l * m] = src[jj];
    }
    // W0^T tile (tj x k), 8-column slabs over the warps.  The E fragment of a slab (B operand, -E^T) does not
    // depend on the column tile but cannot stay in registers for all slabs of a warp: it comes from L2 once per
    // tile, and the fragment of the NEXT slab is requested before the products of the current one (ncu source
    // page: 17 % of the samples waited on these loads when they were issued right before their first use).
    for (int i0 = warp * 8; i0 < k; i0 += kWarps * 8) {
      double acc[RT][2];
      const int ia = i0 + 2 * t, ib = ia + 1;       // output columns of this lane
      const double* ga = Dn + (ia < k ? Ps[r + ia] : 0) * LD + g;
      const double* gb = Dn + (ib < k ? Ps[r + ib] : 0) * LD + g;
#pragma unroll
      for (int rt = 0; rt < RT; rt++) { acc[rt][0] = ga[rt * 8]; acc[rt][1] = gb[rt * 8]; }
      for (int l0 = 0; l0 < r; l0 += 4 * KS) {
        double bneg[KS];
        int prow[KS];
        if (onepass) {
#pragma unroll
          for (int s = 0; s < KS; s++) bneg[s] = -bnext[s];
          if (i0 + kWarps * 8 < k) load_e(bnext, i0 + kWarps * 8, 0);
        } else {
          load_e(bneg, i0, l0);
#pragma unroll
          for (int s = 0; s < KS; s++) bneg[s] = -bneg[s];
        }
#pragma unroll
        for (int s = 0; s < KS; s++) {
          const int l = l0 + 4 * s + t;
          prow[s] = (l < r ? Ps[l] : 0) * LD + g;
        }
#pragma unroll
        for (int s = 0; s < KS; s++) {
          if (l0 + 4 * s < r) {
            const bool lin = l0 + 4 * s + t < r;
#pragma unroll
            for (int rt = 0; rt < RT; rt++) {
              // A[row=jj][k=l] = Dp[l, j] = Dn[P[l]][jj]
              const double a = lin ? Dn[prow[s] + rt * 8] : 0.;
              dmma(acc[rt][0], acc[rt][1], a, bneg[s]);
            }
          }
        }
      }
#pragma unroll
      for (int rt = 0; rt < RT; rt++) {
        const int jj = rt * 8 + g;
        if (jj < tj) {
          if (ia < k) A[(j0 + jj) + (size_t)ia * m] = acc[rt][0];
          if (ib < k) A[(j0 + jj) + (size_t)ib * m] = acc[rt][1];
        }
      }
    }
  }
}

// Trailing update of one 8-column slab C = A[j0:m, c0:c0+cw) with the block
// reflector of the current panel:  C <- C - V (T^T (V^T C)).
// V (mp x NB, explicit unit-lower trapezoid, zero padded to a multiple of 8
// rows) and T (NB x NB upper) live in shared memory; C streams straight from
// global/L2 into fp64 tensor-core fragments (no shared-memory staging) and is
// read twice (second read hits L1), written once.  One warp per slab, no CTA
// barrier: the 8 warps of a CTA work on different slabs asynchronously.
//   phase 1  W^T[n][a]  = sum_i C[i][n] V[i][a]        (A = C^T, B = V)
//   phase T  W2^T       = W^T T                         (A from accumulators
//            via the K-slot permutation a = 8*at + 2t + e, B = T)
//   phase 2  C[i][n]   -= sum_a V[i][a] W2[a][n]       (A = V, B = -W2)
// U = row tiles per prefetch group (2 U loads in flight per lane): the slab
// comes from L2 at ~1-2 k clk under load, memory-level parallelism per warp is
// what bounds the update when few warps are in this phase.
template <int NB, int U = 4>
__device__ __forceinline__ void slab_update(double* __restrict__ Cg, int ldc, int mp,
                                            int cw, const double* __restrict__ Vs,
                                            int ldv, const double* __restrict__ Ts,
                                            int ldt, int lane) {
  constexpr int NT = NB / 8;
  const int g = lane >> 2, t = lane & 3;
  const int nit = (mp + 7) >> 3;
  double wt[NT][2], wu[NT][2];   // two accumulator sets: shorter DMMA chains
#pragma unroll
  for (int q = 0; q < NT; q++) wt[q][0] = wt[q][1] = wu[q][0] = wu[q][1] = 0.;
  const bool gin = g < cw;
  {
    const double* Cn = Cg + (size_t)g * ldc + t;
    const double* Vb = Vs + t + g * ldv;
    for (int it0 = 0; it0 < nit; it0 += U) {
      double a[U][2];
#pragma unroll
      for (int u = 0; u < U; u++)
#pragma unroll
        for (int ks = 0; ks < 2; ks++) {
          const int i = (it0 + u) * 8 + ks * 4;
          a[u][ks] = (gin && i + t < mp) ? Cn[i] : 0.;
        }
#pragma unroll
      for (int u = 0; u < U; u++) {
        const int it = it0 + u;
        if (it < nit) {
          const double* vb = Vb + it * 8;
#pragma unroll
          for (int at = 0; at < NT; at++)
            if (at <= it) {  // V[i][a] = 0 for a > i
              dmma(wt[at][0], wt[at][1], a[u][0], vb[at * 8 * ldv]);
              dmma(wu[at][0], wu[at][1], a[u][1], vb[4 + at * 8 * ldv]);
            }
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < NT; q++) { wt[q][0] += wu[q][0]; wt[q][1] += wu[q][1]; }
  double w2[NT][2];
#pragma unroll
  for (int q = 0; q < NT; q++) w2[q][0] = w2[q][1] = 0.;
#pragma unroll
  for (int atp = 0; atp < NT; atp++)
#pragma unroll
    for (int at = 0; at < NT; at++)
      if (at <= atp) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const double bb = Ts[(at * 8 + 2 * t + e) + (atp * 8 + g) * ldt];
          dmma(w2[atp][0], w2[atp][1], wt[at][e], bb);
        }
      }
  // accumulator layout -> B-fragment layout (natural K order), negated
  double bneg[NB / 4];
#pragma unroll
  for (int s4 = 0; s4 < NB / 4; s4++) {
    const int src = (g << 2) + ((s4 & 1) << 1) + (t >> 1);
    const double v0 = __shfl_sync(0xffffffffu, w2[s4 >> 1][0], src);
    const double v1 = __shfl_sync(0xffffffffu, w2[s4 >> 1][1], src);
    bneg[s4] = -((t & 1) ? v1 : v0);
  }
  const bool c0in = 2 * t < cw, c1in = 2 * t + 1 < cw;
  double* C0 = Cg + (size_t)(2 * t) * ldc + g;
  double* C1 = C0 + ldc;
  const double* Va = Vs + g + t * ldv;
  for (int it0 = 0; it0 < nit; it0 += U) {
    double acc[U][2];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int i = (it0 + u) * 8;
      const bool iin = i + g < mp;
      acc[u][0] = (iin && c0in) ? C0[i] : 0.;
      acc[u][1] = (iin && c1in) ? C1[i] : 0.;
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int it = it0 + u;
      if (it < nit) {
        const double* va = Va + it * 8;
#pragma unroll
        for (int s4 = 0; s4 < NB / 4; s4++)
          if (s4 <= 2 * it + 1)   // V[i][a] = 0 for a > i
            dmma(acc[u][0], acc[u][1], va[s4 * 4 * ldv], bneg[s4]);
        const bool iin = it * 8 + g < mp;
        if (iin && c0in) C0[it * 8] = acc[u][0];
        if (iin && c1in) C1[it * 8] = acc[u][1];
      }
    }
  }
}

// K-slot / column permutation used by slab2_update inside every 8-wide tile:
// sigma = (0,2,1,3,6,4,7,5).  sigma(2q+1) - sigma(2q) = 2 (mod 4) and
// {sigma(2t+e) mod 4 : t} = {0,1,2,3} make both 128-bit fragment loads of V
// from shared memory (ld = 4 mod 16 doubles) bank-conflict free.
__device__ __forceinline__ int sig8(int j) { return (0x57463120u >> (4 * j)) & 7; }

__device__ __forceinline__ double2 ld2(const double* p, bool pred) {
  return pred ? *reinterpret_cast<const double2*>(p) : make_double2(0., 0.);
}

// Trailing update of TWO adjacent 8-column slabs by one warp (same math as
// slab_update).  The LSU wavefront count, not the tensor pipe, bounds
// slab_update (ncu: l1tex data pipe 54-63 % busy over the whole kernel, DMMA
// 31 %): every DMMA takes a 64-bit V fragment from shared memory (2 wavefronts)
// and the 8-byte C accesses touch 4-8 lines per instruction.  Here
//  * every V fragment load is 128 bits wide and feeds 4 DMMAs (two k-steps or
//    two row tiles, times two slabs);
//  * C moves in 16-byte pieces: phase 1 takes rows (2t, 2t+1) of column g as
//    the two k-steps of a tile (K-slot permutation), phase 2 works on 16-row
//    blocks whose two DMMA row tiles are the even / odd rows, so a lane owns
//    rows (2g, 2g+1) of two columns and an instruction covers whole 128-byte
//    lines;
//  * W^T T is chained through the accumulators with the column permutation
//    sigma, W2 is already in B-fragment layout for phase 2 (no shuffles).
// Needs 16-byte aligned columns: ldc even, Cg 16-byte aligned, mp even.
template <int NB, bool TWO = true>
__device__ __forceinline__ void slab2_update(double* __restrict__ Cg, int ldc, int mp,
                                             int cw0, int cw1,
                                             const double* __restrict__ Vs, int ldv,
                                             const double* __restrict__ Ts, int ldt,
                                             int lane) {
  constexpr int NT = NB / 8;
  // row blocks per load group: with 16-column panels there are 4 CTAs of 4
  // warps per SM and on average ~2 warps per scheduler in this phase, so each
  // warp must keep more of the slab in flight to cover the L2 round trip
  constexpr int U1 = NB <= 16 ? 8 : 4, U2 = NB <= 16 ? 4 : 2;
  const int g = lane >> 2, t = lane & 3;
  const int sg = sig8(g), s0 = sig8(2 * t), s1 = sig8(2 * t + 1);
  double wt[2][NT][2];
#pragma unroll
  for (int s = 0; s < 2; s++)
#pragma unroll
    for (int q = 0; q < NT; q++) wt[s][q][0] = wt[s][q][1] = 0.;
  // ---- phase 1: W^T[n][a] = sum_i C[i][n] V[i][a]; k-steps (e) of row block
  // i0 are rows i0 + 2t + e; B column g is reflector 8 at + sigma(g)
  {
    const int nb8 = (mp + 7) >> 3;
    const bool in0 = g < cw0, in1 = TWO && g < cw1;
    const double* C0 = Cg + (size_t)g * ldc + 2 * t;
    const double* C1 = Cg + (size_t)(8 + g) * ldc + 2 * t;
    const double* Vb = Vs + 2 * t + sg * ldv;
    for (int ib0 = 0; ib0 < nb8; ib0 += U1) {
      double2 c[U1][2];
#pragma unroll
      for (int u = 0; u < U1; u++) {
        const int i0 = (ib0 + u) * 8;
        const bool rin = i0 + 2 * t < mp;
        c[u][0] = ld2(C0 + i0, rin && in0);
        if (TWO) c[u][1] = ld2(C1 + i0, rin && in1);
      }
#pragma unroll
      for (int u = 0; u < U1; u++) {
        const int ib = ib0 + u;
        if (ib < nb8) {
#pragma unroll
          for (int at = 0; at < NT; at++)
            if (at <= ib) {   // V[i][a] = 0 for a > i
              const double2 v = *reinterpret_cast<const double2*>(Vb + ib * 8 + at * 8 * ldv);
              dmma(wt[0][at][0], wt[0][at][1], c[u][0].x, v.x);
              if (TWO) dmma(wt[1][at][0], wt[1][at][1], c[u][1].x, v.x);
              dmma(wt[0][at][0], wt[0][at][1], c[u][0].y, v.y);
              if (TWO) dmma(wt[1][at][0], wt[1][at][1], c[u][1].y, v.y);
            }
        }
      }
    }
  }
  // ---- phase T: W2^T = W^T T.  wt[s][at][e] is the A fragment of k-step
  // (at, e) with K slot t <-> reflector 8 at + sigma(2t+e); output column g <->
  // reflector 8 atp + sigma(g), so w2[s][atp][e] = W2[8 atp + sigma(2t+e)][n = g]
  double w2[2][NT][2];
#pragma unroll
  for (int s = 0; s < 2; s++)
#pragma unroll
    for (int q = 0; q < NT; q++) w2[s][q][0] = w2[s][q][1] = 0.;
#pragma unroll
  for (int atp = 0; atp < NT; atp++)
#pragma unroll
    for (int at = 0; at < NT; at++)
      if (at <= atp) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const double bb = Ts[(at * 8 + (e ? s1 : s0)) + (atp * 8 + sg) * ldt];
          dmma(w2[0][atp][0], w2[0][atp][1], wt[0][at][e], bb);
          if (TWO) dmma(w2[1][atp][0], w2[1][atp][1], wt[1][at][e], bb);
        }
      }
#pragma unroll
  for (int s = 0; s < 2; s++)
#pragma unroll
    for (int q = 0; q < NT; q++) { w2[s][q][0] = -w2[s][q][0]; w2[s][q][1] = -w2[s][q][1]; }
  // ---- phase 2: C -= V W2 on 16-row blocks; DMMA row tile u holds rows
  // r0 + 2g + u, so a lane owns rows (2g, 2g+1) of columns 2t and 2t+1
  {
    const int nb16 = (mp + 15) >> 4;
    const bool k00 = 2 * t < cw0, k01 = 2 * t + 1 < cw0, k10 = TWO && 2 * t < cw1, k11 = TWO && 2 * t + 1 < cw1;
    double* P00 = Cg + (size_t)(2 * t) * ldc + 2 * g;
    double* P01 = P00 + ldc;
    double* P10 = Cg + (size_t)(8 + 2 * t) * ldc + 2 * g;
    double* P11 = P10 + ldc;
    const double* Va0 = Vs + 2 * g + s0 * ldv;
    const double* Va1 = Vs + 2 * g + s1 * ldv;
    for (int ib0 = 0; ib0 < nb16; ib0 += U2) {
      double2 a[U2][4];
#pragma unroll
      for (int u = 0; u < U2; u++) {
        const int r0 = (ib0 + u) * 16;
        const bool rin = r0 + 2 * g < mp;
        a[u][0] = ld2(P00 + r0, rin && k00);
        a[u][1] = ld2(P01 + r0, rin && k01);
        if (TWO) {
          a[u][2] = ld2(P10 + r0, rin && k10);
          a[u][3] = ld2(P11 + r0, rin && k11);
        }
      }
#pragma unroll
      for (int u = 0; u < U2; u++) {
        const int ib = ib0 + u;
        if (ib < nb16) {
          const int r0 = ib * 16;
#pragma unroll
          for (int atp = 0; atp < NT; atp++)
            if (atp <= 2 * ib + 1) {   // V[i][a] = 0 for a > i
#pragma unroll
              for (int e = 0; e < 2; e++) {
                const double2 v = *reinterpret_cast<const double2*>((e ? Va1 : Va0) + r0 + atp * 8 * ldv);
                dmma(a[u][0].x, a[u][1].x, v.x, w2[0][atp][e]);
                if (TWO) dmma(a[u][2].x, a[u][3].x, v.x, w2[1][atp][e]);
                dmma(a[u][0].y, a[u][1].y, v.y, w2[0][atp][e]);
                if (TWO) dmma(a[u][2].y, a[u][3].y, v.y, w2[1][atp][e]);
              }
            }
          const bool rin = r0 + 2 * g < mp;
          if (rin && k00) *reinterpret_cast<double2*>(P00 + r0) = a[u][0];
          if (rin && k01) *reinterpret_cast<double2*>(P01 + r0) = a[u][1];
          if (TWO && rin && k10) *reinterpret_cast<double2*>(P10 + r0) = a[u][2];
          if (TWO && rin && k11) *reinterpret_cast<double2*>(P11 + r0) = a[u][3];
        }
      }
    }
  }
}

// Blocked Householder QR of the first k columns of the m x naug factor block,
// reflectors applied to all naug columns (right-looking, panel width NB).
// One CTA per node, 2 CTAs per SM.
//  * the panel (mp x NB) lives in shared memory and is factored in sub-panels
//    of 8 columns: unblocked Householder inside a sub-panel with ONE barrier
//    per column (norms are fused into the previous update, the scaling of a
//    reflector is deferred, the warps that have no column to update compute
//    the V^T v dot products that give the sub-panel's T factor for free);
//    the rest of the panel is then updated on the fp64 tensor pipe
//    (slab_update<8> on shared memory);
//  * the NB x NB T factor is merged from the 8x8 ones (block dlarft);
//  * the trailing matrix is updated slab by slab on the tensor pipe straight
//    from L2 (slab_update<NB>), one warp per slab, no barrier.
//  NTH = 256: 2 CTAs per SM; NTH = 128 (register sub-panel only, NB = 16): 4
//  CTAs per SM, i.e. four Householder column chains in flight per SM, so that
//  the fp64 tensor pipe always finds a CTA in its trailing update.
template <int NB, bool SMALL, int NTH, int MINB = (NTH == 128 ? 4 : 2)>
__global__ void __launch_bounds__(NTH, MINB)
ulv_qr_kernel(const DNode* __restrict__ nodes, const int* __restrict__ list,
              double* __restrict__ fact, double* __restrict__ tfac, int ldv,
              int nowide) {
  static_assert(NTH == 256 || SMALL, "the shared-memory sub-panel variant needs 8 warps");
  constexpr int NWP = NTH / 32;
  extern __shared__ __align__(16) double sm[];
  const DNode nd = nodes[list[blockIdx.x]];
  if (nd.parent < 0 || nd.k == 0) return;
  const int m = nd.m, k = nd.k, naug = nd.naug, tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  constexpr int LDW = ((NB + 15) / 16) * 16 + 4;
  double* Vs = sm;                       // ldv x NB
  double* Ws = Vs + (size_t)ldv * NB;    // LDW x NB : S = V^T V blocks
  double* Ts = Ws + LDW * NB;            // LDW x NB : T
  double* Ys = Ts + LDW * NB;            // LDW x 8
  double* tau = Ys + LDW * 8;            // NB
  double* betas = tau + NB;              // NB
  double* nrm2s = betas + NB;            // NB
  double* scals = nrm2s;                 // NB (1/(alpha-beta) per column)
  double* zs = nrm2s + NB;               // 2 x 16
  double* A = fact + nd.F;
  double* Tg = tfac + nd.T;
  // every column of the factor block starts on a 16-byte boundary (even F, even m)
  const bool wide = !nowide && ((m & 1) == 0) && ((nd.F & 1) == 0);
#ifdef SB200_QR_TIMING
  long long tph[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  long long tlast = clock64();
#define QR_TICK(p) { long long now_ = clock64(); tph[p] += now_ - tlast; tlast = now_; }
#else
#define QR_TICK(p)
#endif
  for (int j0 = 0; j0 < k; j0 += NB) {
    const int jb = min(NB, k - j0), mp = m - j0;
    const int mp8 = (mp + 7) & ~7;
    // ---- load panel (zero padded to mp8 rows / NB columns), clear T.  The
    // panel comes from L2 (it was just written by the trailing update): LDGSTS
    // keeps every element of the panel in flight at once instead of one L2
    // round trip per load -> store pair.
    for (int c = warp; c < NB; c += NWP) {
      const double* src = A + j0 + (size_t)(j0 + c) * m;
      double* dst = Vs + c * ldv;
      const bool cin = c < jb;
      for (int i = lane; i < mp8; i += 32) cp_async8(dst + i, src + i, cin && i < mp);
    }
    cp_async_wait_all();
    for (int idx = tid; idx < NB * NB; idx += NTH) Ts[(idx % NB) + (idx / NB) * LDW] = 0.;
    __syncthreads();
    QR_TICK(0)
    const int nsub = (jb + 7) >> 3;
    for (int sp = 0; sp < nsub; sp++) {
      const int cs = sp * 8, sbw = min(8, jb - cs);
      if (SMALL) {
        // ---- register-resident sub-panel (mp <= 256): warps 0..3 hold the 8
        // columns (2 x 32 rows per lane and warp) in registers, ONE fused
        // 8-value shuffle reduction per column gives the pivot norm, the dot
        // products with the columns still to update AND the v_b^T v_c products
        // of the T factor; the 4 partial results are combined through shared
        // memory behind a 128-thread named barrier.  Warps 4..7 wait at the
        // CTA barrier below.  No per-column CTA barrier, no smem traffic for
        // the column data.
        if (warp < 4) {
          double a[8][2];
          const int r0 = warp * 64 + lane;
#pragma unroll
          for (int q = 0; q < 8; q++)
#pragma unroll
            for (int rr = 0; rr < 2; rr++) {
              const int i = r0 + 32 * rr;
              a[q][rr] = i < mp ? Vs[i + (cs + q) * ldv] : 0.;
            }
          double tr[8];
#pragma unroll
          for (int q = 0; q < 8; q++) tr[q] = 0.;
          double* pair = zs;            // [2 parities][4 warps][8]
          double* diag = Ws;            // [2 parities][8]   (Ws is free during the steps)
#pragma unroll
          for (int cq = 0; cq < 8; cq++) {
            if (cq < sbw) {
              const int c = cs + cq;   // diagonal row: c < 32, lives in warp 0, rr = 0, lane c
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
              // transpose-reduce: 9 shuffles instead of 40; lane L ends up with
              // the warp total of value (L >> 2) & 7
              double rsum;
              {
                const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
                double r1[4], r2[2];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                  const double send = b4 ? p[q] : p[q + 4];
                  const double keep = b4 ? p[q + 4] : p[q];
                  r1[q] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                }
#pragma unroll
                for (int q = 0; q < 2; q++) {
                  const double send = b3 ? r1[q] : r1[q + 2];
                  const double keep = b3 ? r1[q + 2] : r1[q];
                  r2[q] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                }
                {
                  const double send = b2 ? r2[0] : r2[1];
                  const double keep = b2 ? r2[1] : r2[0];
                  rsum = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                }
                rsum += __shfl_xor_sync(0xffffffffu, rsum, 2);
                rsum += __shfl_xor_sync(0xffffffffu, rsum, 1);
              }
              double* pw = pair + (cq & 1) * 32 + warp * 8;
              if ((lane & 3) == 0) pw[lane >> 2] = rsum;
              if (warp == 0 && lane == c) {
#pragma unroll
                for (int q = 0; q < 8; q++) diag[(cq & 1) * 8 + q] = a[q][0];
              }
              asm volatile("bar.sync 1, 128;" ::: "memory");
              const double* pp = pair + (cq & 1) * 32;
              double sm_[8], dg[8];
              {
                const double2* p2 = reinterpret_cast<const double2*>(pp);
                const double2* d2 = reinterpret_cast<const double2*>(diag + (cq & 1) * 8);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                  const double2 a0 = p2[q], a1 = p2[4 + q], a2 = p2[8 + q], a3 = p2[12 + q], dd = d2[q];
                  sm_[2 * q] = (a0.x + a1.x) + (a2.x + a3.x);
                  sm_[2 * q + 1] = (a0.y + a1.y) + (a2.y + a3.y);
                  dg[2 * q] = dd.x; dg[2 * q + 1] = dd.y;
                }
              }
              const double pn = sm_[cq], alpha = dg[cq];
              double tc = 0., scal = 0., beta = alpha;
              if (pn > 0.) {   // dlarfg
                beta = -copysign(sqrt(alpha * alpha + pn), alpha);
                const double d = alpha - beta;
                scal = 1. / d;
                tc = -d / beta;
              }
              const bool isdiag = (warp == 0 && lane == c);
#pragma unroll
              for (int q = 0; q < 8; q++) {
                if (q > cq) {   // apply H_c to the columns to the right
                  const double w = tc * (dg[q] + scal * sm_[q]);
                  const double ws = w * scal;
#pragma unroll
                  for (int rr = 0; rr < 2; rr++)
                    if (r0 + 32 * rr > c) a[q][rr] -= ws * a[cq][rr];
                  if (isdiag) a[q][0] -= w;
                }
              }
#pragma unroll
              for (int rr = 0; rr < 2; rr++)
                if (r0 + 32 * rr > c) a[cq][rr] *= scal;     // v_c
              if (isdiag) a[cq][0] = beta;                    // R(c,c)
              // column cq of the 8x8 T (dlarft): lane a (< 8, warp 0) keeps row a
              {
                double val = (lane == cq) ? tc : 0.;
                double acc = 0.;
#pragma unroll
                for (int b = 0; b < 8; b++)
                  if (b < cq) acc += (b >= lane ? tr[b] : 0.) * (dg[b] + scal * sm_[b]);
                if (lane < cq) val = -tc * acc;
                tr[cq] = val;
              }
            }
          }
#pragma unroll
          for (int q = 0; q < 8; q++)
#pragma unroll
            for (int rr = 0; rr < 2; rr++) {
              const int i = r0 + 32 * rr;
              if (i < mp) Vs[i + (cs + q) * ldv] = a[q][rr];
            }
          if (warp == 0 && lane < 8) {
#pragma unroll
            for (int q = 0; q < 8; q++) Ts[(cs + lane) + (cs + q) * LDW] = (q >= lane) ? tr[q] : 0.;
          }
        }
        QR_TICK(1)
      } else {
      for (int cq = 0; cq < sbw; cq++) {
          const int c = cs + cq;
          const double* x = Vs + c * ldv;
          const int nupd = sbw - 1 - cq;
          // role of this warp in step c:
          //   lead (warp < nupd, or warp 7 when nothing is left to update):
          //        norm of the pivot column + scalars (dlarfg) and, if warp <
          //        nupd, apply H_c to column cc = c+1+warp
          //   dot  (nupd <= warp, ww < cq): raw dot product v_b . x_c for the
          //        earlier reflector b = cs+ww (gives the 8x8 T factor for free)
          //   none: straight to the barrier
          const int ww = warp - nupd;
          const bool upd = warp < nupd, lead = upd || (nupd == 0 && warp == 7);
          const bool dot = !lead && ww < cq;
          const bool writer = nupd > 0 ? warp == 0 : warp == 7;
          if (lead || dot) {
            double* oth = upd ? Vs + (c + 1 + warp) * ldv : Vs + (cs + (dot ? ww : 0)) * ldv;
            const bool use_oth = upd || dot;
            double pn = 0., wr = 0.;
            double xv[SMALL ? 8 : 1], ov[SMALL ? 8 : 1];
            if (SMALL) {
#pragma unroll
              for (int q = 0; q < 8; q++) {
                const int i = lane + 32 * q;
                const bool in = i > c && i < mp;
                xv[q] = in ? x[i] : 0.;
                ov[q] = (in && use_oth) ? oth[i] : 0.;
              }
              double p0 = 0., p1 = 0., w0 = 0., w1 = 0.;
#pragma unroll
              for (int q = 0; q < 8; q += 2) {
                p0 += xv[q] * xv[q]; p1 += xv[q + 1] * xv[q + 1];
                w0 += xv[q] * ov[q]; w1 += xv[q + 1] * ov[q + 1];
              }
              pn = p0 + p1; wr = w0 + w1;
            } else {
              double p0 = 0., p1 = 0., w0 = 0., w1 = 0.;
              int i = c + 1 + lane;
              for (; i + 32 < mp; i += 64) {
                const double a0 = x[i], a1 = x[i + 32];
                p0 += a0 * a0; p1 += a1 * a1;
                if (use_oth) { w0 += a0 * oth[i]; w1 += a1 * oth[i + 32]; }
              }
              if (i < mp) { const double a0 = x[i]; p0 += a0 * a0; if (use_oth) w0 += a0 * oth[i]; }
              pn = p0 + p1; wr = w0 + w1;
            }
            QR_TICK(8)
            if (lead) {
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) {
                pn += __shfl_xor_sync(0xffffffffu, pn, o);
                wr += __shfl_xor_sync(0xffffffffu, wr, o);
              }
              QR_TICK(9)
              const double alpha = x[c];
              double tc = 0., scal = 0., beta = alpha;
              if (pn > 0.) {   // dlarfg
                beta = -copysign(sqrt(alpha * alpha + pn), alpha);
                const double d = alpha - beta;
                scal = 1. / d;
                tc = -d / beta;
              }
              QR_TICK(10)
              if (upd) {
                const double colc = oth[c];
                const double w = tc * (colc + scal * wr);
                const double ws = w * scal;
                if (SMALL) {
#pragma unroll
                  for (int q = 0; q < 8; q++) {
                    const int i = lane + 32 * q;
                    if (i > c && i < mp) oth[i] = ov[q] - ws * xv[q];
                  }
                } else {
                  for (int i = c + 1 + lane; i < mp; i += 32) oth[i] -= ws * x[i];
                }
                if (lane == 0) oth[c] = colc - w;
              }
              if (writer && lane == 0) { betas[c] = beta; tau[c] = tc; scals[c] = scal; }
              QR_TICK(11)
            } else {
              // raw z_b = x_b[c+1:]^T x_c[c+1:] and x_b[c]; the scalings (the
              // previous reflector is still unscaled) are applied by warp 7
              wr = warp_sum(wr);
              if (lane == 0) { zs[(cq & 1) * 16 + ww] = wr; zs[(cq & 1) * 16 + 8 + ww] = oth[c]; }
              if (ww == cq - 1) {   // deferred scaling of the previous reflector
                const double sp_ = scals[c - 1];
                __syncwarp();
                for (int i = c + lane; i < mp; i += 32) oth[i] *= sp_;
              }
            }
          }
          QR_TICK(12)
          __syncthreads();
          QR_TICK(13)
          if (warp == 7 && lane <= cq) {   // column cq of the 8x8 T (dlarft)
            const int a = lane;
            const double tc = tau[c], sc = scals[c];
            double val = tc;
            if (a < cq) {
              const double sprev = scals[c - 1];
              double acc = 0.;
              for (int b = a; b < cq; b++) {
                const double f = (b == cq - 1) ? sprev : 1.;   // v_b unscaled?
                const double z = f * (zs[(cq & 1) * 16 + 8 + b] + sc * zs[(cq & 1) * 16 + b]);
                acc += Ts[(cs + a) + (cs + b) * LDW] * z;
              }
              val = -tc * acc;
            }
            Ts[(cs + a) + (cs + cq) * LDW] = val;
          }
          QR_TICK(14)
        }
        const double scal_prev = scals[cs + sbw - 1];
        QR_TICK(1)
        {  // scale the last reflector of the sub-panel
          const int c = cs + sbw - 1;
          for (int i = c + 1 + tid; i < mp; i += NTH) Vs[i + c * ldv] *= scal_prev;
        }
      }
      __syncthreads();
      // ---- R entries of these columns to global; V explicit (unit diagonal)
      for (int cw = warp; cw < sbw; cw += NWP) {
        const int c = cs + cw;
        double* dst = A + j0 + (size_t)(j0 + c) * m;
        for (int i = lane; i <= c; i += 32) {
          dst[i] = (i == c && !SMALL) ? betas[c] : Vs[i + c * ldv];
          Vs[i + c * ldv] = (i == c) ? 1. : 0.;
        }
      }
      __syncthreads();
      QR_TICK(2)
      // ---- update the rest of the panel with this sub-panel's block reflector
      const int nrest = jb - cs - 8;
      if (nrest > 0) {
        if (warp * 8 < nrest)
          slab_update<8>(Vs + cs + (cs + 8 + warp * 8) * ldv, ldv, mp - cs,
                         min(8, nrest - warp * 8), Vs + cs + cs * ldv, ldv,
                         Ts + cs + cs * LDW, LDW, lane);
        __syncthreads();
      }
      QR_TICK(3)
    }
    // ---- merge the 8x8 T factors into the NB x NB one (block dlarft)
    if (nsub > 1) {
      {  // S(bi,bj) = V_bi^T V_bj for bi < bj, one 8x8 tile per warp
        int tile = 0;
        for (int bj = 1; bj < nsub; bj++)
          for (int bi = 0; bi < bj; bi++, tile++)
            if ((tile % NWP) == warp) {
              const int g = lane >> 2, t = lane & 3;
              double c0 = 0., c1 = 0.;
              const double* va = Vs + t + (bi * 8 + g) * ldv;
              const double* vb = Vs + t + (bj * 8 + g) * ldv;
              for (int i = bj * 8; i < mp8; i += 4) dmma(c0, c1, va[i], vb[i]);
              Ws[(bi * 8 + g) + (bj * 8 + 2 * t) * LDW] = c0;
              Ws[(bi * 8 + g) + (bj * 8 + 2 * t + 1) * LDW] = c1;
            }
      }
      __syncthreads();
      for (int bj = 1; bj < nsub; bj++) {
        const int na = bj * 8, cw = min(8, jb - na);
        for (int idx = tid; idx < na * cw; idx += NTH) {
          const int a = idx % na, cp = idx / na;
          double acc = 0.;
          for (int b = a; b < na; b++) acc += Ts[a + b * LDW] * Ws[b + (na + cp) * LDW];
          Ys[a + cp * LDW] = acc;
        }
        __syncthreads();
        for (int idx = tid; idx < na * cw; idx += NTH) {
          const int a = idx % na, cp = idx / na;
          double acc = 0.;
          for (int d = 0; d <= cp; d++) acc += Ys[a + d * LDW] * Ts[(na + d) + (na + cp) * LDW];
          Ts[a + (na + cp) * LDW] = -acc;
        }
        __syncthreads();
      }
    }
    QR_TICK(4)
    for (int idx = tid; idx < NB * NB; idx += NTH) {
      const int a = idx % NB, c = idx / NB;
      if (a < jb && c < jb) Tg[a + (size_t)(j0 + c) * NB] = Ts[a + c * LDW];
    }
    // ---- V (strictly lower part) back to global
    for (int c = warp; c < jb; c += NWP) {
      double* dst = A + j0 + (size_t)(j0 + c) * m;
      const double* src = Vs + c * ldv;
      for (int i = c + 1 + lane; i < mp; i += 32) dst[i] = src[i];
    }
    QR_TICK(5)
    __syncthreads();
    // ---- trailing update: one warp per 8-column slab, no barrier inside
    const int ntrail = naug - (j0 + jb);
    // (the warp that gets the extra slab rotates with the panel index)
    if (wide) {
      // 16-byte aligned columns: two slabs per warp, 128-bit C and V accesses
      // the slabs are dealt out in contiguous runs whose lengths differ by at most one (pairs of slabs
      // dealt round-robin leave a warp with up to two slabs more than another: 13 % of the update time at
      // m = 256); a run is worked off in pairs, its odd slab alone
      const int nslab = (ntrail + 7) >> 3;
      const int wr = (warp + (j0 / NB) * 3) % NWP;
      const int base = nslab / NWP, rem = nslab - base * NWP;
      int sl = wr * base + min(wr, rem);
      const int send = sl + base + (wr < rem ? 1 : 0);
      for (; sl + 1 < send; sl += 2) {
        const int c0 = j0 + jb + sl * 8;
        slab2_update<NB>(A + j0 + (size_t)c0 * m, m, mp, 8, min(8, naug - c0 - 8), Vs, ldv, Ts, LDW, lane);
      }
      if (sl < send) {
        const int c0 = j0 + jb + sl * 8;
        slab2_update<NB, false>(A + j0 + (size_t)c0 * m, m, mp, min(8, naug - c0), 0, Vs, ldv, Ts, LDW, lane);
      }
    } else {
      for (int sl = (warp + (j0 / NB) * 3) % NWP; sl * 8 < ntrail; sl += NWP) {
        const int c0 = j0 + jb + sl * 8;
        slab_update<NB, (NB == 16 ? 8 : 4)>(A + j0 + (size_t)c0 * m, m, mp, min(8, naug - c0), Vs, ldv, Ts, LDW, lane);
      }
    }
    QR_TICK(6)
    __syncthreads();
    QR_TICK(7)
  }
#ifdef SB200_QR_TIMING
  if (lane == 0 && blockIdx.x < 4) {
    printf("[qr timing] block %d warp %d m %d : load %lld steps %lld fin %lld inpanel %lld merge %lld wb %lld trail %lld wait %lld | pass %lld shfl %lld scalars %lld update %lld other %lld barrier %lld tcol %lld\n",
           blockIdx.x, warp, m, tph[0], tph[1], tph[2], tph[3], tph[4], tph[5], tph[6], tph[7],
           tph[8], tph[9], tph[10], tph[11], tph[12], tph[13], tph[14]);
  }
#endif
}

// Root: Dfull (already assembled in the root factor block by
// ulv_build_inner_kernel, or D itself for a single-node tree) -> LU with
// partial pivoting, in place.                       (factor.hpp:104-107)
__global__ void __launch_bounds__(kThreads)
ulv_root_lu_kernel(double* __restrict__ Ag, int n, int* __restrict__ piv, int use_smem) {
  extern __shared__ double smA[];
  __shared__ double rv[kWarps];
  __shared__ int ri[kWarps];
  __shared__ int pivrow;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // the reduced system at the root is small (sum of two ranks): factor it in
  // shared memory when it fits.  Every column is a latency chain; two barriers
  // per column: each warp finds the pivot itself (same data, same tie-break: no
  // broadcast), rows are swapped across all columns, [barrier], the trailing
  // block gets its rank-1 update with the multipliers formed on the fly,
  // [barrier].  The column scalings 1/pivot are applied in one pass at the end
  // (they commute with the later row interchanges).
  double* A = use_smem ? smA : Ag;
  if (use_smem) {
    double* invp = smA + (size_t)n * n;   // n reciprocal pivots
    for (int idx = tid; idx < n * n; idx += kThreads) A[idx] = Ag[idx];
    __syncthreads();
    for (int j = 0; j < n; j++) {
      double best = -1.;
      int bi = j;
      for (int i = j + lane; i < n; i += 32) {
        const double v = fabs(A[i + (size_t)j * n]);
        if (v > best) { best = v; bi = i; }
      }
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
      }
      const int p = bi;
      const double d = A[p + (size_t)j * n];
      const double inv = d != 0. ? 1. / d : 0.;
      if (tid == 0) { piv[j] = p; invp[j] = inv; }
      if (p != j) {        // uniform: every warp found the same pivot
        __syncthreads();   // every warp has read column j
        for (int c = tid; c < n; c += kThreads) {
          const double t = A[j + (size_t)c * n];
          A[j + (size_t)c * n] = A[p + (size_t)c * n];
          A[p + (size_t)c * n] = t;
        }
        __syncthreads();
      }
      const int nt = n - j - 1;
      for (int idx = tid; idx < nt * nt; idx += kThreads) {
        const int i = j + 1 + idx % nt, c = j + 1 + idx / nt;
        A[i + (size_t)c * n] -= (A[i + (size_t)j * n] * inv) * A[j + (size_t)c * n];
      }
      __syncthreads();
    }
    for (int idx = tid; idx < n * n; idx += kThreads) {
      const int i = idx % n, c = idx / n;
      Ag[idx] = i > c ? A[idx] * invp[c] : A[idx];
    }
    return;
  }
  for (int j = 0; j < n; j++) {
    // pivot search in column j
    double best = -1.;
    int bi = j;
    for (int i = j + tid; i < n; i += kThreads) {
      double v = fabs(A[i + (size_t)j * n]);
      if (v > best) { best = v; bi = i; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      double ov = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) { rv[warp] = best; ri[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
      double b = rv[0]; int p = ri[0];
      for (int w = 1; w < kWarps; w++)
        if (rv[w] > b || (rv[w] == b && ri[w] < p)) { b = rv[w]; p = ri[w]; }
      pivrow = p;
      piv[j] = p;
    }
    __syncthreads();
    const int p = pivrow;
    if (p != j)
      for (int c = tid; c < n; c += kThreads) {
        double t = A[j + (size_t)c * n];
        A[j + (size_t)c * n] = A[p + (size_t)c * n];
        A[p + (size_t)c * n] = t;
      }
    __syncthreads();
    const double d = A[j + (size_t)j * n];
    const double inv = d != 0. ? 1. / d : 0.;
    // scale the column and update the trailing block in one pass (each thread
    // recomputes the multiplier it needs)
    const int nt = n - j - 1;
    for (int idx = tid; idx < nt * nt; idx += kThreads) {
      int i = j + 1 + idx % nt, c = j + 1 + idx / nt;
      A[i + (size_t)c * n] -= (A[i + (size_t)j * n] * inv) * A[j + (size_t)c * n];
    }
    __syncthreads();
    for (int i = j + 1 + tid; i < n; i += kThreads) A[i + (size_t)j * n] *= inv;
    __syncthreads();
  }
  if (use_smem)
    for (int idx = tid; idx < n * n; idx += kThreads) Ag[idx] = A[idx];
}

__global__ void copy_block_kernel(const double* __restrict__ src, double* __restrict__ dst, long long n) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

// ===========================================================================
//                                ULV SOLVE
// ===========================================================================
// Forward sweep over one height class (root excluded)   solve.hpp:69-197
// pf: prefetch the next column block (and E, and at the end the extra columns)
// into L2 while the current one is processed (experiment, SB200_SOLVE_PIPE bit 3)
template <int NB>
__global__ void __launch_bounds__(kThreads)
ulv_fwd_kernel(const DNode* __restrict__ nodes, const int* __restrict__ list,
               const double* __restrict__ vals, const int* __restrict__ perms,
               const double* __restrict__ fact, const double* __restrict__ b, int ldb,
               double* __restrict__ ysol, double* __restrict__ zsol,
               double* __restrict__ fsol, int s, int pf = 0) {
  extern __shared__ __align__(16) double sm[];
  const DNode nd = nodes[list[blockIdx.x]];
  if (nd.parent < 0) return;
  const int col = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m = nd.m, r = nd.u_rank, k = nd.k, rv = nd.v_rank;
  double* f = sm;            // m
  double* fp = f + m;        // m   (P^T f) ; fp[0:r] = ft1, then y in fp[r:]
  double* zc = fp + m;       // v_rows (children z concat) for inner nodes
  double* Rd = zc + max(nd.v_rows, 1);  // 32 x 33 diagonal block
  // lines [0, rows) of columns [c0, c0 + nc) of the factor block -> L2
  auto pf_cols = [&](int c0, int nc, int rows) {
    const double* Ab = fact + nd.F;
    for (int a = warp; a < nc; a += kWarps) {
      const double* src = Ab + (size_t)(c0 + a) * m;
      for (int i = lane * 16; i < rows; i += 32 * 16) prefetch_l2(src + i);
    }
  };
  if (pf && k > 0) {
    const double* E = vals + nd.Eu;
    for (long long i = (long long)tid * 16; i < (long long)k * r; i += kThreads * 16) prefetch_l2(E + i);
    pf_cols(0, min(32, k), min(32, k));
  }
  // ---- gather f
  if (nd.leaf) {
    const double* bb = b + nd.row_off + (size_t)col * ldb;
    for (int i = tid; i < m; i += kThreads) f[i] = bb[i];
  } else {
    const DNode c0 = nodes[nd.ch0], c1 = nodes[nd.ch1];
    const int ru0 = c0.u_rank, ru1 = c1.u_rank, rv0 = c0.v_rank, rv1 = c1.v_rank;
    const double* f0 = fsol + (size_t)c0.f_off * s + (size_t)col * ru0;
    const double* f1 = fsol + (size_t)c1.f_off * s + (size_t)col * ru1;
    const double* z0 = zsol + (size_t)c0.z_off * s + (size_t)col * rv0;
    const double* z1 = zsol + (size_t)c1.z_off * s + (size_t)col * rv1;
    const double* B01 = vals + nd.B01;
    const double* B10 = vals + nd.B10;
    for (int i = tid; i < m; i += kThreads) {
      double v;
      if (i < ru0) { v = f0[i]; for (int j = 0; j < rv1; j++) v -= B01[i + (size_t)j * ru0] * z1[j]; }
      else { int ii = i - ru0; v = f1[ii]; for (int j = 0; j < rv0; j++) v -= B10[ii + (size_t)j * ru1] * z0[j]; }
      f[i] = v;
    }
    for (int i = tid; i < rv0 + rv1; i += kThreads) zc[i] = i < rv0 ? z0[i] : z1[i - rv0];
  }
  __syncthreads();
  const int* P = perms + nd.Pu;
  for (int i = tid; i < m; i += kThreads) fp[i] = f[P[i]];
  __syncthreads();
  double* y = fp + r;
  const double* A = fact + nd.F;
  if (k > 0) {
    // rhs = fp[r:] - E ft1
    const double* E = vals + nd.Eu;
    for (int i = tid; i < k; i += kThreads) {
      double v = fp[r + i];
      for (int l = 0; l < r; l++) v -= E[i + (size_t)l * k] * fp[l];
      f[i] = v;  // reuse f as rhs
    }
    __syncthreads();
    for (int i = tid; i < k; i += kThreads) y[i] = f[i];
    __syncthreads();
    // y = R^{-T} rhs, blocked by 32 columns (L = R^T, solve.hpp:160-161)
    for (int i0 = 0; i0 < k; i0 += 32) {
      const int ib = min(32, k - i0);
      if (pf) {
        if (i0 + 32 < k) pf_cols(i0 + 32, min(32, k - i0 - 32), min(k, i0 + 64));
        if (i0 + 64 >= k) pf_cols(k + (i0 + 32 >= k ? (rv + r) / 2 : 0), (rv + r + 1) / 2, k);   // the extra columns, in two halves
      }
      for (int c = warp; c < ib; c += kWarps) {
        const double* Rc = A + (size_t)(i0 + c) * m;
        double acc = 0.;
        for (int j = lane; j < i0; j += 32) acc += Rc[j] * y[j];
        acc = warp_sum(acc);
        if (lane == 0) y[i0 + c] -= acc;
      }
      for (int idx = tid; idx < ib * ib; idx += kThreads) {
        int a = idx % ib, c = idx / ib;
        Rd[a + c * 33] = (a <= c) ? A[(i0 + a) + (size_t)(i0 + c) * m] : 0.;
      }
      __syncthreads();
      if (warp == 0) {
        double val = lane < ib ? y[i0 + lane] : 0.;
        for (int a = 0; a < ib; a++) {
          if (lane == a) val = val / Rd[a + a * 33];
          const double ya = __shfl_sync(0xffffffffu, val, a);
          if (lane > a && lane < ib) val -= Rd[a + lane * 33] * ya;
        }
        if (lane < ib) y[i0 + lane] = val;
      }
      __syncthreads();
    }
    double* yo = ysol + (size_t)nd.y_off * s + (size_t)col * k;
    for (int i = tid; i < k; i += kThreads) yo[i] = y[i];
  }
  // ---- z = Vt0^H y (+ V^H [z0; z1]) ; ft1 -= (W1 Q0^H) y
  double* zo = zsol + (size_t)nd.z_off * s + (size_t)col * rv;
  double* fo = fsol + (size_t)nd.f_off * s + (size_t)col * r;
  const int* Pv = perms + nd.Pv;
  const double* Ev = vals + nd.Ev;
  const int kv = nd.v_rows - rv;
  for (int c = warp; c < rv + r; c += kWarps) {
    double acc = 0.;
    if (k > 0) {
      const double* col_ = A + (size_t)(k + c) * m;
      for (int i = lane; i < k; i += 32) acc += col_[i] * y[i];
    }
    if (c < rv && !nd.leaf) {
      // basis part: zc[Pv[c]] + sum_i Ev[i,c] zc[Pv[rv+i]]   (solve.hpp:166-167)
      for (int i = lane; i < kv; i += 32) acc += Ev[i + (size_t)c * kv] * zc[Pv[rv + i]];
      if (lane == 0) acc += zc[Pv[c]];
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      if (c < rv) zo[c] = acc;
      else fo[c - rv] = fp[c - rv] - acc;
    }
  }
}

// Root: f = [ft1(c0) - B01 z(c1); ft1(c1) - B10 z(c0)], x = LU^{-1} f
//                                                   (solve.hpp:88-135)
// `node` = 0 with lu = the root's factor block; for a partial factorization
// (partial_forward_solve) node = child 0 of the root and lu = its separate LU
// buffer, and x always goes to the solve workspace.
__global__ void __launch_bounds__(kThreads)
ulv_root_solve_kernel(const DNode* __restrict__ nodes, int node, const double* __restrict__ vals,
                      const double* __restrict__ lu, const int* __restrict__ piv,
                      double* __restrict__ b, int ldb, const double* __restrict__ zsol,
                      const double* __restrict__ fsol, double* __restrict__ xsol, int s,
                      int use_smem) {
  extern __shared__ __align__(16) double sm[];
  const DNode nd = nodes[node];
  const int col = blockIdx.x, tid = threadIdx.x;
  const int n = nd.m;
  double* x = sm;
  double* As = sm + ((n + 1) & ~1);   // LU factors staged in smem when they fit
  if (use_smem)
    for (int idx = tid; idx < n * n; idx += kThreads) As[idx] = lu[idx];
  if (nd.leaf) {
    const double* bb = b + nd.row_off + (size_t)col * ldb;
    for (int i = tid; i < n; i += kThreads) x[i] = bb[i];
  } else {
    const DNode c0 = nodes[nd.ch0], c1 = nodes[nd.ch1];
    const int ru0 = c0.u_rank, ru1 = c1.u_rank, rv0 = c0.v_rank, rv1 = c1.v_rank;
    const double* f0 = fsol + (size_t)c0.f_off * s + (size_t)col * ru0;
    const double* f1 = fsol + (size_t)c1.f_off * s + (size_t)col * ru1;
    const double* z0 = zsol + (size_t)c0.z_off * s + (size_t)col * rv0;
    const double* z1 = zsol + (size_t)c1.z_off * s + (size_t)col * rv1;
    const double* B01 = vals + nd.B01;
    const double* B10 = vals + nd.B10;
    for (int i = tid; i < n; i += kThreads) {
      double v;
      if (i < ru0) { v = f0[i]; for (int j = 0; j < rv1; j++) v -= B01[i + (size_t)j * ru0] * z1[j]; }
      else { int ii = i - ru0; v = f1[ii]; for (int j = 0; j < rv0; j++) v -= B10[ii + (size_t)j * ru1] * z0[j]; }
      x[i] = v;
    }
  }
  __syncthreads();
  const double* A = use_smem ? As : lu;
  if (tid == 0)
    for (int j = 0; j < n; j++) {
      int p = piv[j];
      if (p != j) { double t = x[j]; x[j] = x[p]; x[p] = t; }
    }
  __syncthreads();
  if (n <= 128) {
    // small system (the usual case: sum of two ranks): one warp runs both
    // substitutions with warp-level synchronisation only -- a chain of n short
    // steps instead of 3 n CTA barriers
    if (tid < 32) {
      for (int j = 0; j < n; j++) {   // unit lower
        const double xj = x[j];
        for (int i = j + 1 + tid; i < n; i += 32) x[i] -= A[i + (size_t)j * n] * xj;
        __syncwarp();
      }
      for (int j = n - 1; j >= 0; j--) {  // upper
        const double xj = x[j] / A[j + (size_t)j * n];
        __syncwarp();
        if (tid == 0) x[j] = xj;
        for (int i = tid; i < j; i += 32) x[i] -= A[i + (size_t)j * n] * xj;
        __syncwarp();
      }
    }
    __syncthreads();
  } else {
    for (int j = 0; j < n; j++) {   // unit lower
      const double xj = x[j];
      for (int i = j + 1 + tid; i < n; i += kThreads) x[i] -= A[i + (size_t)j * n] * xj;
      __syncthreads();
    }
    for (int j = n - 1; j >= 0; j--) {  // upper
      if (tid == 0) x[j] /= A[j + (size_t)j * n];
      __syncthreads();
      const double xj = x[j];
      for (int i = tid; i < j; i += kThreads) x[i] -= A[i + (size_t)j * n] * xj;
      __syncthreads();
    }
  }
  if (nd.leaf && node == 0) {
    double* bb = b + (size_t)col * ldb;
    for (int i = tid; i < n; i += kThreads) bb[i] = x[i];
  } else {
    double* xo = xsol + (size_t)nd.x_off * s + (size_t)col * n;
    for (int i = tid; i < n; i += kThreads) xo[i] = x[i];
  }
}

// Backward sweep over one height class (root excluded): x = Q^H [y; x_c]
//                                                   (solve.hpp:199-238)
template <int NB>
__global__ void __launch_bounds__(kThreads)
ulv_bwd_kernel(const DNode* __restrict__ nodes, const int* __restrict__ list,
               const double* __restrict__ fact, const double* __restrict__ tfac,
               double* __restrict__ b, int ldb, const double* __restrict__ ysol,
               double* __restrict__ xsol, int s, int pf = 0) {
  extern __shared__ __align__(16) double sm[];
  const int id = list[blockIdx.x];
  const DNode nd = nodes[id];
  if (nd.parent < 0) return;
  const DNode par = nodes[nd.parent];
  const int col = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m = nd.m, k = nd.k;
  double* v = sm;        // m
  double* w = v + m;     // NB
  double* w2 = w + NB;   // NB
  const int xoff = (par.ch0 == id) ? 0 : nodes[par.ch0].u_rank;
  const double* xp = xsol + (size_t)par.x_off * s + (size_t)col * par.m + xoff;
  const double* yi = ysol + (size_t)nd.y_off * s + (size_t)col * k;
  for (int i = tid; i < m; i += kThreads) v[i] = i < k ? yi[i] : xp[i - k];
  __syncthreads();
  const double* A = fact + nd.F;
  const double* Tg = tfac + nd.T;
  const int nbq = nd.nbq;          // panel width the node was factored with (<= NB)
  const int nblk = (k + nbq - 1) / nbq;
  // pf: pull the reflector block that comes next (and its T) into L2 while this one is applied
  auto pf_block = [&](int pb) {
    const int pj0 = pb * nbq, pjb = min(nbq, k - pj0);
    for (int a = warp; a < pjb; a += kWarps) {
      const double* src = A + pj0 + (size_t)(pj0 + a) * m;
      for (int i = lane * 16; i < m - pj0; i += 32 * 16) prefetch_l2(src + i);
    }
    if (tid < pjb) prefetch_l2(Tg + (size_t)(pj0 + tid) * nbq);
  };
  if (pf && nblk > 1) pf_block(nblk - 2);
  for (int bi = nblk - 1; bi >= 0; bi--) {
    const int j0 = bi * nbq, jb = min(nbq, k - j0);
    if (pf && bi > 1) pf_block(bi - 2);
    for (int a = warp; a < jb; a += kWarps) {
      const double* Va = A + (size_t)(j0 + a) * m;
      double acc = 0.;
      for (int i = j0 + a + 1 + lane; i < m; i += 32) acc += Va[i] * v[i];
      acc = warp_sum(acc);
      if (lane == 0) w[a] = acc + v[j0 + a];
    }
    __syncthreads();
    if (tid < jb) {
      double acc = 0.;
      for (int c = tid; c < jb; c++) acc += Tg[tid + (size_t)(j0 + c) * nbq] * w[c];
      w2[tid] = acc;
    }
    __syncthreads();
    for (int i = j0 + tid; i < m; i += kThreads) {
      double acc = 0.;
      const int amax = min(jb, i - j0);  // a with j0+a < i
      for (int a = 0; a < amax; a++) acc += A[i + (size_t)(j0 + a) * m] * w2[a];
      if (i - j0 < jb) acc += w2[i - j0];
      v[i] -= acc;
    }
    __syncthreads();
  }
  if (nd.leaf) {
    double* bb = b + nd.row_off + (size_t)col * ldb;
    for (int i = tid; i < m; i += kThreads) bb[i] = v[i];
  } else {
    double* xo = xsol + (size_t)nd.x_off * s + (size_t)col * m;
    for (int i = tid; i < m; i += kThreads) xo[i] = v[i];
  }
}

// ---------------------------------------------------------------------------
// Streaming variants of the two solve sweeps (single right-hand side per CTA).
// ulv_fwd_kernel / ulv_bwd_kernel walk the factor block in column blocks and
// every block starts with a dependent global load: ~15 exposed HBM round trips
// per leaf, 3-4x the time the bytes need.  None of those loads depends on the
// vector being solved, so here the factor block is streamed through shared
// memory by cp.async (LDGSTS) one column chunk AHEAD of its use, double
// buffered: the chain of barriers stays, the memory latency leaves it, and
// three CTAs per SM keep ~100 KB of loads in flight (HBM-bound by design:
// the sweep reads the block once).  Same arithmetic as the kernels above.
// ---------------------------------------------------------------------------

// Backward sweep: x = Q^H [y; x_c], reflector blocks of `nbq` columns (the
// panel width of the class, <= nbw) applied last to first.
__global__ void __launch_bounds__(kThreads)
ulv_bwd_pipe_kernel(const DNode* __restrict__ nodes, const int* __restrict__ list,
                    const double* __restrict__ fact, const double* __restrict__ tfac,
                    double* __restrict__ b, int ldb, const double* __restrict__ ysol,
                    double* __restrict__ xsol, int s, int ldbuf, int nbw) {
  extern __shared__ __align__(16) double sm[];
  const int id = list[blockIdx.x];
  const DNode nd = nodes[id];
  if (nd.parent < 0) return;
  const DNode par = nodes[nd.parent];
  const int col = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m = nd.m, k = nd.k;
  double* v = sm;                              // ldbuf
  double* w = v + ldbuf;                       // 32
  double* w2 = w + 32;                         // 32
  double* Tb = w2 + 32;                        // 2 x (nbw x nbw)
  double* Vb = Tb + 2 * nbw * nbw;             // 2 x (ldbuf x nbw)
  const int xoff = (par.ch0 == id) ? 0 : nodes[par.ch0].u_rank;
  const double* xp = xsol + (size_t)par.x_off * s + (size_t)col * par.m + xoff;
  const double* yi = ysol + (size_t)nd.y_off * s + (size_t)col * k;
  const double* A = fact + nd.F;
  const double* Tg = tfac + nd.T;
  const int nbq = nd.nbq;
  const int nblk = (k + nbq - 1) / nbq;
  // rows j0..m-1 of the block's columns (the reflector tails; the entries on and
  // above the diagonal are R and are masked out below) + its T block
  auto issue = [&](int bi, int buf) {
    const int j0 = bi * nbq, jb = min(nbq, k - j0), rows = m - j0;
    double* dst = Vb + (size_t)buf * ldbuf * nbw;
    for (int a = warp; a < jb; a += kWarps) {
      const double* src = A + j0 + (size_t)(j0 + a) * m;
      for (int i = lane; i < rows; i += 32) cp_async8(dst + i + a * ldbuf, src + i, true);
    }
    double* td = Tb + buf * nbw * nbw;
    for (int idx = tid; idx < jb * nbw; idx += kThreads) {
      const int r = idx % nbw, c = idx / nbw;
      if (r < jb) cp_async8(td + r + c * nbw, Tg + r + (size_t)(j0 + c) * nbq, true);
    }
    cp_async_commit();
  };
  if (nblk > 0) issue(nblk - 1, 0);            // in flight while v is gathered
  for (int i = tid; i < m; i += kThreads) v[i] = i < k ? yi[i] : xp[i - k];
  int buf = 0;
  for (int bi = nblk - 1; bi >= 0; bi--, buf ^= 1) {
    if (bi > 0) { issue(bi - 1, buf ^ 1); cp_async_wait<1>(); }
    else cp_async_wait<0>();
    __syncthreads();
    const int j0 = bi * nbq, jb = min(nbq, k - j0), rows = m - j0;
    const double* Vc = Vb + (size_t)buf * ldbuf * nbw;
    const double* Tc = Tb + buf * nbw * nbw;
    for (int a = warp; a < jb; a += kWarps) {
      const double* Va = Vc + a * ldbuf;
      double acc = 0.;
      for (int i = a + 1 + lane; i < rows; i += 32) acc += Va[i] * v[j0 + i];
      acc = warp_sum(acc);
      if (lane == 0) w[a] = acc + v[j0 + a];
    }
    __syncthreads();
    if (tid < jb) {
      double acc = 0.;
      for (int c = tid; c < jb; c++) acc += Tc[tid + c * nbw] * w[c];
      w2[tid] = acc;
    }
    __syncthreads();
    for (int i = tid; i < rows; i += kThreads) {
      double acc = 0.;
      const int amax = min(jb, i);       // columns a with j0 + a < j0 + i
      for (int a = 0; a < amax; a++) acc += Vc[i + a * ldbuf] * w2[a];
      if (i < jb) acc += w2[i];
      v[j0 + i] -= acc;
    }
    __syncthreads();
  }
  __syncthreads();
  if (nd.leaf) {
    double* bb = b + nd.row_off + (size_t)col * ldb;
    for (int i = tid; i < m; i += kThreads) bb[i] = v[i];
  } else {
    double* xo = xsol + (size_t)nd.x_off * s + (size_t)col * m;
    for (int i = tid; i < m; i += kThreads) xo[i] = v[i];
  }
}

// Forward sweep: the first k rows of the factor block are streamed in chunks
// of kFC columns: chunks inside [0, k) are steps of the triangular solve
// y = R^{-T} rhs, the chunks of the last r_v + r_u columns are the products
// z = Vt0^H y and ft1 -= (W1 Q0^H) y.
constexpr int kFC = 16;
__global__ void __launch_bounds__(kThreads)
ulv_fwd_pipe_kernel(const DNode* __restrict__ nodes, const int* __restrict__ list,
                    const double* __restrict__ vals, const int* __restrict__ perms,
                    const double* __restrict__ fact, const double* __restrict__ b, int ldb,
                    double* __restrict__ ysol, double* __restrict__ zsol,
                    double* __restrict__ fsol, int s, int ldm, int ldbuf) {
  extern __shared__ __align__(16) double sm[];
  const DNode nd = nodes[list[blockIdx.x]];
  if (nd.parent < 0) return;
  const int col = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m = nd.m, r = nd.u_rank, k = nd.k, rv = nd.v_rank;
  double* f = sm;              // ldm
  double* fp = f + ldm;        // ldm  (P^T f) ; fp[0:r] = ft1, then y in fp[r:]
  double* zc = fp + ldm;       // ldm  children z concat (inner nodes)
  double* Cb = zc + ldm;       // 2 x (ldbuf x kFC)
  const double* A = fact + nd.F;
  const int nextra = rv + r;
  const int nsolve = (k + kFC - 1) / kFC;                // chunks of the triangular solve
  const int nchunk = k > 0 ? nsolve + (nextra + kFC - 1) / kFC : 0;
  // chunk c: columns [c0, c0 + cw) of the factor block, rows [0, nr)
  auto issue = [&](int c, int buf) {
    int c0, cw, nr;
    if (c < nsolve) { c0 = c * kFC; cw = min(kFC, k - c0); nr = c0 + cw; }
    else { c0 = k + (c - nsolve) * kFC; cw = min(kFC, k + nextra - c0); nr = k; }
    double* dst = Cb + (size_t)buf * ldbuf * kFC;
    for (int a = warp; a < cw; a += kWarps) {
      const double* src = A + (size_t)(c0 + a) * m;
      for (int i = lane; i < nr; i += 32) cp_async8(dst + i + a * ldbuf, src + i, true);
    }
    cp_async_commit();
  };
  if (nchunk > 0) issue(0, 0);       // in flight while the right-hand side is assembled
  // ---- gather f
  if (nd.leaf) {
    const double* bb = b + nd.row_off + (size_t)col * ldb;
    for (int i = tid; i < m; i += kThreads) f[i] = bb[i];
  } else {
    const DNode c0 = nodes[nd.ch0], c1 = nodes[nd.ch1];
    const int ru0 = c0.u_rank, ru1 = c1.u_rank, rv0 = c0.v_rank, rv1 = c1.v_rank;
    const double* f0 = fsol + (size_t)c0.f_off * s + (size_t)col * ru0;
    const double* f1 = fsol + (size_t)c1.f_off * s + (size_t)col * ru1;
    const double* z0 = zsol + (size_t)c0.z_off * s + (size_t)col * rv0;
    const double* z1 = zsol + (size_t)c1.z_off * s + (size_t)col * rv1;
    const double* B01 = vals + nd.B01;
    const double* B10 = vals + nd.B10;
    for (int i = tid; i < m; i += kThreads) {
      double v;
      if (i < ru0) { v = f0[i]; for (int j = 0; j < rv1; j++) v -= B01[i + (size_t)j * ru0] * z1[j]; }
      else { int ii = i - ru0; v = f1[ii]; for (int j = 0; j < rv0; j++) v -= B10[ii + (size_t)j * ru1] * z0[j]; }
      f[i] = v;
    }
    for (int i = tid; i < rv0 + rv1; i += kThreads) zc[i] = i < rv0 ? z0[i] : z1[i - rv0];
  }
  __syncthreads();
  const int* P = perms + nd.Pu;
  for (int i = tid; i < m; i += kThreads) fp[i] = f[P[i]];
  __syncthreads();
  double* y = fp + r;
  if (k > 0) {
    // rhs = fp[r:] - E ft1
    const double* E = vals + nd.Eu;
    for (int i = tid; i < k; i += kThreads) {
      double v = fp[r + i];
      for (int l = 0; l < r; l++) v -= E[i + (size_t)l * k] * fp[l];
      f[i] = v;  // reuse f as rhs
    }
    __syncthreads();
    for (int i = tid; i < k; i += kThreads) y[i] = f[i];
    __syncthreads();
  }
  double* zo = zsol + (size_t)nd.z_off * s + (size_t)col * rv;
  double* fo = fsol + (size_t)nd.f_off * s + (size_t)col * r;
  const int* Pv = perms + nd.Pv;
  const double* Ev = vals + nd.Ev;
  const int kv = nd.v_rows - rv;
  int buf = 0;
  for (int c = 0; c < nchunk; c++, buf ^= 1) {
    if (c + 1 < nchunk) { issue(c + 1, buf ^ 1); cp_async_wait<1>(); }
    else cp_async_wait<0>();
    __syncthreads();
    const double* Cc = Cb + (size_t)buf * ldbuf * kFC;
    if (c < nsolve) {
      // subtract the contribution of the solved prefix, then a cw-step
      // substitution with the diagonal block (L = R^T, solve.hpp:160-161)
      const int c0 = c * kFC, cw = min(kFC, k - c0);
      for (int a = warp; a < cw; a += kWarps) {
        const double* Ra = Cc + a * ldbuf;
        double acc = 0.;
        for (int j = lane; j < c0; j += 32) acc += Ra[j] * y[j];
        acc = warp_sum(acc);
        if (lane == 0) y[c0 + a] -= acc;
      }
      __syncthreads();
      if (warp == 0) {
        // the substitution is a chain of cw dependent steps: keep it to
        // multiply + shuffle + fma by taking the reciprocal pivots and this
        // lane's column of the diagonal block into registers beforehand
        const bool lin = lane < cw;
        double val = lin ? y[c0 + lane] : 0.;
        const double* Rl = Cc + c0 + (lin ? lane : 0) * ldbuf;     // column c0 + lane, rows c0..
        const double dinv = lin ? 1. / Rl[lane] : 0.;
        double rc[kFC];
#pragma unroll
        for (int a = 0; a < kFC; a++) rc[a] = (lin && a < lane) ? Rl[a] : 0.;
#pragma unroll
        for (int a = 0; a < kFC; a++) {
          if (lane == a) val *= dinv;
          const double ya = __shfl_sync(0xffffffffu, val, a);
          val -= rc[a] * ya;          // rc[a] = 0 unless a < lane < cw
        }
        if (lin) y[c0 + lane] = val;
      }
      __syncthreads();
    } else {
      // columns k + e of the factor block: z = Vt0^H y (+ V^H [z0; z1]) ; ft1 -= (W1 Q0^H) y
      const int e0 = (c - nsolve) * kFC, cw = min(kFC, nextra - e0);
      for (int a = warp; a < cw; a += kWarps) {
        const int e = e0 + a;
        const double* Xa = Cc + a * ldbuf;
        double acc = 0.;
        for (int i = lane; i < k; i += 32) acc += Xa[i] * y[i];
        if (e < rv && !nd.leaf) {
          for (int i = lane; i < kv; i += 32) acc += Ev[i + (size_t)e * kv] * zc[Pv[rv + i]];
          if (lane == 0) acc += zc[Pv[e]];
        }
        acc = warp_sum(acc);
        if (lane == 0) {
          if (e < rv) zo[e] = acc;
          else fo[e - rv] = fp[e - rv] - acc;
        }
      }
      __syncthreads();     // the buffer is refilled two chunks later
    }
  }
  if (k > 0) {
    double* yo = ysol + (size_t)nd.y_off * s + (size_t)col * k;
    for (int i = tid; i < k; i += kThreads) yo[i] = y[i];
  } else {
    // nothing to eliminate: z = V^H [z0; z1] (or 0 at a leaf), ft1 = P^T f
    for (int e = warp; e < nextra; e += kWarps) {
      double acc = 0.;
      if (e < rv && !nd.leaf) {
        for (int i = lane; i < kv; i += 32) acc += Ev[i + (size_t)e * kv] * zc[Pv[rv + i]];
        if (lane == 0) acc += zc[Pv[e]];
      }
      acc = warp_sum(acc);
      if (lane == 0) {
        if (e < rv) zo[e] = acc;
        else fo[e - rv] = fp[e - rv] - acc;
      }
    }
  }
}

// ===========================================================================
//            SCHUR COMPLEMENT OF THE (0,0) BLOCK / PARTIAL FACTORIZATION
// ===========================================================================
// Building blocks of partial_factor / Schur_update / Schur_product_* (reference
// HSSMatrix.factor.hpp:44-50, HSSMatrix.Schur.hpp:35-215): the tree sweeps are
// the apply / ULV kernels above run over the node lists of the two subtrees;
// what is left are small dense products with the reduced blocks.

// C = alpha op(A) op(B) + beta C, column-major, any sizes: one CTA per 128 x 64
// tile of C on the fp64 tensor pipe (smem_gemm reads its operands through
// predicated loads, so they may live in global memory).
template <bool TA, bool TB>
__global__ void __launch_bounds__(kThreads)
dense_gemm_kernel(int M, int N, int K, double alpha, const double* A, int lda,
                  const double* B, int ldb, double beta, double* C, int ldc) {
  const int i0 = blockIdx.x * 128, j0 = blockIdx.y * 64;
  const int mm = min(128, M - i0), nn = min(64, N - j0);
  const double* Ab = TA ? A + (size_t)i0 * lda : A + i0;
  const double* Bb = TB ? B + j0 : B + (size_t)j0 * ldb;
  smem_gemm<TA, TB>(mm, nn, K, alpha, Ab, lda, Bb, ldb, beta, C + i0 + (size_t)j0 * ldc, ldc,
                    threadIdx.x >> 5, kWarps, threadIdx.x & 31);
}

// Y = U X with U = P_u [I; E_u] of `node` (u_rows x u_rank), X u_rank x ncols
//                                       (HSSBasisID::apply, HSSBasisID.hpp:155-169)
__global__ void __launch_bounds__(kThreads)
basis_apply_kernel(const DNode* __restrict__ nodes, int node, const double* __restrict__ vals,
                   const int* __restrict__ perms, const double* __restrict__ X, int ldx,
                   int ncols, double* __restrict__ Y, int ldy) {
  const DNode nd = nodes[node];
  const int n = nd.u_rows, r = nd.u_rank, k = n - r;
  const int* P = perms + nd.Pu;
  const double* E = vals + nd.Eu;
  const long long tot = (long long)n * ncols;
  for (long long idx = blockIdx.x * (long long)kThreads + threadIdx.x; idx < tot;
       idx += (long long)gridDim.x * kThreads) {
    const int i = (int)(idx % n), c = (int)(idx / n);
    const double* x = X + (size_t)c * ldx;
    double v;
    if (i < r) v = x[i];
    else {
      v = 0.;
      for (int l = 0; l < r; l++) v += E[(i - r) + (size_t)l * k] * x[l];
    }
    Y[P[i] + (size_t)c * ldy] = v;
  }
}

// X <- (LU)^{-1} X, one CTA per column     (DenseMatrix::solve, DenseMatrix.cpp:625-640)
__global__ void __launch_bounds__(kThreads)
lu_solve_kernel(const double* __restrict__ A, int n, const int* __restrict__ piv,
                double* __restrict__ X, int ldx) {
  extern __shared__ __align__(16) double sm[];
  double* x = sm;
  const int tid = threadIdx.x;
  double* xg = X + (size_t)blockIdx.x * ldx;
  for (int i = tid; i < n; i += kThreads) x[i] = xg[i];
  __syncthreads();
  if (tid == 0)
    for (int j = 0; j < n; j++) {
      const int p = piv[j];
      if (p != j) { const double t = x[j]; x[j] = x[p]; x[p] = t; }
    }
  __syncthreads();
  if (n <= 128) {
    // small system (the usual case: sum of two ranks): one warp runs both
    // substitutions with warp-level synchronisation only -- a chain of n short
    // steps instead of 3 n CTA barriers
    if (tid < 32) {
      for (int j = 0; j < n; j++) {   // unit lower
        const double xj = x[j];
        for (int i = j + 1 + tid; i < n; i += 32) x[i] -= A[i + (size_t)j * n] * xj;
        __syncwarp();
      }
      for (int j = n - 1; j >= 0; j--) {  // upper
        const double xj = x[j] / A[j + (size_t)j * n];
        __syncwarp();
        if (tid == 0) x[j] = xj;
        for (int i = tid; i < j; i += 32) x[i] -= A[i + (size_t)j * n] * xj;
        __syncwarp();
      }
    }
    __syncthreads();
  } else {
    for (int j = 0; j < n; j++) {   // unit lower
      const double xj = x[j];
      for (int i = j + 1 + tid; i < n; i += kThreads) x[i] -= A[i + (size_t)j * n] * xj;
      __syncthreads();
    }
    for (int j = n - 1; j >= 0; j--) {  // upper
      if (tid == 0) x[j] /= A[j + (size_t)j * n];
      __syncthreads();
      const double xj = x[j];
      for (int i = tid; i < j; i += kThreads) x[i] -= A[i + (size_t)j * n] * xj;
      __syncthreads();
    }
  }
  for (int i = tid; i < n; i += kThreads) xg[i] = x[i];
}

// dst (cols x rows, ld ldd) = src^T (src rows x cols, ld lds)
__global__ void transpose_kernel(const double* __restrict__ src, int lds, int rows, int cols,
                                 double* __restrict__ dst, int ldd) {
  const long long tot = (long long)rows * cols;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < tot;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % rows), j = (int)(idx / rows);
    dst[j + (size_t)i * ldd] = src[i + (size_t)j * lds];
  }
}

// reduced_rhs = Vhat^H x + V^H [z(c0); z(c1)] at the sub-root of a partial
// factorization                                     (solve.hpp:136-152)
__global__ void __launch_bounds__(kThreads)
partial_reduced_rhs_kernel(const DNode* __restrict__ nodes, int node,
                           const double* __restrict__ vals, const int* __restrict__ perms,
                           const double* __restrict__ vhat, const double* __restrict__ xsol,
                           const double* __restrict__ zsol, double* __restrict__ red, int ldred,
                           int s) {
  extern __shared__ __align__(16) double sm[];
  double* zc = sm;
  const DNode nd = nodes[node];
  const int col = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m = nd.m, rv = nd.v_rank;
  const double* x = xsol + (size_t)nd.x_off * s + (size_t)col * m;
  if (!nd.leaf) {
    const DNode c0 = nodes[nd.ch0], c1 = nodes[nd.ch1];
    const int rv0 = c0.v_rank, rv1 = c1.v_rank;
    const double* z0 = zsol + (size_t)c0.z_off * s + (size_t)col * rv0;
    const double* z1 = zsol + (size_t)c1.z_off * s + (size_t)col * rv1;
    for (int i = tid; i < rv0 + rv1; i += kThreads) zc[i] = i < rv0 ? z0[i] : z1[i - rv0];
  }
  __syncthreads();
  const int* Pv = perms + nd.Pv;
  const double* Ev = vals + nd.Ev;
  const int kv = nd.v_rows - rv;
  for (int c = warp; c < rv; c += kWarps) {
    double acc = 0.;
    const double* vc = vhat + (size_t)c * m;
    for (int i = lane; i < m; i += 32) acc += vc[i] * x[i];
    if (!nd.leaf) {
      for (int i = lane; i < kv; i += 32) acc += Ev[i + (size_t)c * kv] * zc[Pv[rv + i]];
      if (lane == 0) acc += zc[Pv[c]];
    }
    acc = warp_sum(acc);
    if (lane == 0) red[c + (size_t)col * ldred] = acc;
  }
}

}  // namespace

// ===========================================================================
//                                 HOST SIDE
// ===========================================================================

HSSEngine::HSSEngine(HSSHost&& host) : H_(std::move(host)) {
  int dev = 0;
  SB200_CUDA(cudaGetDevice(&dev));
  SB200_CUDA(cudaDeviceGetAttribute(&nsm_, cudaDevAttrMultiProcessorCount, dev));
  if (const char* e = std::getenv("SB200_QR_REGPANEL")) qr_regpanel_ = std::atoi(e);
  // classes with m <= 256: 0 = 32-column panels, 256 threads, 2 CTAs/SM;
  // 1 = 16-column panels, 128 threads, 4 CTAs/SM; 2 = 16-column panels, 256 threads
  if (const char* e = std::getenv("SB200_QR_VARIANT")) qr_variant_ = std::atoi(e);
  if (const char* e = std::getenv("SB200_APPLY_MM_MIN")) mm_min_ = std::max(1, std::atoi(e));   // rhs count from which the GEMM-shaped apply kernels run
  if (const char* e = std::getenv("SB200_QR_NOWIDE")) qr_nowide_ = std::atoi(e);   // 1: 64-bit one-slab trailing update
  if (const char* e = std::getenv("SB200_ELIM_VARIANT")) elim_variant_ = std::atoi(e);   // see factor_classes
  // cp.async-streamed solve sweeps: bit 0 backward, bit 1 forward (default: both);
  // bits 2 / 3 (with bit 0 / 1 clear): the non-streamed kernels with L2 prefetch of the next block
  if (const char* e = std::getenv("SB200_SOLVE_PIPE")) solve_pipe_ = std::atoi(e);
  if (const char* e = std::getenv("SB200_GRAPH")) use_graph_ = std::atoi(e);
  // 1: left-looking warp-specialised TMA-fed QR (ulv_qr3.cuh) for classes with m <= 256; default: the right-looking kernel
  if (const char* e = std::getenv("SB200_QR3")) qr3_ = std::atoi(e);
  build_tables();
}
HSSEngine::~HSSEngine() {
  for (auto& e : ev_) if (e) cudaEventDestroy(e);
  drop_graphs();
  dist_close();
}

void HSSEngine::drop_graphs() {
  for (auto& g : graphs_) if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
  graphs_.clear();
  seen_.clear();
}

// Run `body` (a fixed sequence of kernel launches on `st`) through a CUDA graph:
// captured the first time a key is seen, replayed afterwards.  Not used on the
// legacy default stream (capture is not allowed there) nor while the per-kernel
// profiling events are on.
void HSSEngine::run_graphed(const std::string& key, cudaStream_t st, const std::function<void()>& body) {
  if (!use_graph_ || profile_ || st == nullptr || st == cudaStreamLegacy || st == cudaStreamPerThread) {
    body();
    return;
  }
  auto it = graphs_.find(key);
  if (it == graphs_.end() && seen_[key]++ == 0) {   // first time: plain launches (one-off calls never pay for a capture;
    body();                                         // libraries inside the sequence get their lazy set-up done)
    return;
  }
  if (it == graphs_.end()) {
    if (graphs_.size() > 64) drop_graphs();   // operands keep changing: do not hoard executables
    cudaGraph_t g = nullptr;
    const long long l0 = launches_;
    SB200_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    try {
      body();
    } catch (...) {
      cudaStreamEndCapture(st, &g);
      if (g) cudaGraphDestroy(g);
      throw;
    }
    SB200_CUDA(cudaStreamEndCapture(st, &g));
    GraphEntry e;
    e.launches = launches_ - l0;
    launches_ = l0;
    const cudaError_t rc = cudaGraphInstantiate(&e.exec, g, 0);
    cudaGraphDestroy(g);
    if (rc != cudaSuccess) { cudaGetLastError(); body(); return; }   // fall back to plain launches
    it = graphs_.emplace(key, e).first;
  }
  SB200_CUDA(cudaGraphLaunch(it->second.exec, st));
  launches_ += it->second.launches;
}

void HSSEngine::set_profile(bool on) {
  profile_ = on;
  if (on && !ev_[0]) {
    SB200_CUDA(cudaEventCreate(&ev_[0]));
    SB200_CUDA(cudaEventCreate(&ev_[1]));
  }
}

float HSSEngine::qr_leaf_ms() {
  if (!ev_[0]) return 0.f;
  float ms = 0.f;
  SB200_CUDA(cudaEventSynchronize(ev_[1]));
  SB200_CUDA(cudaEventElapsedTime(&ms, ev_[0], ev_[1]));
  return ms;
}

void HSSEngine::build_tables() {
  const int N = int(H_.nodes.size());
  hn_.resize(N);
  int woff = 0;
  long long foff = 0, toff = 0;
  int yoff = 0, zoff = 0, fo = 0, xo = 0;
  int maxm = 0;
  for (int i = 0; i < N; i++) {
    const auto& n = H_.nodes[i];
    maxm = std::max(maxm, n.leaf() ? std::max(n.rows, n.cols)
                                   : H_.nodes[n.ch0].u_rank + H_.nodes[n.ch1].u_rank);
  }
  nb_ = maxm <= 768 ? 32 : 16;
  class_gmm_.assign(H_.hptr.size(), 0);
  for (int i = 0; i < N; i++) {
    const auto& n = H_.nodes[i];
    const int mi = n.leaf() ? std::max(n.rows, n.cols)
                            : H_.nodes[n.ch0].u_rank + H_.nodes[n.ch1].u_rank;
    int& g = class_gmm_[n.height];
    g = std::max(g, std::max(mi, std::max(n.u_rows, n.v_rows)));
    if (!n.leaf()) g = std::max(g, H_.nodes[n.ch0].v_rank + H_.nodes[n.ch1].v_rank);
  }
  if (maxm > 1600)
    throw std::invalid_argument("HSS block larger than 1600 not supported");
  for (int i = 0; i < N; i++) {
    const auto& n = H_.nodes[i];
    DNode& d = hn_[i];
    d.ch0 = n.ch0; d.ch1 = n.ch1; d.parent = n.parent; d.leaf = n.leaf();
    d.rows = n.rows; d.cols = n.cols; d.row_off = n.row_off; d.col_off = n.col_off;
    d.u_rows = n.u_rows; d.u_rank = n.u_rank; d.v_rows = n.v_rows; d.v_rank = n.v_rank;
    d.D = n.off_D; d.Eu = n.off_Eu; d.Ev = n.off_Ev; d.B01 = n.off_B01; d.B10 = n.off_B10;
    d.Pu = n.off_Pu; d.Pv = n.off_Pv;
    d.w_off = woff;
    woff += std::max(n.u_rank, n.v_rank);
    d.m = n.leaf() ? n.rows : H_.nodes[n.ch0].u_rank + H_.nodes[n.ch1].u_rank;
    if (i == 0) {
      d.k = 0; d.naug = d.m;  // root: m x m LU block
    } else {
      d.k = d.m - n.u_rank;
      d.naug = d.k + n.v_rank + n.u_rank;
    }
    d.F = foff;
    foff += ((long long)d.m * d.naug + 1) & ~1LL;   // even offsets: 16-byte aligned columns when m is even
    d.T = toff;
    toff += (long long)nb_ * d.k;
    d.y_off = yoff; yoff += d.k;
    d.z_off = zoff; zoff += n.v_rank;
    d.f_off = fo; fo += n.u_rank;
    d.x_off = xo; xo += d.m;
  }
  ws_total_ = woff;
  tot_k_ = yoff; tot_rv_ = zoff; tot_ru_ = fo; tot_m_ = xo;
  fact_len_ = foff;
  fact_nnz_ = foff + toff;
  dn_.upload(hn_.data(), hn_.size());
  vals_.upload(H_.vals.data(), H_.vals.size());
  perms_.upload(H_.perms.data(), H_.perms.size());
  SB200_CUDA(cudaStreamSynchronize(0));
  set_partition(1, 0);
}

void HSSEngine::ensure_apply_ws(int s) {
  if (s <= apply_s_) return;
  drop_graphs();
  t1_.alloc((size_t)std::max(ws_total_, 1) * s);
  t2_.alloc((size_t)std::max(ws_total_, 1) * s);
  apply_s_ = s;
}

void HSSEngine::ensure_solve_ws(int s) {
  if (s <= solve_s_) return;
  drop_graphs();
  ysol_.alloc((size_t)std::max<long long>(tot_k_, 1) * s);
  zsol_.alloc((size_t)std::max<long long>(tot_rv_, 1) * s);
  fsol_.alloc((size_t)std::max<long long>(tot_ru_, 1) * s);
  xsol_.alloc((size_t)std::max<long long>(tot_m_, 1) * s);
  solve_s_ = s;
}

template <typename K> static void set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024)
    SB200_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}

// One TMA descriptor per node for ulv_qr3_kernel's reflector ring: the node's
// factor block as a 2-D fp64 tensor {m rows (contiguous), naug columns}, tiles of
// 16 x 16 written to shared memory with the 128-byte swizzle.  128 bytes per node
// in `out` (zeros for nodes the kernel does not feed through TMA).
static void build_qr3_tmaps(const std::vector<DNode>& nodes, double* fact_base, DevBuf<unsigned char>& out) {
  using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    SB200_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) throw std::runtime_error("cuTensorMapEncodeTiled not available");
    encode = reinterpret_cast<EncodeFn>(fn);
  }
  static_assert(sizeof(CUtensorMap) == 128, "tensor map size");
  std::vector<CUtensorMap> host(nodes.size());
  std::memset(host.data(), 0, host.size() * sizeof(CUtensorMap));
  for (size_t i = 0; i < nodes.size(); i++) {
    const DNode& d = nodes[i];
    if (d.parent < 0 || d.k == 0 || d.m > qr3::MAXM || (d.m & 1) || (d.F & 1)) continue;
    const cuuint64_t dims[2] = {(cuuint64_t)d.m, (cuuint64_t)d.naug};
    const cuuint64_t strides[1] = {(cuuint64_t)d.m * sizeof(double)};
    const cuuint32_t box[2] = {16, 16}, estr[2] = {1, 1};
    const CUresult r = encode(&host[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, fact_base + d.F, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
      throw std::runtime_error("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ") for an " +
                               std::to_string(d.m) + " x " + std::to_string(d.naug) + " factor block");
  }
  out.upload(reinterpret_cast<const unsigned char*>(host.data()), host.size() * sizeof(CUtensorMap));
  SB200_CUDA(cudaStreamSynchronize(0));
}

__global__ void copy2d_kernel(double* __restrict__ dst, long long ldd,
                              const double* __restrict__ src, long long lds,
                              int rows, int cols) {
  const long long tot = (long long)rows * cols;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < tot;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % rows), j = (int)(idx / rows);
    dst[i + j * ldd] = src[i + j * lds];
  }
}

static void copy2d(double* dst, long long ldd, const double* src, long long lds,
                   int rows, int cols, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return;
  const long long tot = (long long)rows * cols;
  copy2d_kernel<<<(unsigned)std::min<long long>((tot + 255) / 256, 1024), 256, 0, st>>>(dst, ldd, src, lds, rows, cols);
}

// ---------------------------------------------------------------------------
// node lists: the sweeps run over "owned" nodes (everything on one GPU; the
// subtree of this rank's cut node when the tree is sharded) and over the "top"
// nodes above the cut (replicated on every rank)
// ---------------------------------------------------------------------------
void HSSEngine::make_lists(NodeLists& L, const std::vector<int>& nodes) {
  int maxh = 0;
  for (int i : nodes) maxh = std::max(maxh, H_.nodes[i].height);
  L.hptr.assign(nodes.empty() ? 1 : maxh + 2, 0);
  for (int i : nodes) L.hptr[H_.nodes[i].height + 1]++;
  for (size_t h = 0; h + 1 < L.hptr.size(); h++) L.hptr[h + 1] += L.hptr[h];
  L.host.assign(nodes.size(), 0);
  std::vector<int> pos(L.hptr.begin(), L.hptr.end());
  for (int i : nodes) L.host[pos[H_.nodes[i].height]++] = i;   // `nodes` is in pre-order
  const int nh = L.classes();
  L.max_m.assign(nh, 0);
  L.max_k.assign(nh, 0);
  L.soff.assign(nodes.size(), 0);
  L.smax = 0;
  for (int h = 0; h < nh; h++) {
    long long o = 0;
    for (int q = L.hptr[h]; q < L.hptr[h + 1]; q++) {
      const DNode& d = hn_[L.host[q]];
      L.max_m[h] = std::max(L.max_m[h], std::max(d.m, std::max(d.rows * d.leaf, d.cols * d.leaf)));
      L.max_m[h] = std::max(L.max_m[h], std::max(d.u_rows, d.v_rows));
      if (!d.leaf) {
        // the apply kernels stage the children's t1 / t2 pieces (v_rank and
        // u_rank rows each) of an inner node in one buffer sized by max_m: the
        // V side can be larger than the U side for a nonsymmetric matrix
        const DNode& a = hn_[d.ch0];
        const DNode& b = hn_[d.ch1];
        L.max_m[h] = std::max(L.max_m[h], std::max(a.u_rank + b.u_rank, a.v_rank + b.v_rank));
      }
      L.max_k[h] = std::max(L.max_k[h], d.k);
      L.soff[q] = o;
      if (!d.leaf) o += (long long)d.m * d.m;
    }
    L.smax = std::max(L.smax, o);
    // panel width by the class maximum over the WHOLE tree, so that every list a
    // node appears in (owned / top / Schur subtrees) factors it the same way
    const int nbq = class_nb(h, std::max(class_gmm_[h], 1));
    for (int q = L.hptr[h]; q < L.hptr[h + 1]; q++) hn_[L.host[q]].nbq = nbq;
  }
  L.list.upload(L.host.data(), L.host.size());
  L.dsoff.upload(L.soff.data(), L.soff.size());
  SB200_CUDA(cudaStreamSynchronize(0));
}

int HSSEngine::class_nb(int h, int max_m) const {
  if (nb_ != 32 || max_m > 256) return nb_;
  if (qr3_) return 16;
  return (qr_regpanel_ && qr_variant_ >= 1) ? 16 : 32;
}

void HSSEngine::set_partition(int nparts, int part) {
  const int N = int(H_.nodes.size());
  std::vector<int> own, top;
  cut_.clear();
  if (nparts <= 1) {
    nparts = 1; part = 0;
    for (int i = 0; i < N; i++) own.push_back(i);
  } else {
    if (nparts & (nparts - 1)) throw std::invalid_argument("number of parts must be a power of two");
    if (part < 0 || part >= nparts) throw std::invalid_argument("part out of range");
    int depth = 0;
    while ((1 << depth) < nparts) depth++;
    for (int i = 0; i < N; i++) {
      const auto& n = H_.nodes[i];
      if (n.depth < depth) {
        if (n.leaf()) throw std::invalid_argument("HSS tree too shallow for this many parts");
        top.push_back(i);
      } else if (n.depth == depth) cut_.push_back(i);   // pre-order = left to right
    }
    if ((int)cut_.size() != nparts) throw std::logic_error("cut size mismatch");
    // subtree of cut_[part]: contiguous in pre-order
    const int lo = cut_[part];
    const int hi = (part + 1 < nparts) ? cut_[part + 1] : N;
    for (int i = lo; i < hi; i++)
      if (H_.nodes[i].depth >= depth) own.push_back(i);
    // pre-order subtree of `lo` ends where depth returns to <= depth
    own.clear();
    own.push_back(lo);
    for (int i = lo + 1; i < N && H_.nodes[i].depth > depth; i++) own.push_back(i);
  }
  drop_graphs();
  nparts_ = nparts; part_ = part;
  make_lists(own_, own);
  make_lists(top_, top);
  dn_.upload(hn_.data(), hn_.size());   // make_lists set the per-class panel widths
  SB200_CUDA(cudaStreamSynchronize(0));
  factored_ = false;
}

void HSSEngine::owned_range(int* lo, int* hi) const {
  const auto& n = H_.nodes[nparts_ > 1 ? cut_[part_] : 0];
  *lo = n.row_off;
  *hi = n.row_off + n.rows;
}

void HSSEngine::dist_sizes(int s, long long* out) const {
  long long a = 0, f = 0, v = 0;
  for (int c : cut_) {
    const DNode& d = hn_[c];
    a = std::max<long long>(a, (long long)std::max(d.u_rank, d.v_rank) * s);
    f = std::max<long long>(f, (long long)d.u_rank * (d.v_rank + d.u_rank));
    v = std::max<long long>(v, (long long)(d.v_rank + d.u_rank) * s);
  }
  out[0] = std::max<long long>(a, 1); out[1] = std::max<long long>(f, 1); out[2] = std::max<long long>(v, 1);
}

static std::string gkey(const char* op, std::initializer_list<const void*> ptrs, std::initializer_list<long long> ints) {
  std::string k(op);
  char buf[40];
  for (const void* p : ptrs) { std::snprintf(buf, sizeof buf, "|%p", p); k += buf; }
  for (long long v : ints) { std::snprintf(buf, sizeof buf, "|%lld", v); k += buf; }
  return k;
}

// ------------------------------------------------------------------- apply
void HSSEngine::run_up(const NodeLists& L, bool T, int s, const double* dB, int ldB, cudaStream_t st) {
  for (int h = 0; h < L.classes(); h++) {
    const int cnt = L.hptr[h + 1] - L.hptr[h];
    if (!cnt) continue;
    const int* lst = L.list.p + L.hptr[h];
    const int ldp = smem_ld(std::max(L.max_m[h], 1));
    const size_t smem_mm = sizeof(double) * (size_t)ldp * kMS;
    if (s >= mm_min_ && smem_mm <= kMaxSmem) {   // GEMM-shaped: kMS right-hand sides per CTA
      dim3 grid(cnt, (s + kMS - 1) / kMS);
      if (T) { set_smem(hss_up_mm_kernel<true>, smem_mm);
        hss_up_mm_kernel<true><<<grid, kThreads, smem_mm, st>>>(dn_.p, lst, vals_.p, perms_.p, dB, ldB, t1_.p, s, ldp);
      } else { set_smem(hss_up_mm_kernel<false>, smem_mm);
        hss_up_mm_kernel<false><<<grid, kThreads, smem_mm, st>>>(dn_.p, lst, vals_.p, perms_.p, dB, ldB, t1_.p, s, ldp);
      }
      launches_++;
      continue;
    }
    size_t smem = sizeof(double) * (size_t)std::max(L.max_m[h], 1);
    dim3 grid(cnt, s);
    if (T) { set_smem(hss_up_kernel<true>, smem);
      hss_up_kernel<true><<<grid, kThreads, smem, st>>>(dn_.p, lst, vals_.p, perms_.p, dB, ldB, t1_.p, s);
    } else { set_smem(hss_up_kernel<false>, smem);
      hss_up_kernel<false><<<grid, kThreads, smem, st>>>(dn_.p, lst, vals_.p, perms_.p, dB, ldB, t1_.p, s);
    }
    launches_++;
  }
}

void HSSEngine::run_down(const NodeLists& L, bool T, int s, const double* dB, int ldB,
                         double* dC, int ldC, bool leaves, cudaStream_t st, double beta) {
  for (int h = L.classes() - 1; h >= 1; h--) {
    const int cnt = L.hptr[h + 1] - L.hptr[h];
    if (!cnt) continue;
    const int* lst = L.list.p + L.hptr[h];
    const int ldp = smem_ld(std::max(L.max_m[h], 1));
    const size_t smem_mm = sizeof(double) * (size_t)ldp * kMS * 4;
    if (s >= mm_min_ && smem_mm <= kMaxSmem) {
      dim3 grid(cnt, (s + kMS - 1) / kMS);
      if (T) { set_smem(hss_down_mm_kernel<true>, smem_mm);
        hss_down_mm_kernel<true><<<grid, kThreads, smem_mm, st>>>(dn_.p, lst, vals_.p, perms_.p, t1_.p, t2_.p, s, ldp);
      } else { set_smem(hss_down_mm_kernel<false>, smem_mm);
        hss_down_mm_kernel<false><<<grid, kThreads, smem_mm, st>>>(dn_.p, lst, vals_.p, perms_.p, t1_.p, t2_.p, s, ldp);
      }
      launches_++;
      continue;
    }
    size_t smem = sizeof(double) * (size_t)(2 * std::max(L.max_m[h], 1) + 8);
    dim3 grid(cnt, s);
    if (T) { set_smem(hss_down_kernel<true>, smem);
      hss_down_kernel<true><<<grid, kThreads, smem, st>>>(dn_.p, lst, vals_.p, perms_.p, t1_.p, t2_.p, s);
    } else { set_smem(hss_down_kernel<false>, smem);
      hss_down_kernel<false><<<grid, kThreads, smem, st>>>(dn_.p, lst, vals_.p, perms_.p, t1_.p, t2_.p, s);
    }
    launches_++;
  }
  if (leaves && L.classes() > 0 && L.hptr[1] > L.hptr[0]) {
    const int cnt = L.hptr[1] - L.hptr[0];
    const int ldp = smem_ld(std::max(L.max_m[0], 1));
    const size_t smem_mm = sizeof(double) * (size_t)ldp * kMS * 3;
    if (s >= mm_min_ && smem_mm <= kMaxSmem) {
      dim3 grid(cnt, (s + kMS - 1) / kMS);
      if (T) { set_smem(hss_leaf_mm_kernel<true>, smem_mm);
        hss_leaf_mm_kernel<true><<<grid, kThreads, smem_mm, st>>>(dn_.p, L.list.p, vals_.p, perms_.p, dB, ldB, t2_.p, dC, ldC, s, ldp, beta);
      } else { set_smem(hss_leaf_mm_kernel<false>, smem_mm);
        hss_leaf_mm_kernel<false><<<grid, kThreads, smem_mm, st>>>(dn_.p, L.list.p, vals_.p, perms_.p, dB, ldB, t2_.p, dC, ldC, s, ldp, beta);
      }
      launches_++;
      return;
    }
    size_t smem = sizeof(double) * (size_t)(3 * std::max(L.max_m[0], 1) + 8);
    dim3 grid(cnt, s);
    if (T) { set_smem(hss_leaf_kernel<true>, smem);
      hss_leaf_kernel<true><<<grid, kThreads, smem, st>>>(dn_.p, L.list.p, vals_.p, perms_.p, dB, ldB, t2_.p, dC, ldC, s, beta);
    } else { set_smem(hss_leaf_kernel<false>, smem);
      hss_leaf_kernel<false><<<grid, kThreads, smem, st>>>(dn_.p, L.list.p, vals_.p, perms_.p, dB, ldB, t2_.p, dC, ldC, s, beta);
    }
    launches_++;
  }
}

void HSSEngine::mult(char trans, int s, const double* dB, int ldB, double* dC,
                     int ldC, cudaStream_t st, double beta) {
  if (nparts_ > 1) throw std::logic_error("sharded matrix: use the dist_* entry points");
  const bool T = !(trans == 'N' || trans == 'n');
  if (s <= 0) return;
  ensure_apply_ws(s);
  long long bb;
  std::memcpy(&bb, &beta, sizeof bb);
  run_graphed(gkey("mult", {dB, dC}, {T, s, ldB, ldC, bb}), st, [&] {
    run_up(own_, T, s, dB, ldB, st);
    run_down(own_, T, s, dB, ldB, dC, ldC, true, st, beta);
  });
  SB200_CUDA(cudaGetLastError());
}

// sharded apply: local up-sweep, export t1 of this rank's cut node ...
void HSSEngine::dist_mult_begin(char trans, int s, const double* dB, int ldB,
                                double* send, cudaStream_t st) {
  const bool T = !(trans == 'N' || trans == 'n');
  ensure_apply_ws(s);
  run_graphed(gkey("mult_begin", {dB, send}, {T, s, ldB}), st, [&] {
    run_up(own_, T, s, dB, ldB, st);
    const DNode& d = hn_[cut_[part_]];
    const int r = T ? d.u_rank : d.v_rank;
    copy2d(send, r, t1_.p + (size_t)d.w_off * s, r, r, s, st);
  });
  SB200_CUDA(cudaGetLastError());
}
// ... import every cut node's t1, sweep the replicated top, finish locally
void HSSEngine::dist_mult_end(char trans, int s, const double* dB, int ldB, double* dC,
                              int ldC, const double* recv, cudaStream_t st) {
  const bool T = !(trans == 'N' || trans == 'n');
  long long sz[3];
  dist_sizes(s, sz);
  run_graphed(gkey("mult_end", {dB, dC, recv}, {T, s, ldB, ldC}), st, [&] {
    for (int c = 0; c < nparts_; c++) {
      const DNode& d = hn_[cut_[c]];
      const int r = T ? d.u_rank : d.v_rank;
      copy2d(t1_.p + (size_t)d.w_off * s, r, recv + (size_t)c * sz[0], r, r, s, st);
    }
    run_up(top_, T, s, dB, ldB, st);
    run_down(top_, T, s, dB, ldB, dC, ldC, false, st);
    run_down(own_, T, s, dB, ldB, dC, ldC, true, st);
    // the cut node itself is an inner (or leaf) node of own_: its down kernel ran above
  });
  SB200_CUDA(cudaGetLastError());
}

void HSSEngine::shift(double sigma, cudaStream_t st) {
  const int cnt = own_.classes() ? own_.hptr[1] - own_.hptr[0] : 0;
  if (cnt) hss_shift_kernel<<<cnt, 128, 0, st>>>(dn_.p, own_.list.p, vals_.p, sigma);
  launches_++;
  SB200_CUDA(cudaGetLastError());
  factored_ = false;
}

void HSSEngine::sync_host_values() {
  SB200_CUDA(cudaMemcpy(H_.vals.data(), vals_.p, H_.vals.size() * sizeof(double),
                        cudaMemcpyDeviceToHost));
}

void HSSEngine::export_ulv(double* factors, double* tfactors, long long* sizes) {
  if (!factored_) throw std::logic_error("ULV factors requested before factor()");
  const long long nf = fact_len_, nt = (long long)nb_ * tot_k_;
  if (sizes) { sizes[0] = nf; sizes[1] = nt; }
  SB200_CUDA(cudaDeviceSynchronize());
  if (factors) SB200_CUDA(cudaMemcpy(factors, fact_.p, nf * sizeof(double), cudaMemcpyDeviceToHost));
  if (tfactors) SB200_CUDA(cudaMemcpy(tfactors, tfac_.p, nt * sizeof(double), cudaMemcpyDeviceToHost));
}

template <int NB> static size_t qr_smem(int ldv) {
  constexpr int LDW = ((NB + 15) / 16) * 16 + 4;
  return sizeof(double) * ((size_t)ldv * NB + 2 * LDW * NB + LDW * 8 + 3 * NB + 64);
}

// ------------------------------------------------------------------ extract
void HSSEngine::extract(int nI, const int* I, int nJ, const int* J, double* dB, int ldB,
                        bool add, cudaStream_t st) {
  if (nI <= 0 || nJ <= 0) return;
  // leaves in index order (pre-order numbering = left to right)
  if (leaf_ids_.empty()) {
    for (int i = 0; i < (int)H_.nodes.size(); i++)
      if (H_.nodes[i].leaf()) leaf_ids_.push_back(i);
    max_depth_ = 0;
    for (const auto& n : H_.nodes) max_depth_ = std::max(max_depth_, n.depth);
  }
  const int maxd = std::max(max_depth_, 1), maxr = std::max(H_.max_rank(), 1);
  auto build_paths = [&](int cnt, const int* idx, bool cols, std::vector<int>& paths) {
    paths.assign((size_t)cnt * maxd, -1);
    for (int q = 0; q < cnt; q++) {
      const int i = idx[q];
      if (i < 0 || i >= (cols ? H_.cols() : H_.rows())) throw std::invalid_argument("extract: index out of range");
      int lo = 0, hi = (int)leaf_ids_.size() - 1;     // last leaf with offset <= i
      while (lo < hi) {
        const int mid = (lo + hi + 1) / 2;
        const auto& n = H_.nodes[leaf_ids_[mid]];
        if ((cols ? n.col_off : n.row_off) <= i) lo = mid; else hi = mid - 1;
      }
      for (int t = leaf_ids_[lo]; t > 0; t = H_.nodes[t].parent)
        paths[(size_t)q * maxd + H_.nodes[t].depth - 1] = t;
    }
  };
  std::vector<int> pI, pJ;
  build_paths(nI, I, false, pI);
  build_paths(nJ, J, true, pJ);
  DevBuf<int> dI, dJ, dpI, dpJ;
  dI.upload(I, nI, st); dJ.upload(J, nJ, st);
  dpI.upload(pI.data(), pI.size(), st); dpJ.upload(pJ.data(), pJ.size(), st);
  DevBuf<double> cu((size_t)nI * maxd * maxr), cv((size_t)nJ * maxd * maxr);
  const size_t smem = sizeof(double) * 2 * (size_t)maxr + sizeof(int) * (size_t)(maxr + 8);
  set_smem(hss_basis_rows_kernel<false>, smem);
  set_smem(hss_basis_rows_kernel<true>, smem);
  hss_basis_rows_kernel<false><<<nI, 128, smem, st>>>(dn_.p, vals_.p, perms_.p, dI.p, dpI.p, maxd, maxr, cu.p);
  hss_basis_rows_kernel<true><<<nJ, 128, smem, st>>>(dn_.p, vals_.p, perms_.p, dJ.p, dpJ.p, maxd, maxr, cv.p);
  const long long tot = (long long)nI * nJ;
  hss_extract_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(dn_.p, vals_.p, dI.p, dJ.p, nI, nJ, dpI.p, dpJ.p,
                                                                    maxd, maxr, cu.p, cv.p, dB, ldB, add ? 1 : 0);
  launches_ += 3;
  SB200_CUDA(cudaGetLastError());
  SB200_CUDA(cudaStreamSynchronize(st));   // the index / path buffers are locals
}

// ------------------------------------------------------------------ factor
void HSSEngine::factor_prepare(bool whole) {
  if (whole) {
    if (H_.rows() != H_.cols())
      throw std::invalid_argument("ULV factorization needs a square matrix");
    for (auto& n : H_.nodes)
      if (n.leaf() && n.rows != n.cols)
        throw std::invalid_argument("ULV factorization needs square diagonal blocks");
  }
  // zero-filled on allocation: the alignment padding between blocks is never written
  const size_t nf = (size_t)std::max<long long>(fact_len_, 1), nt = (size_t)std::max<long long>((long long)nb_ * tot_k_, 1);
  if (fact_.n < nf || tfac_.n < nt) drop_graphs();
  if (fact_.n < nf) { fact_.alloc(nf); SB200_CUDA(cudaMemset(fact_.p, 0, nf * sizeof(double))); tmaps_for_ = nullptr; }
  if (qr3_ && tmaps_for_ != fact_.p) {   // TMA descriptors of the factor blocks (they carry the arena's address)
    build_qr3_tmaps(hn_, fact_.p, tmaps_);
    tmaps_for_ = fact_.p;
  }
  if (tfac_.n < nt) { tfac_.alloc(nt); SB200_CUDA(cudaMemset(tfac_.p, 0, nt * sizeof(double))); }
  const void* old_piv = rootpiv_.p;
  const void* old_scr = scratch_.p;
  rootpiv_.ensure(std::max(hn_[0].m, 1));
  scratch_.ensure((size_t)std::max<long long>(std::max(std::max(own_.smax, top_.smax), sub0_.smax), 1));
  if (old_piv != rootpiv_.p || old_scr != scratch_.p) drop_graphs();
}

void HSSEngine::factor_classes(const NodeLists& L, bool time_leaf, cudaStream_t st, int lu_node,
                               double* lu_dst, int* lu_piv) {
  for (int h = 0; h < L.classes(); h++) {
    const int cnt = L.hptr[h + 1] - L.hptr[h];
    if (!cnt) continue;
    const int* lst = L.list.p + L.hptr[h];
    const long long* so = L.dsoff.p + L.hptr[h];
    const int mm = std::max(L.max_m[h], 1);
    if (h == 0) {
      ulv_vh_leaf_kernel<<<cnt, kThreads, 0, st>>>(dn_.p, lst, vals_.p, perms_.p, fact_.p);
      launches_++;
    } else {
      // inverse permutation of V + (when it fits) the two children's Vt1 blocks
      const size_t need = sizeof(double) * ((size_t)(mm + 1) / 2 + 2 + (size_t)mm * mm);
      const int stage = need <= 160 * 1024;
      size_t smem = stage ? need : sizeof(int) * (size_t)(mm + 8);
      set_smem(ulv_build_inner_kernel, smem);
      // few nodes: several CTAs per node (the kernel slices its entry loops over gridDim.y)
      const int ny = cnt >= 2 * nsm_ ? 1 : std::min(8, std::max(1, 2 * nsm_ / cnt));
      ulv_build_inner_kernel<<<dim3(cnt, ny), kThreads, smem, st>>>(dn_.p, lst, vals_.p, perms_.p, fact_.p, scratch_.p, so, stage);
      launches_++;
    }
    if (cnt == 1 && L.host[L.hptr[h]] == lu_node) {   // the root (of the factored subtree): LU
      const DNode& root = hn_[lu_node];
      const long long n2 = (long long)root.m * root.m;
      const double* src = root.leaf ? vals_.p + root.D : scratch_.p;  // slab offset 0 of its class
      double* dst = lu_dst ? lu_dst : fact_.p + root.F;
      if (n2 > 0) copy_block_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(src, dst, n2);
      {
        const size_t bytes = sizeof(double) * ((size_t)root.m * root.m + root.m);   // + reciprocal pivots
        const int use_smem = bytes <= 200 * 1024;
        set_smem(ulv_root_lu_kernel, use_smem ? bytes : 0);
        ulv_root_lu_kernel<<<1, kThreads, use_smem ? bytes : 0, st>>>(dst, root.m, lu_piv ? lu_piv : rootpiv_.p, use_smem);
      }
      launches_ += 2;
      continue;
    }
    // elimination
    if (mm <= 640 && elim_variant_ == 1) {          // 16-column tiles, 4 CTAs per SM
      size_t smem = sizeof(double) * (size_t)mm * 17 + sizeof(int) * (size_t)(mm + 8);
      set_smem(ulv_eliminate_kernel<16, 4, false>, smem);
      ulv_eliminate_kernel<16, 4, false><<<cnt, kThreads, smem, st>>>(dn_.p, lst, vals_.p, perms_.p, fact_.p, scratch_.p, so);
    } else if (mm <= 640 && elim_variant_ == 2) {   // 16-column tiles, prefetched one ahead (cp.async)
      size_t smem = sizeof(double) * (size_t)mm * 17 * 2 + sizeof(int) * (size_t)(mm + 8);
      set_smem(ulv_eliminate_kernel<16, 3, true>, smem);
      ulv_eliminate_kernel<16, 3, true><<<cnt, kThreads, smem, st>>>(dn_.p, lst, vals_.p, perms_.p, fact_.p, scratch_.p, so);
    } else if (mm <= 640) {
      size_t smem = sizeof(double) * (size_t)mm * 33 + sizeof(int) * (size_t)(mm + 8);
      set_smem(ulv_eliminate_kernel<32>, smem);
      ulv_eliminate_kernel<32><<<cnt, kThreads, smem, st>>>(dn_.p, lst, vals_.p, perms_.p, fact_.p, scratch_.p, so);
    } else {
      size_t smem = sizeof(double) * (size_t)mm * 9 + sizeof(int) * (size_t)(mm + 8);
      set_smem(ulv_eliminate_kernel<8>, smem);
      ulv_eliminate_kernel<8><<<cnt, kThreads, smem, st>>>(dn_.p, lst, vals_.p, perms_.p, fact_.p, scratch_.p, so);
    }
    launches_++;
    const int ldv = smem_ld(mm);
    const int gmm = std::max(class_gmm_[h], mm);   // kernel variant by the tree-wide class maximum (see make_lists)
    const bool timed = profile_ && time_leaf && h == 0;
    if (timed) SB200_CUDA(cudaEventRecord(ev_[0], st));
    {
      // one launch per class: the whole blocked QR of every node of the class
      const int nowide = qr_nowide_ ? 1 : 0;
      if (nb_ == 32 && gmm <= 256 && qr3_) {          // optional left-looking TMA-fed kernel
        set_smem(qr3::ulv_qr3_kernel, qr3::SMEM_BYTES);
        qr3::ulv_qr3_kernel<<<cnt, qr3::NTHREADS, qr3::SMEM_BYTES, st>>>(dn_.p, lst, fact_.p, tfac_.p, tmaps_.p);
      } else if (nb_ == 32 && gmm <= 256 && qr_regpanel_ && qr_variant_ == 1) {   // default for m <= 256
        size_t smem = qr_smem<16>(ldv); set_smem(ulv_qr_kernel<16, true, 128>, smem);
        ulv_qr_kernel<16, true, 128><<<cnt, 128, smem, st>>>(dn_.p, lst, fact_.p, tfac_.p, ldv, nowide);
      } else if (nb_ == 32 && gmm <= 256 && qr_regpanel_ && qr_variant_ == 2) {
        size_t smem = qr_smem<16>(ldv); set_smem(ulv_qr_kernel<16, true, 256>, smem);
        ulv_qr_kernel<16, true, 256><<<cnt, kThreads, smem, st>>>(dn_.p, lst, fact_.p, tfac_.p, ldv, nowide);
      } else if (nb_ == 32 && gmm <= 256 && qr_regpanel_) {
        size_t smem = qr_smem<32>(ldv); set_smem(ulv_qr_kernel<32, true, 256>, smem);
        ulv_qr_kernel<32, true, 256><<<cnt, kThreads, smem, st>>>(dn_.p, lst, fact_.p, tfac_.p, ldv, nowide);
      } else if (nb_ == 32) {
        size_t smem = qr_smem<32>(ldv); set_smem(ulv_qr_kernel<32, false, 256>, smem);
        ulv_qr_kernel<32, false, 256><<<cnt, kThreads, smem, st>>>(dn_.p, lst, fact_.p, tfac_.p, ldv, nowide);
      } else {
        size_t smem = qr_smem<16>(ldv); set_smem(ulv_qr_kernel<16, false, 256>, smem);
        ulv_qr_kernel<16, false, 256><<<cnt, kThreads, smem, st>>>(dn_.p, lst, fact_.p, tfac_.p, ldv, nowide);
      }
      launches_++;
    }
    if (timed) SB200_CUDA(cudaEventRecord(ev_[1], st));
  }
}

void HSSEngine::factor(cudaStream_t st) {
  if (nparts_ > 1) throw std::logic_error("sharded matrix: use the dist_* entry points");
  factor_prepare();
  run_graphed("factor", st, [&] { factor_classes(own_, true, st); });
  SB200_CUDA(cudaGetLastError());
  factored_ = true;
  pf_ok_ = false;   // the factor arena now holds the full factorization
}

void HSSEngine::dist_factor_begin(double* send, cudaStream_t st) {
  factor_prepare();
  run_graphed(gkey("factor_begin", {send}, {}), st, [&] {
    factor_classes(own_, true, st);
    // export [Vt1 | Dt^T] of this rank's cut node: rows k..m, columns k..naug of F
    const DNode& d = hn_[cut_[part_]];
    copy2d(send, d.u_rank, fact_.p + d.F + d.k + (size_t)d.k * d.m, d.m, d.u_rank, d.v_rank + d.u_rank, st);
  });
  SB200_CUDA(cudaGetLastError());
}

void HSSEngine::dist_factor_end(const double* recv, cudaStream_t st) {
  long long sz[3];
  dist_sizes(1, sz);
  run_graphed(gkey("factor_end", {recv}, {}), st, [&] {
    for (int c = 0; c < nparts_; c++) {
      const DNode& d = hn_[cut_[c]];
      copy2d(fact_.p + d.F + d.k + (size_t)d.k * d.m, d.m, recv + (size_t)c * sz[1], d.u_rank,
             d.u_rank, d.v_rank + d.u_rank, st);
    }
    factor_classes(top_, false, st);
  });
  SB200_CUDA(cudaGetLastError());
  factored_ = true;
}


// ---------------------------------------------------------------------------
// Sharded operations with the exchange inside the engine (NCCL bound at run time)
// ---------------------------------------------------------------------------
namespace {
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
const NcclApi& nccl_api() {
  static NcclApi api;
  static bool done = false;
  if (!done) {
    // the copy a host framework (torch) has already loaded is found by its soname
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) throw std::runtime_error(std::string("NCCL is not available: ") + dlerror());
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(h, "ncclAllGather"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    if (!api.GetUniqueId || !api.CommInitRank || !api.AllGather || !api.CommDestroy)
      throw std::runtime_error("NCCL symbols not found");
    done = true;
  }
  return api;
}
void nccl_check(ncclResult_t r, const char* what) {
  if (r != ncclSuccess) {
    const auto& a = nccl_api();
    throw std::runtime_error(std::string("NCCL error in ") + what + ": " + (a.GetErrorString ? a.GetErrorString(r) : "?"));
  }
}
}  // namespace

void HSSEngine::nccl_unique_id(char* out128) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclUniqueId id;
  nccl_check(nccl_api().GetUniqueId(&id), "ncclGetUniqueId");
  std::memcpy(out128, &id, sizeof id);
}

void HSSEngine::dist_init(int nparts, int part, const char* unique_id128) {
  set_partition(nparts, part);
  dist_close();
  if (nparts <= 1) return;
  ncclUniqueId id;
  std::memcpy(&id, unique_id128, sizeof id);
  ncclComm_t comm = nullptr;
  nccl_check(nccl_api().CommInitRank(&comm, nparts, id, part), "ncclCommInitRank");
  nccl_comm_ = comm;
}

void HSSEngine::dist_close() {
  if (!nccl_comm_) return;
  try { nccl_api().CommDestroy(static_cast<ncclComm_t>(nccl_comm_)); } catch (...) { }
  nccl_comm_ = nullptr;
}

void HSSEngine::all_gather(int which, long long count, cudaStream_t st) {
  nccl_check(nccl_api().AllGather(xsend_[which].p, xrecv_[which].p, (size_t)count, ncclDouble,
                                  static_cast<ncclComm_t>(nccl_comm_), st), "ncclAllGather");
}

void HSSEngine::dist_mult(char trans, int s, const double* dB, int ldB, double* dC, int ldC, cudaStream_t st) {
  if (!nccl_comm_) throw std::logic_error("dist_mult: call dist_init first");
  const bool T = !(trans == 'N' || trans == 'n');
  if (s <= 0) return;
  ensure_apply_ws(s);
  long long sz[3];
  dist_sizes(s, sz);
  if (xsend_[0].n < (size_t)sz[0]) { drop_graphs(); xsend_[0].alloc(sz[0]); xrecv_[0].alloc(sz[0] * nparts_); }
  run_graphed(gkey("dist_mult", {dB, dC}, {T, s, ldB, ldC}), st, [&] {
    run_up(own_, T, s, dB, ldB, st);
    {
      const DNode& d = hn_[cut_[part_]];
      const int r = T ? d.u_rank : d.v_rank;
      copy2d(xsend_[0].p, r, t1_.p + (size_t)d.w_off * s, r, r, s, st);
    }
    all_gather(0, sz[0], st);
    for (int c = 0; c < nparts_; c++) {
      const DNode& d = hn_[cut_[c]];
      const int r = T ? d.u_rank : d.v_rank;
      copy2d(t1_.p + (size_t)d.w_off * s, r, xrecv_[0].p + (size_t)c * sz[0], r, r, s, st);
    }
    run_up(top_, T, s, dB, ldB, st);
    run_down(top_, T, s, dB, ldB, dC, ldC, false, st);
    run_down(own_, T, s, dB, ldB, dC, ldC, true, st);
  });
  SB200_CUDA(cudaGetLastError());
}

void HSSEngine::dist_factor(cudaStream_t st) {
  if (!nccl_comm_) throw std::logic_error("dist_factor: call dist_init first");
  factor_prepare();
  long long sz[3];
  dist_sizes(1, sz);
  if (xsend_[1].n < (size_t)sz[1]) { drop_graphs(); xsend_[1].alloc(sz[1]); xrecv_[1].alloc(sz[1] * nparts_); }
  run_graphed("dist_factor", st, [&] {
    factor_classes(own_, true, st);
    {
      const DNode& d = hn_[cut_[part_]];
      copy2d(xsend_[1].p, d.u_rank, fact_.p + d.F + d.k + (size_t)d.k * d.m, d.m, d.u_rank, d.v_rank + d.u_rank, st);
    }
    all_gather(1, sz[1], st);
    for (int c = 0; c < nparts_; c++) {
      const DNode& d = hn_[cut_[c]];
      copy2d(fact_.p + d.F + d.k + (size_t)d.k * d.m, d.m, xrecv_[1].p + (size_t)c * sz[1], d.u_rank,
             d.u_rank, d.v_rank + d.u_rank, st);
    }
    factor_classes(top_, false, st);
  });
  SB200_CUDA(cudaGetLastError());
  factored_ = true;
}

void HSSEngine::dist_solve(int s, double* dB, int ldB, cudaStream_t st) {
  if (!nccl_comm_) throw std::logic_error("dist_solve: call dist_init first");
  if (!factored_) throw std::logic_error("solve called before factor");
  if (s <= 0) return;
  ensure_solve_ws(s);
  long long sz[3];
  dist_sizes(s, sz);
  if (xsend_[2].n < (size_t)sz[2]) { drop_graphs(); xsend_[2].alloc(sz[2]); xrecv_[2].alloc(sz[2] * nparts_); }
  run_graphed(gkey("dist_solve", {dB}, {s, ldB}), st, [&] {
    solve_fwd(own_, s, dB, ldB, st);
    {
      const DNode& d = hn_[cut_[part_]];
      const int ld = d.v_rank + d.u_rank;
      copy2d(xsend_[2].p, ld, zsol_.p + (size_t)d.z_off * s, d.v_rank, d.v_rank, s, st);
      copy2d(xsend_[2].p + d.v_rank, ld, fsol_.p + (size_t)d.f_off * s, d.u_rank, d.u_rank, s, st);
    }
    all_gather(2, sz[2], st);
    for (int c = 0; c < nparts_; c++) {
      const DNode& d = hn_[cut_[c]];
      const int ld = d.v_rank + d.u_rank;
      const double* src = xrecv_[2].p + (size_t)c * sz[2];
      copy2d(zsol_.p + (size_t)d.z_off * s, d.v_rank, src, ld, d.v_rank, s, st);
      copy2d(fsol_.p + (size_t)d.f_off * s, d.u_rank, src + d.v_rank, ld, d.u_rank, s, st);
    }
    solve_fwd(top_, s, dB, ldB, st);
    solve_root(s, dB, ldB, st);
    solve_bwd(top_, s, dB, ldB, st);
    solve_bwd(own_, s, dB, ldB, st);
  });
  SB200_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------
// Schur complement of the (0,0) block (HSS fronts)
// ---------------------------------------------------------------------------
static void dgemm(bool ta, bool tb, int M, int N, int K, double alpha, const double* A, int lda,
                  const double* B, int ldb, double beta, double* C, int ldc, cudaStream_t st) {
  if (M <= 0 || N <= 0) return;
  dim3 grid((M + 127) / 128, (N + 63) / 64);
  if (!ta && !tb) dense_gemm_kernel<false, false><<<grid, kThreads, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  else if (ta && !tb) dense_gemm_kernel<true, false><<<grid, kThreads, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  else if (!ta && tb) dense_gemm_kernel<false, true><<<grid, kThreads, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  else dense_gemm_kernel<true, true><<<grid, kThreads, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
}

void HSSEngine::schur_lists() {
  if (nparts_ > 1) throw std::logic_error("sharded matrix: Schur operations need the whole tree on one GPU");
  if (sub_ok_) return;
  const auto& root = H_.nodes[0];
  if (root.leaf()) throw std::logic_error("Schur complement needs a root with two children");
  const int N = int(H_.nodes.size());
  std::vector<int> s0, s1;   // pre-order: subtree(ch0) = [ch0, ch1), subtree(ch1) = [ch1, N)
  for (int i = root.ch0; i < root.ch1; i++) s0.push_back(i);
  for (int i = root.ch1; i < N; i++) s1.push_back(i);
  make_lists(sub0_, s0);
  make_lists(sub1_, s1);
  dn_.upload(hn_.data(), hn_.size());
  const size_t nz = (size_t)std::max(std::max(H_.rows(), H_.cols()), 1);
  zeros_.alloc(nz);
  SB200_CUDA(cudaMemset(zeros_.p, 0, nz * sizeof(double)));
  SB200_CUDA(cudaStreamSynchronize(0));
  sub_ok_ = true;
}

void HSSEngine::schur_sizes(int* out) const {
  const DNode& r = hn_[0];
  if (r.leaf) throw std::logic_error("Schur complement needs a root with two children");
  const DNode &c0 = hn_[r.ch0], &c1 = hn_[r.ch1];
  out[0] = c1.rows; out[1] = c1.cols; out[2] = c0.v_rank; out[3] = c0.m;
  out[4] = c1.v_rank; out[5] = c1.u_rank; out[6] = c0.rows;
}

// partial_factor: ULV-factor the subtree of child 0 with child 0 as its root (LU
// of its reduced block D0) and keep Vhat = Vh(child 0)      (factor.hpp:44-50)
void HSSEngine::partial_factor(cudaStream_t st) {
  schur_lists();
  const int ch0 = hn_[0].ch0;
  const DNode& c0 = hn_[ch0];
  if (c0.rows != c0.cols) throw std::invalid_argument("partial_factor needs a square (0,0) block");
  for (int i : sub0_.host)
    if (H_.nodes[i].leaf() && H_.nodes[i].rows != H_.nodes[i].cols)
      throw std::invalid_argument("ULV factorization needs square diagonal blocks");
  factor_prepare(false);
  const int m0 = c0.m, rv0 = c0.v_rank;
  pf_lu_.ensure((size_t)std::max(m0 * m0, 1));
  pf_piv_.ensure((size_t)std::max(m0, 1));
  vhat_.ensure((size_t)std::max(m0 * rv0, 1));
  factor_classes(sub0_, false, st, ch0, pf_lu_.p, pf_piv_.p);
  copy2d(vhat_.p, m0, fact_.p + c0.F + (size_t)c0.k * c0.m, c0.m, m0, rv0, st);
  launches_++;
  SB200_CUDA(cudaGetLastError());
  factored_ = false;
  pf_ok_ = true;
  pfwd_s_ = 0;
}

// Schur_update: Theta = U1big B10, DUB01 = D0^{-1} U0 B01, Phi = V1big DUB01^H
//                                                   (HSSMatrix.Schur.hpp:40-59)
void HSSEngine::schur_update(double* dTheta, int ldT, double* dDUB01, int ldD, double* dPhi,
                             int ldP, cudaStream_t st) {
  if (!pf_ok_) throw std::logic_error("Schur_update called before partial_factor");
  const DNode& r = hn_[0];
  const int ch0 = r.ch0;
  const DNode &c0 = hn_[ch0], &c1 = hn_[r.ch1];
  const int m0 = c0.m, ru0 = c0.u_rank, rv0 = c0.v_rank, ru1 = c1.u_rank, rv1 = c1.v_rank;
  if (m0 > 0 && rv1 > 0) {
    if (ru0 > 0) {
      const long long tot = (long long)m0 * rv1;
      basis_apply_kernel<<<(unsigned)std::min<long long>((tot + kThreads - 1) / kThreads, 1024), kThreads, 0, st>>>(
          dn_.p, ch0, vals_.p, perms_.p, vals_.p + r.B01, ru0, rv1, dDUB01, ldD);
    } else {
      SB200_CUDA(cudaMemset2DAsync(dDUB01, sizeof(double) * ldD, 0, sizeof(double) * m0, rv1, st));
    }
    lu_solve_kernel<<<rv1, kThreads, sizeof(double) * (size_t)m0, st>>>(pf_lu_.p, m0, pf_piv_.p, dDUB01, ldD);
    launches_ += 2;
  }
  const int smax = std::max(std::max(rv0, m0), 1);
  ensure_apply_ws(smax);
  SB200_CUDA(cudaMemsetAsync(t1_.p, 0, t1_.n * sizeof(double), st));
  if (rv0 > 0 && c1.rows > 0) {      // Theta: down-sweep of child 1 started from t2(ch1) = B10
    const int s = rv0;
    if (ru1 > 0)
      SB200_CUDA(cudaMemcpyAsync(t2_.p + (size_t)c1.w_off * s, vals_.p + r.B10,
                                 sizeof(double) * (size_t)ru1 * rv0, cudaMemcpyDeviceToDevice, st));
    run_down(sub1_, false, s, zeros_.p, 0, dTheta - c1.row_off, ldT, true, st);
  }
  if (m0 > 0 && c1.cols > 0) {       // Phi: transposed down-sweep from t2(ch1) = DUB01^H
    const int s = m0;
    if (rv1 > 0) {
      const long long tot = (long long)m0 * rv1;
      transpose_kernel<<<(unsigned)std::min<long long>((tot + 255) / 256, 1024), 256, 0, st>>>(
          dDUB01, ldD, m0, rv1, t2_.p + (size_t)c1.w_off * s, rv1);
      launches_++;
    }
    run_down(sub1_, true, s, zeros_.p, 0, dPhi - c1.col_off, ldP, true, st);
  }
  SB200_CUDA(cudaGetLastError());
}

// Schur_product_direct                              (HSSMatrix.Schur.hpp:73-137)
//   Sr = H11 R   - Theta Vhat^H DUB01 (V1big^H R)
//   Sc = H11^H R - Phi Vhat B10^H (U1big^H R)
// V1big^H R / U1big^H R are the up-sweeps of the two products with H11, reused.
void HSSEngine::schur_product_direct(const double* dTheta, int ldT, const double* dDUB01, int ldD,
                                     const double* dPhi, int ldP, int c, const double* dR, int ldR,
                                     double* dSr, int ldSr, double* dSc, int ldSc,
                                     cudaStream_t st) {
  if (!pf_ok_) throw std::logic_error("Schur_product_direct called before partial_factor");
  if (c <= 0) return;
  const DNode& r = hn_[0];
  const DNode &c0 = hn_[r.ch0], &c1 = hn_[r.ch1];
  if (c1.rows != c1.cols) throw std::invalid_argument("Schur_product_direct needs a square (1,1) block");
  const int m0 = c0.m, rv0 = c0.v_rank, ru1 = c1.u_rank, rv1 = c1.v_rank;
  ensure_apply_ws(c);
  stmp_[0].ensure((size_t)std::max(m0, 1) * c);
  stmp_[1].ensure((size_t)std::max(rv0, 1) * c);
  double* tA = stmp_[0].p;   // m0 x c
  double* tB = stmp_[1].p;   // rv0 x c
  double* t1c1 = t1_.p + (size_t)c1.w_off * c;
  double* t2c1 = t2_.p + (size_t)c1.w_off * c;
  // ---- Sr
  run_up(sub1_, false, c, dR - c1.col_off, ldR, st);
  SB200_CUDA(cudaMemsetAsync(t2c1, 0, sizeof(double) * (size_t)std::max(ru1, 1) * c, st));
  run_down(sub1_, false, c, dR - c1.col_off, ldR, dSr - c1.row_off, ldSr, true, st);
  dgemm(false, false, m0, c, rv1, 1., dDUB01, ldD, t1c1, std::max(rv1, 1), 0., tA, std::max(m0, 1), st);
  dgemm(true, false, rv0, c, m0, 1., vhat_.p, std::max(m0, 1), tA, std::max(m0, 1), 0., tB, std::max(rv0, 1), st);
  dgemm(false, false, c1.rows, c, rv0, -1., dTheta, ldT, tB, std::max(rv0, 1), 1., dSr, ldSr, st);
  // ---- Sc
  run_up(sub1_, true, c, dR - c1.row_off, ldR, st);
  SB200_CUDA(cudaMemsetAsync(t2c1, 0, sizeof(double) * (size_t)std::max(rv1, 1) * c, st));
  run_down(sub1_, true, c, dR - c1.row_off, ldR, dSc - c1.col_off, ldSc, true, st);
  dgemm(true, false, rv0, c, ru1, 1., vals_.p + r.B10, std::max(ru1, 1), t1c1, std::max(ru1, 1), 0., tB, std::max(rv0, 1), st);
  dgemm(false, false, m0, c, rv0, 1., vhat_.p, std::max(m0, 1), tB, std::max(rv0, 1), 0., tA, std::max(m0, 1), st);
  dgemm(false, false, c1.cols, c, m0, -1., dPhi, ldP, tA, std::max(m0, 1), 1., dSc, ldSc, st);
  launches_ += 6;
  SB200_CUDA(cudaGetLastError());
}

// Schur_product_indirect                            (HSSMatrix.Schur.hpp:139-215)
// with W = B10 Vhat^H DUB01 (ru1 x rv1):
//   Sr = Sr1 - U1big (B10 (V0big^H R0) + W (V1big^H R1))
//   Sc = Sc1 - V1big (B01^H (U0big^H R0) + W^H (U1big^H R1))
// (Sr1, Sc1: rows of H [R0; R1] / H^H [R0; R1] that belong to block 1)
void HSSEngine::schur_product_indirect(const double* dDUB01, int ldD, int c, const double* dR0,
                                       int ldR0, const double* dR1, int ldR1, const double* dSr1,
                                       int ldSr1, const double* dSc1, int ldSc1, double* dSr,
                                       int ldSr, double* dSc, int ldSc, cudaStream_t st) {
  if (!pf_ok_) throw std::logic_error("Schur_product_indirect called before partial_factor");
  if (c <= 0) return;
  const DNode& r = hn_[0];
  const DNode &c0 = hn_[r.ch0], &c1 = hn_[r.ch1];
  const int m0 = c0.m, ru0 = c0.u_rank, rv0 = c0.v_rank, ru1 = c1.u_rank, rv1 = c1.v_rank;
  ensure_apply_ws(c);
  auto pad = [](long long n) { return (size_t)std::max<long long>(n, 1); };
  const size_t nV0 = pad((long long)rv0 * c), nU0 = pad((long long)ru0 * c), nV1 = pad((long long)rv1 * c),
               nU1 = pad((long long)ru1 * c), nVtD = pad((long long)rv0 * rv1), nW = pad((long long)ru1 * rv1);
  stmp_[2].ensure(nV0 + nU0 + nV1 + nU1 + nVtD + nW);
  double* V0tR0 = stmp_[2].p;
  double* U0tR0 = V0tR0 + nV0;
  double* V1tR1 = U0tR0 + nU0;
  double* U1tR1 = V1tR1 + nV1;
  double* VtD = U1tR1 + nU1;
  double* W = VtD + nVtD;
  double* t1c0 = t1_.p + (size_t)c0.w_off * c;
  double* t1c1 = t1_.p + (size_t)c1.w_off * c;
  double* t2c1 = t2_.p + (size_t)c1.w_off * c;
  // apply_UtVt_big on both subtrees = the four up-sweeps
  run_up(sub0_, false, c, dR0 - c0.col_off, ldR0, st);
  copy2d(V0tR0, std::max(rv0, 1), t1c0, std::max(rv0, 1), rv0, c, st);
  run_up(sub0_, true, c, dR0 - c0.row_off, ldR0, st);
  copy2d(U0tR0, std::max(ru0, 1), t1c0, std::max(ru0, 1), ru0, c, st);
  run_up(sub1_, false, c, dR1 - c1.col_off, ldR1, st);
  copy2d(V1tR1, std::max(rv1, 1), t1c1, std::max(rv1, 1), rv1, c, st);
  run_up(sub1_, true, c, dR1 - c1.row_off, ldR1, st);
  copy2d(U1tR1, std::max(ru1, 1), t1c1, std::max(ru1, 1), ru1, c, st);
  // W = B10 (Vhat^H DUB01)
  dgemm(true, false, rv0, rv1, m0, 1., vhat_.p, std::max(m0, 1), dDUB01, ldD, 0., VtD, std::max(rv0, 1), st);
  dgemm(false, false, ru1, rv1, rv0, 1., vals_.p + r.B10, std::max(ru1, 1), VtD, std::max(rv0, 1), 0., W, std::max(ru1, 1), st);
  SB200_CUDA(cudaMemsetAsync(t1_.p, 0, t1_.n * sizeof(double), st));
  // ---- Sr = Sr1 + U1big t2,  t2(ch1) = -(B10 V0tR0 + W V1tR1)
  dgemm(false, false, ru1, c, rv0, -1., vals_.p + r.B10, std::max(ru1, 1), V0tR0, std::max(rv0, 1), 0., t2c1, std::max(ru1, 1), st);
  dgemm(false, false, ru1, c, rv1, -1., W, std::max(ru1, 1), V1tR1, std::max(rv1, 1), 1., t2c1, std::max(ru1, 1), st);
  if (dSr != dSr1) copy2d(dSr, ldSr, dSr1, ldSr1, c1.rows, c, st);
  run_down(sub1_, false, c, zeros_.p, 0, dSr - c1.row_off, ldSr, true, st, 1.);
  // ---- Sc = Sc1 + V1big t2,  t2(ch1) = -(B01^H U0tR0 + W^H U1tR1)
  dgemm(true, false, rv1, c, ru0, -1., vals_.p + r.B01, std::max(ru0, 1), U0tR0, std::max(ru0, 1), 0., t2c1, std::max(rv1, 1), st);
  dgemm(true, false, rv1, c, ru1, -1., W, std::max(ru1, 1), U1tR1, std::max(ru1, 1), 1., t2c1, std::max(rv1, 1), st);
  if (dSc != dSc1) copy2d(dSc, ldSc, dSc1, ldSc1, c1.cols, c, st);
  run_down(sub1_, true, c, zeros_.p, 0, dSc - c1.col_off, ldSc, true, st, 1.);
  launches_ += 12;
  SB200_CUDA(cudaGetLastError());
}

// child(0)->forward_solve(w, b, partial = true)      (solve.hpp:52-60,133-152)
void HSSEngine::partial_forward_solve(int s, double* dB0, int ldB, double* dRed, int ldRed,
                                      cudaStream_t st) {
  if (!pf_ok_) throw std::logic_error("partial forward_solve called before partial_factor");
  if (s <= 0) return;
  ensure_solve_ws(s);
  const int ch0 = hn_[0].ch0;
  const DNode& c0 = hn_[ch0];
  const int h0 = H_.nodes[ch0].height;
  solve_fwd(sub0_, s, dB0, ldB, st, h0);       // the classes below child 0
  solve_root(s, dB0, ldB, st, ch0, pf_lu_.p, pf_piv_.p);
  if (c0.v_rank > 0) {
    const size_t smem = sizeof(double) * (size_t)std::max(c0.v_rows, 1);
    set_smem(partial_reduced_rhs_kernel, smem);
    partial_reduced_rhs_kernel<<<s, kThreads, smem, st>>>(dn_.p, ch0, vals_.p, perms_.p, vhat_.p, xsol_.p,
                                                          zsol_.p, dRed, ldRed, s);
    launches_++;
  }
  pfwd_s_ = s;
  SB200_CUDA(cudaGetLastError());
}

double* HSSEngine::partial_x(int s) {
  if (pfwd_s_ != s || s <= 0) throw std::logic_error("no partial forward_solve with this number of right-hand sides");
  return xsol_.p + (size_t)hn_[hn_[0].ch0].x_off * s;
}

// child(0)->backward_solve(w, x)                    (solve.hpp:62-66,199-238)
void HSSEngine::partial_backward_solve(int s, double* dX0, int ldX, cudaStream_t st) {
  double* x = partial_x(s);
  const int ch0 = hn_[0].ch0;
  const DNode& c0 = hn_[ch0];
  if (c0.leaf) copy2d(dX0, ldX, x, c0.m, c0.m, s, st);
  else solve_bwd(sub0_, s, dX0, ldX, st, H_.nodes[ch0].height);
  launches_++;
  SB200_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------- solve
void HSSEngine::solve_fwd(const NodeLists& L, int s, double* dB, int ldB, cudaStream_t st, int nclass) {
  const int nh = nclass < 0 ? L.classes() : std::min(nclass, L.classes());
  for (int h = 0; h < nh; h++) {
    const int cnt = L.hptr[h + 1] - L.hptr[h];
    if (!cnt) continue;
    const int mm = std::max(L.max_m[h], 1);
    dim3 grid(cnt, s);
    if (solve_pipe_ & 2) {      // stream the factor block through shared memory (cp.async, double buffered)
      const int ldm = (mm + 1) & ~1;
      const int ldbuf = mm | 1;   // odd: the column-strided reads of the diagonal block are conflict free
      const size_t psm = sizeof(double) * ((size_t)3 * ldm + (size_t)2 * ldbuf * kFC);
      if (psm <= kMaxSmem) {
        set_smem(ulv_fwd_pipe_kernel, psm);
        ulv_fwd_pipe_kernel<<<grid, kThreads, psm, st>>>(dn_.p, L.list.p + L.hptr[h], vals_.p, perms_.p, fact_.p, dB, ldB,
                                                         ysol_.p, zsol_.p, fsol_.p, s, ldm, ldbuf);
        launches_++;
        continue;
      }
    }
    size_t smem = sizeof(double) * (size_t)(3 * mm + 32 * 33 + 8);
    set_smem(ulv_fwd_kernel<32>, smem);
    ulv_fwd_kernel<32><<<grid, kThreads, smem, st>>>(dn_.p, L.list.p + L.hptr[h], vals_.p, perms_.p, fact_.p, dB, ldB, ysol_.p, zsol_.p, fsol_.p, s, (solve_pipe_ & 8) ? 1 : 0);
    launches_++;
  }
}

void HSSEngine::solve_root(int s, double* dB, int ldB, cudaStream_t st, int node,
                           const double* lu, const int* piv) {
  const size_t nn = (size_t)std::max(hn_[node].m, 1);
  const int use_smem = sizeof(double) * (nn * nn + nn + 2) <= 200 * 1024;
  size_t smem = sizeof(double) * (use_smem ? nn * nn + nn + 2 : nn);
  set_smem(ulv_root_solve_kernel, smem);
  ulv_root_solve_kernel<<<s, kThreads, smem, st>>>(dn_.p, node, vals_.p, lu ? lu : fact_.p + hn_[node].F,
                                                   piv ? piv : rootpiv_.p, dB, ldB, zsol_.p, fsol_.p,
                                                   xsol_.p, s, use_smem);
  launches_++;
}

void HSSEngine::solve_bwd(const NodeLists& L, int s, double* dB, int ldB, cudaStream_t st, int nclass) {
  const int nh = nclass < 0 ? L.classes() : std::min(nclass, L.classes());
  for (int h = nh - 1; h >= 0; h--) {
    const int cnt = L.hptr[h + 1] - L.hptr[h];
    if (!cnt) continue;
    const int mm = std::max(L.max_m[h], 1);
    dim3 grid(cnt, s);
    const int* lst = L.list.p + L.hptr[h];
    if (solve_pipe_ & 1) {
      const int ldbuf = (mm + 1) & ~1;
      const int nbw = std::max(hn_[L.host[L.hptr[h]]].nbq, 1);     // panel width of the class
      const size_t psm = sizeof(double) * ((size_t)ldbuf + 64 + (size_t)2 * nbw * nbw + (size_t)2 * ldbuf * nbw);
      if (psm <= kMaxSmem) {
        set_smem(ulv_bwd_pipe_kernel, psm);
        ulv_bwd_pipe_kernel<<<grid, kThreads, psm, st>>>(dn_.p, lst, fact_.p, tfac_.p, dB, ldB, ysol_.p, xsol_.p, s,
                                                         ldbuf, nbw);
        launches_++;
        continue;
      }
    }
    size_t smem = sizeof(double) * (size_t)(mm + 2 * 32 + 8);
    set_smem(ulv_bwd_kernel<32>, smem);   // the panel width (<= 32) is a per-node field
    ulv_bwd_kernel<32><<<grid, kThreads, smem, st>>>(dn_.p, lst, fact_.p, tfac_.p, dB, ldB, ysol_.p, xsol_.p, s, (solve_pipe_ & 4) ? 1 : 0);
    launches_++;
  }
}

void HSSEngine::solve(int s, double* dB, int ldB, cudaStream_t st) {
  if (nparts_ > 1) throw std::logic_error("sharded matrix: use the dist_* entry points");
  if (!factored_) throw std::logic_error("solve called before factor");
  if (s <= 0) return;
  ensure_solve_ws(s);
  run_graphed(gkey("solve", {dB}, {s, ldB}), st, [&] {
    solve_fwd(own_, s, dB, ldB, st);     // the kernels skip the root
    solve_root(s, dB, ldB, st);
    solve_bwd(own_, s, dB, ldB, st);
  });
  SB200_CUDA(cudaGetLastError());
}

// forward_solve / backward_solve (reference HSSMatrix.solve.hpp:52-66): the
// WorkSolve state between the two calls lives in the engine's workspaces.
void HSSEngine::forward_solve(int s, double* dB, int ldB, cudaStream_t st) {
  if (nparts_ > 1) throw std::logic_error("sharded matrix: use the dist_* entry points");
  if (!factored_) throw std::logic_error("forward_solve called before factor");
  if (s <= 0) return;
  ensure_solve_ws(s);
  solve_fwd(own_, s, dB, ldB, st);
  solve_root(s, dB, ldB, st);
  fwd_s_ = s;
  SB200_CUDA(cudaGetLastError());
}

void HSSEngine::backward_solve(int s, double* dB, int ldB, cudaStream_t st) {
  if (nparts_ > 1) throw std::logic_error("sharded matrix: use the dist_* entry points");
  if (fwd_s_ != s || s <= 0) throw std::logic_error("backward_solve needs a forward_solve with the same number of right-hand sides");
  solve_bwd(own_, s, dB, ldB, st);
  SB200_CUDA(cudaGetLastError());
}

void HSSEngine::dist_solve_begin(int s, double* dB, int ldB, double* send, cudaStream_t st) {
  if (!factored_) throw std::logic_error("solve called before factor");
  ensure_solve_ws(s);
  run_graphed(gkey("solve_begin", {dB, send}, {s, ldB}), st, [&] {
    solve_fwd(own_, s, dB, ldB, st);
    const DNode& d = hn_[cut_[part_]];
    // export [z ; ft1] of the cut node, (r_v + r_u) x s
    const int ld = d.v_rank + d.u_rank;
    copy2d(send, ld, zsol_.p + (size_t)d.z_off * s, d.v_rank, d.v_rank, s, st);
    copy2d(send + d.v_rank, ld, fsol_.p + (size_t)d.f_off * s, d.u_rank, d.u_rank, s, st);
  });
  SB200_CUDA(cudaGetLastError());
}

void HSSEngine::dist_solve_end(int s, double* dB, int ldB, const double* recv, cudaStream_t st) {
  long long sz[3];
  dist_sizes(s, sz);
  run_graphed(gkey("solve_end", {dB, recv}, {s, ldB}), st, [&] {
    for (int c = 0; c < nparts_; c++) {
      const DNode& d = hn_[cut_[c]];
      const int ld = d.v_rank + d.u_rank;
      const double* src = recv + (size_t)c * sz[2];
      copy2d(zsol_.p + (size_t)d.z_off * s, d.v_rank, src, ld, d.v_rank, s, st);
      copy2d(fsol_.p + (size_t)d.f_off * s, d.u_rank, src + d.v_rank, ld, d.u_rank, s, st);
    }
    solve_fwd(top_, s, dB, ldB, st);
    solve_root(s, dB, ldB, st);
    solve_bwd(top_, s, dB, ldB, st);
    solve_bwd(own_, s, dB, ldB, st);
  });
  SB200_CUDA(cudaGetLastError());
}


// ---------------------------------------------------------------------------
// Roofline denominator, measured live: register-resident mma.sync.m8n8k4.f64
// stream (8 independent accumulators per warp, 2 CTAs of 256 threads per SM),
// the same microbenchmark as profiles/microbench/fp64_peak.cu.  Returns TFLOP/s.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters, double a, double b) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; i++) { c[i][0] = threadIdx.x * 1e-3 + i; c[i][1] = i; }
  const double av = a + threadIdx.x * 1e-9, bv = b;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) dmma(c[i][0], c[i][1], av, bv);
  }
  double s = 0.;
#pragma unroll
  for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

double measure_fp64_dmma_peak_tflops() {
  int dev = 0, nsm = 0;
  SB200_CUDA(cudaGetDevice(&dev));
  SB200_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  const int blocks = 2 * nsm, threads = 256, iters = 20000;
  DevBuf<double> out((size_t)blocks * threads);
  cudaEvent_t e0, e1;
  SB200_CUDA(cudaEventCreate(&e0));
  SB200_CUDA(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int r = 0; r < 4; r++) {   // first pass warms up
    SB200_CUDA(cudaEventRecord(e0, 0));
    dmma_peak_kernel<<<blocks, threads>>>(out.p, iters, 1.0000001, 1e-9);
    SB200_CUDA(cudaEventRecord(e1, 0));
    SB200_CUDA(cudaEventSynchronize(e1));
    float t = 0.f;
    SB200_CUDA(cudaEventElapsedTime(&t, e0, e1));
    if (r) best = std::min(best, t);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  const double flops = 2. * 8 * 8 * 4 * 8. * iters * (double)(blocks * threads / 32);
  return flops / (best * 1e-3) / 1e12;
}

// ---------------------------------------------------------------------------
// Test / microbenchmark hook: `count` copies of one m x naug block (QR of the
// first k columns) through the leaf QR kernels, outside any HSS tree.
// variant 0: ulv_qr_kernel<16, true, 128> (right-looking), 1: qr3::ulv_qr3_kernel
// ---------------------------------------------------------------------------
void debug_qr_batch(int m, int k, int naug, int count, const double* hA, double* hOut, double* hT, int variant,
                    int reps, float* ms) {
  if (m < 1 || m > 256 || k < 0 || k > m || naug < k || count < 1)
    throw std::invalid_argument("debug_qr_batch: need 1 <= m <= 256, 0 <= k <= m <= ..., naug >= k");
  const long long fsz = (((long long)m * naug) + 1) & ~1LL, tsz = 32LL * std::max(k, 1);
  std::vector<DNode> nodes(count);
  std::vector<int> list(count);
  for (int i = 0; i < count; i++) {
    DNode d{};
    d.parent = 0; d.leaf = 1; d.m = m; d.k = k; d.naug = naug; d.nbq = 16;
    d.F = fsz * i; d.T = tsz * i;
    nodes[i] = d; list[i] = i;
  }
  DevBuf<DNode> dn; dn.upload(nodes.data(), nodes.size());
  DevBuf<int> dl; dl.upload(list.data(), list.size());
  DevBuf<double> src((size_t)fsz * count), fact((size_t)fsz * count), tf((size_t)tsz * count);
  SB200_CUDA(cudaMemset(src.p, 0, sizeof(double) * fsz * count));
  for (int i = 0; i < count; i++)
    SB200_CUDA(cudaMemcpy(src.p + fsz * i, hA, sizeof(double) * (size_t)m * naug, cudaMemcpyHostToDevice));
  DevBuf<unsigned char> tmaps;
  if (variant == 1) build_qr3_tmaps(nodes, fact.p, tmaps);
  cudaEvent_t e0, e1;
  SB200_CUDA(cudaEventCreate(&e0));
  SB200_CUDA(cudaEventCreate(&e1));
  float best = 1e30f;
  const int ldv = smem_ld(m);
  for (int r = 0; r < std::max(reps, 1); r++) {
    SB200_CUDA(cudaMemcpy(fact.p, src.p, sizeof(double) * fsz * count, cudaMemcpyDeviceToDevice));
    SB200_CUDA(cudaMemset(tf.p, 0, sizeof(double) * tsz * count));
    SB200_CUDA(cudaEventRecord(e0, 0));
    if (variant == 1) {
      set_smem(qr3::ulv_qr3_kernel, qr3::SMEM_BYTES);
      qr3::ulv_qr3_kernel<<<count, qr3::NTHREADS, qr3::SMEM_BYTES>>>(dn.p, dl.p, fact.p, tf.p, tmaps.p);
    } else {
      size_t smem = qr_smem<16>(ldv);
      set_smem(ulv_qr_kernel<16, true, 128>, smem);
      ulv_qr_kernel<16, true, 128><<<count, 128, smem>>>(dn.p, dl.p, fact.p, tf.p, ldv, 0);
    }
    SB200_CUDA(cudaEventRecord(e1, 0));
    SB200_CUDA(cudaEventSynchronize(e1));
    SB200_CUDA(cudaGetLastError());
    float t = 0.f;
    SB200_CUDA(cudaEventElapsedTime(&t, e0, e1));
    best = std::min(best, t);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (ms) *ms = best;
  // the last copy (its neighbours ran concurrently: races between CTAs would show)
  if (hOut) SB200_CUDA(cudaMemcpy(hOut, fact.p + fsz * (count - 1), sizeof(double) * (size_t)m * naug, cudaMemcpyDeviceToHost));
  if (hT) SB200_CUDA(cudaMemcpy(hT, tf.p + tsz * (count - 1), sizeof(double) * 16 * (size_t)std::max(k, 1), cudaMemcpyDeviceToHost));
}

}  // namespace sb200
