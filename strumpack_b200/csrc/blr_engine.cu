// strumpack_b200 -- block low-rank (BLR) matrices on sm_100a:
// tile compression, right-looking BLR LU, triangular solves, mat-vec.
//
// Replaces (reference):
//   BLRMatrix::compress / compress_and_factor (RL)  src/BLR/BLRMatrix.cpp:91-241
//   LRTile(RRQR) / DenseTile::LU                     src/BLR/LRTile.cpp:55-75, DenseTile.cpp:111-117
//   trsm / laswp on tiles                            src/BLR/LRTile.cpp:296-311, BLRTileBLAS.hpp
//   BLRMatrix::solve, trsm(L,L/U), gemv              src/BLR/BLRMatrix.hpp:118-122, .cpp:1667-1763
// and the library-call GPU path construct_and_partial_factor_gpu
// (src/BLR/BLRMatrix.GPU.cpp:70-262, src/BLR/BLRBatch.cpp) -- here every step
// is a hand-written batched kernel over the tiles of a block row/column:
//   K11 blr_copy_tiles + id_cpqr (sb200_cpqr.cuh) + blr_extract_lr : RRQR tiles
//   K12 blr_getrf (+pivot threshold), blr_trsm_lower (L^{-1} P U_ij, U^{-T} V_ji^T)
//   K13 blr_schur : C_kj -= U_ki (V_ki U_ij) V_ij, fused, fp64 tensor pipe
//   K14 blr_lr_gemv / blr_trsm_lower on the right-hand side
#include "blr_engine.hpp"

#include <algorithm>
#include <cstring>
#include <stdexcept>

#include "sb200_cpqr.cuh"

namespace sb200 {

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = 8;

struct TileDesc {     // one off-diagonal tile of the current step
  int i, j;           // tile coordinates
  int ro, co, m, n;   // row/col offset and size in the matrix
  long long lr;       // offset of [U (m x rc) | Vt (n x rc)] in the LR arena
  int rc;             // rank capacity
  int adm;            // 0: inadmissible tile, stays dense whatever its numerical rank
};

// ---------------------------------------------------------------- diagonal LU
// Blocked LU with partial pivoting of one diagonal tile, in place (ld = N), plus
// the reference's small-pivot replacement (DenseTile::LU, DenseTile.cpp:111-117).
// One CTA; the NB-column panel is factored in shared memory (one latency chain
// per column), the row swaps are applied to the rest of the tile, U12 =
// L11^{-1} A12 is formed in shared memory and the trailing block is updated on
// the fp64 tensor pipe with L21 (smem) x U12 (smem), C streamed from global.
// ipiv: LAPACK style (1-based, local); g: the same permutation as a gather.
template <int NB>
__global__ void __launch_bounds__(kThreads)
blr_getrf_kernel(double* __restrict__ A, long long ld, int n, int* __restrict__ ipiv,
                 int* __restrict__ g, double thresh, int ldp) {
  extern __shared__ double sm[];
  double* P = sm;                        // ldp x NB panel (rows j0..n)
  double* U12 = P + (size_t)ldp * NB;    // NB x (n - j0 - NB), ld = NB + 4... stored [k + c*LDU]
  constexpr int LDU = NB + 4;
  __shared__ double rv[kWarps];
  __shared__ int ri[kWarps];
  __shared__ int pivrow;
  __shared__ int pl[NB];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int j0 = 0; j0 < n; j0 += NB) {
    const int jb = min(NB, n - j0), mp = n - j0;
    for (int c = warp; c < jb; c += kWarps)
      for (int i = lane; i < mp; i += 32) P[i + c * ldp] = A[(j0 + i) + (j0 + c) * ld];
    __syncthreads();
    for (int c = 0; c < jb; c++) {
      double best = -1.;
      int bi = c;
      for (int i = c + tid; i < mp; i += kThreads) {
        double v = fabs(P[i + c * ldp]);
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
        pl[c] = p;
        ipiv[j0 + c] = j0 + p + 1;
      }
      __syncthreads();
      const int p = pivrow;
      if (p != c && tid < jb) {
        double t = P[c + tid * ldp];
        P[c + tid * ldp] = P[p + tid * ldp];
        P[p + tid * ldp] = t;
      }
      __syncthreads();
      const double d = P[c + c * ldp];
      const double inv = d != 0. ? 1. / d : 0.;
      // rank-1 update of the remaining panel columns, one warp per column; the
      // multipliers are recomputed on the fly and stored by the last warp pass
      for (int cc = c + 1 + warp; cc < jb; cc += kWarps) {
        const double u = P[c + cc * ldp];
        for (int i = c + 1 + lane; i < mp; i += 32) P[i + cc * ldp] -= (P[i + c * ldp] * inv) * u;
      }
      __syncthreads();
      for (int i = c + 1 + tid; i < mp; i += kThreads) P[i + c * ldp] *= inv;
      __syncthreads();
    }
    // row swaps on the columns outside the panel
    for (int col = tid; col < n; col += kThreads) {
      if (col >= j0 && col < j0 + jb) continue;
      double* a = A + col * ld + j0;
      for (int c = 0; c < jb; c++) {
        const int p = pl[c];
        if (p != c) { double t = a[c]; a[c] = a[p]; a[p] = t; }
      }
    }
    // panel back to global
    for (int c = warp; c < jb; c += kWarps)
      for (int i = lane; i < mp; i += 32) A[(j0 + i) + (j0 + c) * ld] = P[i + c * ldp];
    __syncthreads();
    const int nt = n - j0 - jb;        // trailing columns
    if (nt <= 0) break;
    // U12 = L11^{-1} A12, one thread per column, result to smem and global
    for (int col = tid; col < nt; col += kThreads) {
      double* a = A + (j0 + jb + col) * ld + j0;
      double x[NB];
#pragma unroll
      for (int k = 0; k < NB; k++) x[k] = k < jb ? a[k] : 0.;
#pragma unroll
      for (int k = 0; k < NB; k++) {
#pragma unroll
        for (int i = 0; i < NB; i++)
          if (i > k && i < jb) x[i] -= P[i + k * ldp] * x[k];
      }
#pragma unroll
      for (int k = 0; k < NB; k++) {
        if (k < jb) a[k] = x[k];
        U12[k + col * LDU] = x[k];
      }
    }
    __syncthreads();
    // A22 -= L21 U12 : 16x16 output tiles over the warps, K = jb
    {
      const int mr = mp - jb;
      const int gq = lane >> 2, t = lane & 3;
      const int tm = (mr + 15) >> 4, tn = (nt + 15) >> 4;
      double* C = A + (j0 + jb) + (j0 + jb) * ld;
      const double* L21 = P + jb;
      for (int tile = warp; tile < tm * tn; tile += kWarps) {
        const int i0 = (tile % tm) << 4, c0 = (tile / tm) << 4;
        double c[2][2][2];
#pragma unroll
        for (int ti = 0; ti < 2; ti++)
#pragma unroll
          for (int tj = 0; tj < 2; tj++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
              const int ii = i0 + ti * 8 + gq, jj = c0 + tj * 8 + 2 * t + e;
              c[ti][tj][e] = (ii < mr && jj < nt) ? C[ii + jj * ld] : 0.;
            }
#pragma unroll
        for (int k0 = 0; k0 < NB; k0 += 4) {
          const int kk = k0 + t;
          const bool kin = kk < jb;
          const int ia0 = i0 + gq, ia1 = ia0 + 8, jb0 = c0 + gq, jb1 = jb0 + 8;
          const double a0 = (kin && ia0 < mr) ? -L21[ia0 + kk * ldp] : 0.;
          const double a1 = (kin && ia1 < mr) ? -L21[ia1 + kk * ldp] : 0.;
          const double b0 = (kin && jb0 < nt) ? U12[kk + jb0 * LDU] : 0.;
          const double b1 = (kin && jb1 < nt) ? U12[kk + jb1 * LDU] : 0.;
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
              const int ii = i0 + ti * 8 + gq, jj = c0 + tj * 8 + 2 * t + e;
              if (ii < mr && jj < nt) C[ii + jj * ld] = c[ti][tj][e];
            }
      }
    }
    __syncthreads();
  }
  if (thresh > 0.)
    for (int i = tid; i < n; i += kThreads) {
      double d = A[i + i * ld];
      if (fabs(d) < thresh) A[i + i * ld] = d < 0 ? -thresh : thresh;
    }
  __syncthreads();
  if (tid == 0) {
    for (int i = 0; i < n; i++) g[i] = i;
    for (int i = 0; i < n; i++) {
      int p = ipiv[i] - 1;
      if (p != i) { int t = g[i]; g[i] = g[p]; g[p] = t; }
    }
  }
}

// UT = (upper triangle of A)^T, so that solves with U from the right become
// solves with the lower triangular U^T from the left
__global__ void blr_upper_transpose_kernel(const double* __restrict__ A, long long ld, int n,
                                           double* __restrict__ UT) {
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n * n; idx += gridDim.x * blockDim.x) {
    int r = idx % n, c = idx / n;   // UT[r, c] = U[c, r] for c <= r
    UT[r + (size_t)c * n] = (c <= r) ? A[c + r * ld] : 0.;
  }
}

// ------------------------------------------------------------ tile compression
__global__ void __launch_bounds__(kThreads)
blr_copy_tiles_kernel(const double* __restrict__ A, long long ld, const TileDesc* __restrict__ tiles,
                      double* __restrict__ scratch, long long stride) {
  const TileDesc t = tiles[blockIdx.x];
  double* dst = scratch + blockIdx.x * stride;
  const double* src = A + t.ro + (size_t)t.co * ld;
  for (int c = blockIdx.y * kWarps + (threadIdx.x >> 5); c < t.n; c += gridDim.y * kWarps)
    for (int r = threadIdx.x & 31; r < t.m; r += 32) dst[r + (size_t)c * t.m] = src[r + c * ld];
}

// U = Q[:, pivots] (m x rank), Vt = R^T (n x rank)   (LRTile RRQR ctor, LRTile.cpp:55-65)
__global__ void __launch_bounds__(kThreads)
blr_extract_lr_kernel(const TileDesc* __restrict__ tiles, const double* __restrict__ scratch,
                      long long stride, const double* __restrict__ R, long long rstride,
                      const int* __restrict__ order, int ostride, const int* __restrict__ ranks_step,
                      double* __restrict__ lr, int* __restrict__ rank_tab, int nb) {
  const TileDesc t = tiles[blockIdx.x];
  const int rank = t.adm ? ranks_step[blockIdx.x] : -1;   // inadmissible: a DenseTile (BLRMatrix.cpp:146-147)
  if (threadIdx.x == 0) rank_tab[t.i + t.j * nb] = rank;
  if (rank <= 0) return;
  const double* Q = scratch + blockIdx.x * stride;
  const double* Rt = R + blockIdx.x * rstride;
  const int* ord = order + (size_t)blockIdx.x * ostride;
  double* U = lr + t.lr;
  double* Vt = U + (size_t)t.m * t.rc;
  for (int idx = threadIdx.x; idx < t.m * rank; idx += kThreads) {
    int r = idx % t.m, a = idx / t.m;
    U[r + (size_t)a * t.m] = Q[r + (size_t)ord[a] * t.m];
  }
  for (int idx = threadIdx.x; idx < t.n * rank; idx += kThreads) {
    int c = idx % t.n, a = idx / t.n;
    Vt[c + (size_t)a * t.n] = Rt[a + (size_t)c * t.rc];
  }
}

// ------------------------------------------------------------ triangular solves
// B <- L^{-1} B[g, :] for the columns of B, L lower triangular (unit or not),
// one warp per column, blocked by 32 rows.  Used for
//   U_ij <- L_ii^{-1} P_i U_ij         (laswp + trsm L,L,N,U; BLRMatrix.cpp:152-155)
//   Vt_ji <- U_ii^{-T} Vt_ji           (trsm R,U,N,N on V;     BLRMatrix.cpp:165-166)
//   and for the right-hand sides of the solve.
__global__ void __launch_bounds__(kThreads)
blr_trsm_lower_kernel(const double* __restrict__ L, long long ldl, int n, int unit,
                      const int* __restrict__ g, const SolveTask* __restrict__ tasks,
                      const int* __restrict__ ncols_dyn) {
  extern __shared__ double sm[];
  const SolveTask t = tasks[blockIdx.x];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ncols = ncols_dyn ? ncols_dyn[blockIdx.x] : t.ncols;
  const int col = blockIdx.y * kWarps + warp;
  if (col >= ncols) return;
  double* x = sm + (size_t)warp * n;
  double* b = t.B + (size_t)col * t.ldb;
  for (int i = lane; i < n; i += 32) x[i] = b[g ? g[i] : i];
  __syncwarp();
  for (int i0 = 0; i0 < n; i0 += 32) {
    const int i = i0 + lane;
    double v = i < n ? x[i] : 0.;
    if (i < n) {
      double a0 = 0., a1 = 0.;
      int j = 0;
      for (; j + 2 <= i0; j += 2) {
        a0 += L[i + j * ldl] * x[j];
        a1 += L[i + (j + 1) * ldl] * x[j + 1];
      }
      v -= a0 + a1;
    }
    const int ib = min(32, n - i0);
    for (int a = 0; a < ib; a++) {
      if (!unit && lane == a) v /= L[(i0 + a) + (i0 + a) * ldl];
      const double xa = __shfl_sync(0xffffffffu, v, a);
      if (lane > a && i < n) v -= L[i + (i0 + a) * ldl] * xa;
    }
    if (i < n) x[i] = v;
    __syncwarp();
  }
  for (int i = lane; i < n; i += 32) b[i] = x[i];
}

// B <- U^{-1} B (back substitution, U upper triangular non-unit), one warp per column
__global__ void __launch_bounds__(kThreads)
blr_trsm_upper_kernel(const double* __restrict__ U, long long ldu, int n, double* __restrict__ B,
                      long long ldb, int ncols) {
  extern __shared__ double sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col = blockIdx.x * kWarps + warp;
  if (col >= ncols) return;
  double* x = sm + (size_t)warp * n;
  double* b = B + (size_t)col * ldb;
  for (int i = lane; i < n; i += 32) x[i] = b[i];
  __syncwarp();
  const int nblk = (n + 31) / 32;
  for (int bi = nblk - 1; bi >= 0; bi--) {
    const int i0 = bi * 32, i = i0 + lane, ib = min(32, n - i0);
    double v = i < n ? x[i] : 0.;
    if (i < n) {
      double a0 = 0.;
      for (int j = i0 + ib; j < n; j++) a0 += U[i + j * ldu] * x[j];
      v -= a0;
    }
    for (int a = ib - 1; a >= 0; a--) {
      if (lane == a) v /= U[(i0 + a) + (i0 + a) * ldu];
      const double xa = __shfl_sync(0xffffffffu, v, a);
      if (lane < a) v -= U[i + (i0 + a) * ldu] * xa;
    }
    if (i < n) x[i] = v;
    __syncwarp();
  }
  for (int i = lane; i < n; i += 32) b[i] = x[i];
}

// ranks of the tiles of one kind (0: slots with i<j, 1: slots with i>j) as the
// dynamic column counts of the trsm launch
// kind 2: the dense (incompressible, rank_tab < 0) tiles with i<j: all their columns
__global__ void blr_select_kernel(const TileDesc* __restrict__ td, int cnt, int kind,
                                  const int* __restrict__ rank_tab, int nb, int* __restrict__ sel) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= cnt) return;
  const bool upper = td[q].i < td[q].j;
  const int r = rank_tab[td[q].i + td[q].j * nb];
  if (kind == 2) sel[q] = (upper && r < 0) ? td[q].n : 0;
  else sel[q] = ((kind == 0) == upper && r > 0) ? r : 0;
}

// Dense tiles below the diagonal: A_ji <- A_ji U_ii^{-1} (trsm R,U,N,N on a
// DenseTile, reference BLRTileBLAS / DenseTile::trsm_b).  One CTA per tile of
// the step, one thread per row; tiles that are low rank return at once.
__global__ void __launch_bounds__(kThreads)
blr_trsm_right_upper_dense_kernel(double* __restrict__ A, long long ld, const double* __restrict__ U,
                                  int nu, const TileDesc* __restrict__ tiles,
                                  const int* __restrict__ rank_tab, int nb) {
  const TileDesc t = tiles[blockIdx.x];
  if (t.i < t.j || rank_tab[t.i + t.j * nb] >= 0) return;
  double* B = A + t.ro + (size_t)t.co * ld;   // t.m x nu
  for (int r = threadIdx.x; r < t.m; r += kThreads) {
    for (int c = 0; c < nu; c++) {
      double v = B[r + (size_t)c * ld];
      for (int a = 0; a < c; a++) v -= B[r + (size_t)a * ld] * U[a + (size_t)c * ld];
      B[r + (size_t)c * ld] = v / U[c + (size_t)c * ld];
    }
  }
}

// ------------------------------------------------------------------ Schur update
// C_kj -= U_ki (Vt_ki^T U_ij) Vt_ij^T for every pair (k, j) of the trailing
// matrix, one CTA per pair (RL update, BLRMatrix.cpp:170-185; the 3-stage
// batched GEMM of BLRBatch.hpp:85-118 fused: the rank x rank core never leaves
// shared memory).  Ranks above KC are processed in chunks of KC columns.
// mode 0 (right-looking): the CTAs are the (k, j) pairs behind step `pstep`,
//   each subtracts the contribution of that one step;
// mode 1 (left-looking, BLRFactorAlgorithm::LL, BLRMatrix.cpp:186-212): the
//   CTAs are the tiles of block row and block column `pstep` (diagonal tile
//   included), each subtracts the contributions of ALL earlier steps, in the
//   same order and with the same arithmetic as mode 0 -- the factors are
//   bitwise those of the right-looking schedule;
// mode 2: the trailing block behind a partial factorization of `pstep` steps
//   (the LL update of A22, BLRMatrix.cpp:998-1013).
template <int KC>
__global__ void __launch_bounds__(kThreads)
blr_schur_kernel(double* A, long long ld, const int* __restrict__ off, int nb, int pstep,
                 const double* __restrict__ lr, const long long* __restrict__ lroff,
                 const int* __restrict__ rcap, const int* __restrict__ rank_tab, int ldp,
                 const double* __restrict__ ident, int ldi, int mode) {
  extern __shared__ double sm[];
  int k, j, sbeg, send;
  if (mode == 0) {
    const int nrem = nb - pstep - 1;
    k = pstep + 1 + blockIdx.x % nrem; j = pstep + 1 + blockIdx.x / nrem;
    sbeg = pstep; send = pstep + 1;
  } else if (mode == 1) {
    const int nrow = nb - pstep;          // tiles (pstep, pstep .. nb-1), then (pstep+1 .. nb-1, pstep)
    if ((int)blockIdx.x < nrow) { k = pstep; j = pstep + blockIdx.x; }
    else { k = pstep + 1 + (blockIdx.x - nrow); j = pstep; }
    sbeg = 0; send = pstep;
  } else {
    const int nrem = nb - pstep;
    k = pstep + blockIdx.x % nrem; j = pstep + blockIdx.x / nrem;
    sbeg = 0; send = pstep;
  }
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int istep = sbeg; istep < send; istep++) {
  int ra = rank_tab[k + istep * nb], rb = rank_tab[istep + j * nb];
  if (ra == 0 || rb == 0) continue;    // zero tile: nothing to subtract
  __syncthreads();                     // the shared buffers of the previous step are free
  const int mk = off[k + 1] - off[k], mi = off[istep + 1] - off[istep], nj = off[j + 1] - off[j];
  constexpr int LDM = KC + 4;
  double* S1 = sm;                       // ldp x KC
  double* S2 = S1 + (size_t)ldp * KC;    // ldp x KC
  double* W = S2 + (size_t)ldp * KC;     // ldp x KC
  double* Ms = W + (size_t)ldp * KC;     // LDM x KC
  // a tile is U Vt^T; a dense (incompressible) tile T is taken as T I^T with
  // U = the tile itself inside A and Vt = the identity (DenseTile operands of
  // the reference's gemm(LRTile/DenseTile, ...) overloads, BLRTileBLAS.hpp)
  const double* Uki = lr + lroff[k + istep * nb];
  const double* Vtki = Uki + (size_t)mk * rcap[k + istep * nb];
  long long ldUki = mk, ldVki = mi;
  if (ra < 0) { Uki = A + off[k] + (size_t)off[istep] * ld; ldUki = ld; Vtki = ident; ldVki = ldi; ra = mi; }
  const double* Uij = lr + lroff[istep + j * nb];
  const double* Vtij = Uij + (size_t)mi * rcap[istep + j * nb];
  long long ldUij = mi, ldVij = nj;
  if (rb < 0) { Uij = A + off[istep] + (size_t)off[j] * ld; ldUij = ld; Vtij = ident; ldVij = ldi; rb = nj; }
  double* C = A + off[k] + (size_t)off[j] * ld;
  for (int cb = 0; cb < rb; cb += KC) {
    const int kb = min(KC, rb - cb);
    for (int ca = 0; ca < ra; ca += KC) {
      const int ka = min(KC, ra - ca);
      __syncthreads();
      for (int c = warp; c < ka; c += kWarps)
        for (int r = lane; r < mi; r += 32) S1[r + c * ldp] = Vtki[r + (size_t)(ca + c) * ldVki];
      for (int c = warp; c < kb; c += kWarps)
        for (int r = lane; r < mi; r += 32) S2[r + c * ldp] = Uij[r + (size_t)(cb + c) * ldUij];
      __syncthreads();
      smem_gemm<true, false>(ka, kb, mi, 1., S1, ldp, S2, ldp, 0., Ms, LDM, warp, kWarps, lane);
      __syncthreads();
      for (int c = warp; c < ka; c += kWarps)
        for (int r = lane; r < mk; r += 32) S1[r + c * ldp] = Uki[r + (size_t)(ca + c) * ldUki];
      __syncthreads();
      smem_gemm<false, false>(mk, kb, ka, 1., S1, ldp, Ms, LDM, ca ? 1. : 0., W, ldp, warp, kWarps, lane);
    }
    __syncthreads();
    for (int c = warp; c < kb; c += kWarps)
      for (int r = lane; r < nj; r += 32) S2[r + c * ldp] = Vtij[r + (size_t)(cb + c) * ldVij];
    __syncthreads();
    // C (mk x nj, global) -= W (mk x kb) * S2^T (kb x nj): 16x16 tiles over the warps
    const int g = lane >> 2, t = lane & 3;
    const int tm = (mk + 15) >> 4, tn = (nj + 15) >> 4;
    for (int tile = warp; tile < tm * tn; tile += kWarps) {
      const int i0 = (tile % tm) << 4, j0 = (tile / tm) << 4;
      double c[2][2][2];
#pragma unroll
      for (int ti = 0; ti < 2; ti++)
#pragma unroll
        for (int tj = 0; tj < 2; tj++)
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const int ii = i0 + ti * 8 + g, jj = j0 + tj * 8 + 2 * t + e;
            c[ti][tj][e] = (ii < mk && jj < nj) ? C[ii + (size_t)jj * ld] : 0.;
          }
      for (int k0 = 0; k0 < kb; k0 += 4) {
        const int kk = k0 + t;
        const bool kin = kk < kb;
        const int ia0 = i0 + g, ia1 = ia0 + 8, jb0 = j0 + g, jb1 = jb0 + 8;
        const double a0 = (kin && ia0 < mk) ? -W[ia0 + kk * ldp] : 0.;
        const double a1 = (kin && ia1 < mk) ? -W[ia1 + kk * ldp] : 0.;
        const double b0 = (kin && jb0 < nj) ? S2[jb0 + kk * ldp] : 0.;
        const double b1 = (kin && jb1 < nj) ? S2[jb1 + kk * ldp] : 0.;
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
            const int ii = i0 + ti * 8 + g, jj = j0 + tj * 8 + 2 * t + e;
            if (ii < mk && jj < nj) C[ii + (size_t)jj * ld] = c[ti][tj][e];
          }
    }
  }
  }   // istep
}

// ------------------------------------------------------------- vector kernels
// y_k += alpha * T_kj x_j for a list of LR tiles (gemv_a, LRTile.cpp:313-327),
// one CTA per tile, atomics into y (several tiles feed the same y block).
struct GemvTask { int k, j; };

__global__ void __launch_bounds__(kThreads)
blr_lr_gemv_kernel(const GemvTask* __restrict__ tasks, const int* __restrict__ off, int nb,
                   const double* __restrict__ lr, const long long* __restrict__ lroff,
                   const int* __restrict__ rcap, const int* __restrict__ rank_tab,
                   const double* __restrict__ x, long long ldx, double* __restrict__ y,
                   long long ldy, double alpha, const double* __restrict__ A, long long ld,
                   int trans = 0) {
  extern __shared__ double sm[];
  const GemvTask tk = tasks[blockIdx.x];
  const int col = blockIdx.y;
  const int r = rank_tab[tk.k + tk.j * nb];
  if (r == 0) return;
  const int tm = off[tk.k + 1] - off[tk.k], tn = off[tk.j + 1] - off[tk.j];
  // trans: y_j += alpha * T_kj^T x_k  (T^T = Vt U^T: the two factors swap roles)
  const int m = trans ? tn : tm, n = trans ? tm : tn;
  const int xb = trans ? tk.k : tk.j, yb = trans ? tk.j : tk.k;
  if (r < 0) {   // dense tile, lives in A (DenseTile::gemv_a)
    const double* D = A + off[tk.k] + (size_t)off[tk.j] * ld;
    const double* xj = x + off[xb] + col * ldx;
    double* yk = y + off[yb] + col * ldy;
    if (!trans) {
      for (int i = threadIdx.x; i < m; i += kThreads) {
        double acc = 0.;
        for (int c = 0; c < n; c++) acc += D[i + (size_t)c * ld] * xj[c];
        atomicAdd(yk + i, alpha * acc);
      }
    } else {
      const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
      for (int c = warp; c < m; c += kWarps) {      // column c of D = row c of D^T
        double acc = 0.;
        for (int i = lane; i < n; i += 32) acc += D[i + (size_t)c * ld] * xj[i];
        acc = warp_sum(acc);
        if (lane == 0) atomicAdd(yk + c, alpha * acc);
      }
    }
    return;
  }
  const double* U0 = lr + lroff[tk.k + tk.j * nb];
  const double* Vt0 = U0 + (size_t)tm * rcap[tk.k + tk.j * nb];
  const double* U = trans ? Vt0 : U0;      // m x r
  const double* Vt = trans ? U0 : Vt0;     // n x r
  const double* xj = x + off[xb] + col * ldx;
  double* yk = y + off[yb] + col * ldy;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  double* tv = sm;   // r
  for (int a = warp; a < r; a += kWarps) {
    const double* va = Vt + (size_t)a * n;
    double acc = 0.;
    for (int i = lane; i < n; i += 32) acc += va[i] * xj[i];
    acc = warp_sum(acc);
    if (lane == 0) tv[a] = alpha * acc;
  }
  __syncthreads();
  for (int i = tid; i < m; i += kThreads) {
    double acc = 0.;
    for (int a = 0; a < r; a++) acc += U[i + (size_t)a * m] * tv[a];
    atomicAdd(yk + i, acc);
  }
}

// y_i += D_ii x_i (dense diagonal tiles of a compressed, unfactored matrix)
__global__ void __launch_bounds__(kThreads)
blr_diag_gemv_kernel(const double* __restrict__ A, long long ld, const int* __restrict__ off,
                     const double* __restrict__ x, long long ldx, double* __restrict__ y,
                     long long ldy, int trans) {
  const int i = blockIdx.x, col = blockIdx.y;
  const int m = off[i + 1] - off[i];
  const double* D = A + off[i] + (size_t)off[i] * ld;
  const double* xi = x + off[i] + col * ldx;
  double* yi = y + off[i] + col * ldy;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (!trans) {
    for (int r = threadIdx.x; r < m; r += kThreads) {
      double acc = 0.;
      for (int c = 0; c < m; c++) acc += D[r + c * ld] * xi[c];
      atomicAdd(yi + r, acc);
    }
  } else {
    for (int c = warp; c < m; c += kWarps) {
      double acc = 0.;
      for (int r = lane; r < m; r += 32) acc += D[r + c * ld] * xi[r];
      acc = warp_sum(acc);
      if (lane == 0) atomicAdd(yi + c, acc);
    }
  }
}

__global__ void blr_zero_kernel(double* y, long long ld, int n, int s) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < (long long)n * s;
       idx += (long long)gridDim.x * blockDim.x)
    y[(idx % n) + (idx / n) * ld] = 0.;
}

}  // namespace

// ===========================================================================
//                                 host side
// ===========================================================================
template <typename K> static void set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024)
    SB200_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}

static void refine(std::vector<int>& tiles, int size, int leaf) {
  // ClusterTree::refine (reference src/structured/ClusterTree.hpp:104-114)
  if (size >= 2 * leaf) {
    refine(tiles, size / 2, leaf);
    refine(tiles, size - size / 2, leaf);
  } else tiles.push_back(size);
}

BLREngine::BLREngine(int n, const double* hostA, int ldA, const BLROpts& o, bool do_factor,
                     bool device_input)
    : n_(n), opts_(o) {
  if (n <= 0) throw std::invalid_argument("empty matrix");
  std::vector<int> tiles;
  if (!o.tiles1.empty()) {
    long long sum = 0;
    for (int t : o.tiles1) { if (t <= 0) throw std::invalid_argument("BLR: tile sizes must be positive"); sum += t; }
    if (sum != n) throw std::invalid_argument("BLR: the given tiles do not add up to the matrix size");
    tiles = o.tiles1;
  } else refine(tiles, n, std::max(1, o.leaf_size));
  n1_ = n;
  nsteps_ = int(tiles.size());
  setup(tiles, do_factor);
  SB200_CUDA(cudaMemcpy2D(A_.p, sizeof(double) * n, hostA, sizeof(double) * ldA,
                          sizeof(double) * n, n,
                          device_input ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
  run(do_factor);
}

BLREngine::BLREngine(int n1, int n2, const double* A11, int ld11, const double* A12, int ld12,
                     const double* A21, int ld21, const double* A22, int ld22, const BLROpts& o,
                     bool device_input)
    : n_(n1 + n2), opts_(o) {
  if (n1 <= 0 || n2 < 0) throw std::invalid_argument("partial factorization needs n1 > 0, n2 >= 0");
  std::vector<int> tiles;
  auto given = [&](const std::vector<int>& t, int n) {
    long long sum = 0;
    for (int v : t) { if (v <= 0) throw std::invalid_argument("BLR: tile sizes must be positive"); sum += v; }
    if (sum != n) throw std::invalid_argument("BLR: the given tiles do not add up to the block size");
    tiles.insert(tiles.end(), t.begin(), t.end());
  };
  if (!o.tiles1.empty()) given(o.tiles1, n1);
  else refine(tiles, n1, std::max(1, o.leaf_size));
  n1_ = n1;
  nsteps_ = int(tiles.size());
  if (n2 > 0) {
    if (!o.tiles2.empty()) given(o.tiles2, n2);
    else refine(tiles, n2, std::max(1, o.leaf_size));
  }
  setup(tiles, true);
  const cudaMemcpyKind kind = device_input ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  const size_t n = (size_t)n_, w = sizeof(double);
  SB200_CUDA(cudaMemcpy2D(A_.p, w * n, A11, w * ld11, w * n1, n1, kind));
  if (n2 > 0) {
    SB200_CUDA(cudaMemcpy2D(A_.p + (size_t)n1 * n, w * n, A12, w * ld12, w * n1, n2, kind));
    SB200_CUDA(cudaMemcpy2D(A_.p + n1, w * n, A21, w * ld21, w * n2, n1, kind));
    SB200_CUDA(cudaMemcpy2D(A_.p + n1 + (size_t)n1 * n, w * n, A22, w * ld22, w * n2, n2, kind));
  }
  run(true);
}

void BLREngine::setup(const std::vector<int>& tiles, bool do_factor) {
  const int n = n_;
  nb_ = int(tiles.size());
  off_.assign(nb_ + 1, 0);
  for (int t = 0; t < nb_; t++) off_[t + 1] = off_[t] + tiles[t];
  maxtile_ = *std::max_element(tiles.begin(), tiles.end());
  if (!opts_.admissible.empty() &&
      (opts_.nadm != nsteps_ || opts_.admissible.size() != (size_t)opts_.nadm * opts_.nadm))
    throw std::invalid_argument("BLR admissibility matrix must be (tiles of the eliminated block)^2 = " +
                                std::to_string(nsteps_) + "^2");
  if (maxtile_ > 1024) throw std::invalid_argument("BLR tiles larger than 1024 are not supported");
  // LR arena: tile (i,j) gets [U m x rc | Vt n x rc], rc = min(m,n)/2 (a tile
  // is kept low-rank only if rank*(m+n) <= m*n, BLRMatrix.cpp:563-570).  In a
  // partial factorization the tiles of the trailing (2,2) block stay dense.
  lroff_.assign((size_t)nb_ * nb_, -1);
  rcap_.assign((size_t)nb_ * nb_, 0);
  long long o_ = 0;
  for (int j = 0; j < nb_; j++)
    for (int i = 0; i < nb_; i++) {
      if (i == j) continue;
      if (do_factor && std::min(i, j) >= nsteps_) continue;
      const int m = tiles[i], nn = tiles[j];
      const int rc = std::max(1, std::min(std::min(m, nn) / 2, opts_.max_rank));
      lroff_[i + (size_t)j * nb_] = o_;
      rcap_[i + (size_t)j * nb_] = rc;
      o_ += (long long)(m + nn) * rc;
    }
  A_.alloc((size_t)n * n);
  lr_.alloc((size_t)std::max<long long>(o_, 1));
  doff_.upload(off_.data(), off_.size());
  dlroff_.upload(lroff_.data(), lroff_.size());
  drcap_.upload(rcap_.data(), rcap_.size());
  std::vector<int> rk((size_t)nb_ * nb_, 0);
  drank_.upload(rk.data(), rk.size());
  piv_.alloc(n);
  gperm_.alloc(n);
  SB200_CUDA(cudaStreamSynchronize(0));
}

void BLREngine::run(bool do_factor) {
  cudaStream_t st = 0;
  const int nb = nb_, n = n_;
  const long long tstride = (long long)maxtile_ * maxtile_;
  const int ntmax = 2 * (nb - 1);
  DevBuf<double> scratch((size_t)std::max(ntmax, 1) * tstride), Rbuf((size_t)std::max(ntmax, 1) * tstride),
      UT((size_t)tstride);
  DevBuf<int> order((size_t)std::max(ntmax, 1) * maxtile_), ranks_step(std::max(ntmax, 1)),
      sel(std::max(ntmax, 1));
  SB200_CUDA(cudaMemsetAsync(Rbuf.p, 0, sizeof(double) * Rbuf.n, st));
  // all per-step descriptors are built up front: no host sync inside the loop
  std::vector<TileDesc> td;
  std::vector<IDTask> idt;
  std::vector<SolveTask> stL;
  std::vector<int> step_ptr(nb + 1, 0);
  auto add_tile = [&](int i, int j, int slot) {
    TileDesc t;
    t.i = i; t.j = j; t.ro = off_[i]; t.co = off_[j];
    t.m = off_[i + 1] - off_[i]; t.n = off_[j + 1] - off_[j];
    t.lr = lroff_[i + (size_t)j * nb]; t.rc = rcap_[i + (size_t)j * nb];
    t.adm = 1;
    if (!opts_.admissible.empty() && i < opts_.nadm && j < opts_.nadm)
      t.adm = opts_.admissible[i + (size_t)j * opts_.nadm] != 0;
    td.push_back(t);
    IDTask q;
    q.M = scratch.p + slot * tstride; q.R = Rbuf.p + slot * tstride;
    q.ns = t.m; q.nc = t.n; q.rcap = t.rc;
    q.order = order.p + (size_t)slot * maxtile_; q.rank = ranks_step.p + slot;
    q.E = nullptr; q.strict = 1;
    idt.push_back(q);
  };
  const int nsteps = do_factor ? nsteps_ : nb;   // partial factorization: the tiles of A11 only
  for (int i = 0; i < nsteps; i++) {
    int slot = 0;
    if (do_factor) {
      for (int j = i + 1; j < nb; j++) { add_tile(i, j, slot++); add_tile(j, i, slot++); }
    } else {
      for (int j = 0; j < nb; j++) if (j != i) add_tile(i, j, slot++);   // compress only: row i
    }
    step_ptr[i + 1] = int(td.size());
  }
  // solve tasks: left: U of (i,j) ; right: Vt of (j,i)
  for (size_t q = 0; q < td.size(); q++) {
    const TileDesc& t = td[q];
    SolveTask s;
    if (t.i < t.j) { s.B = lr_.p + t.lr; s.ldb = t.m; }                          // U_ij, m x rank
    else { s.B = lr_.p + t.lr + (long long)t.m * t.rc; s.ldb = t.n; }            // Vt_ji, n x rank
    s.ncols = 0;
    stL.push_back(s);
  }
  // the same tiles kept dense (when they turn out incompressible): operands inside A
  std::vector<SolveTask> stD;
  for (const TileDesc& t : td) {
    SolveTask s;
    s.B = A_.p + t.ro + (size_t)t.co * n; s.ldb = n; s.ncols = 0;
    stD.push_back(s);
  }
  DevBuf<SolveTask> dstD; dstD.upload(stD.data(), stD.size(), st);
  std::vector<double> hid((size_t)maxtile_ * maxtile_, 0.);
  for (int q = 0; q < maxtile_; q++) hid[q + (size_t)q * maxtile_] = 1.;
  DevBuf<double> ident; ident.upload(hid.data(), hid.size(), st);
  DevBuf<TileDesc> dtd; dtd.upload(td.data(), td.size(), st);
  DevBuf<IDTask> didt; didt.upload(idt.data(), idt.size(), st);
  DevBuf<SolveTask> dst; dst.upload(stL.data(), stL.size(), st);
  const size_t cpqr_smem = (sizeof(double) + sizeof(int)) * (size_t)maxtile_ + 16;
  set_smem(id_cpqr_kernel, cpqr_smem);
  const int ldp = smem_ld(maxtile_);
  const int KC = maxtile_ <= 256 ? 32 : 16;
  const bool ll = do_factor && opts_.factor_algorithm == 1;
  auto schur_launch = [&](int pstep, int mode, int nblocks) {
    const size_t smem = sizeof(double) * ((size_t)3 * ldp * KC + (size_t)(KC + 4) * KC);
    if (KC == 32) { set_smem(blr_schur_kernel<32>, smem);
      blr_schur_kernel<32><<<nblocks, kThreads, smem, st>>>(A_.p, n, doff_.p, nb, pstep, lr_.p, dlroff_.p, drcap_.p, drank_.p, ldp, ident.p, maxtile_, mode);
    } else { set_smem(blr_schur_kernel<16>, smem);
      blr_schur_kernel<16><<<nblocks, kThreads, smem, st>>>(A_.p, n, doff_.p, nb, pstep, lr_.p, dlroff_.p, drcap_.p, drank_.p, ldp, ident.p, maxtile_, mode);
    }
    launches_++;
  };
  for (int i = 0; i < nsteps; i++) {
    const int m = off_[i + 1] - off_[i];
    const int cnt = step_ptr[i + 1] - step_ptr[i];
    double* Aii = A_.p + off_[i] + (size_t)off_[i] * n;
    // left-looking schedule: block row / column i receive the updates of all earlier steps now
    if (ll && i > 0) schur_launch(i, 1, 2 * (nb - i) - 1);
    if (do_factor) {
      {
        const int lp = smem_ld(maxtile_);
        if (maxtile_ <= 256) {
          const size_t smem = sizeof(double) * ((size_t)lp * 32 + (size_t)36 * maxtile_);
          set_smem(blr_getrf_kernel<32>, smem);
          blr_getrf_kernel<32><<<1, kThreads, smem, st>>>(Aii, n, m, piv_.p + off_[i], gperm_.p + off_[i],
                                                          opts_.pivot_threshold, lp);
        } else {
          const size_t smem = sizeof(double) * ((size_t)lp * 16 + (size_t)20 * maxtile_);
          set_smem(blr_getrf_kernel<16>, smem);
          blr_getrf_kernel<16><<<1, kThreads, smem, st>>>(Aii, n, m, piv_.p + off_[i], gperm_.p + off_[i],
                                                          opts_.pivot_threshold, lp);
        }
      }
      blr_upper_transpose_kernel<<<64, 256, 0, st>>>(Aii, n, m, UT.p);
      launches_ += 2;
    }
    if (!cnt) continue;
    const TileDesc* tds = dtd.p + step_ptr[i];
    // K11: compress the tiles of block row / column i
    blr_copy_tiles_kernel<<<dim3(cnt, 8), kThreads, 0, st>>>(A_.p, n, tds, scratch.p, tstride);
    id_cpqr_kernel<<<cnt, kCpqrThreads, cpqr_smem, st>>>(didt.p + step_ptr[i], opts_.rel_tol,
                                                         opts_.abs_tol, opts_.max_rank);
    blr_extract_lr_kernel<<<cnt, kThreads, 0, st>>>(tds, scratch.p, tstride, Rbuf.p, tstride, order.p,
                                                    maxtile_, ranks_step.p, lr_.p, drank_.p, nb);
    SB200_CUDA(cudaMemsetAsync(Rbuf.p, 0, sizeof(double) * (size_t)cnt * tstride, st));
    launches_ += 3;
    if (!do_factor) continue;
    // K12: U_ij <- L^{-1} P U_ij (even slots), Vt_ji <- U^{-T} Vt_ji (odd slots).
    // Both lists are interleaved in one descriptor array; each launch covers
    // all tiles and the kernel skips the ones of the other kind by ncols = 0.
    {
      const size_t smem = sizeof(double) * (size_t)kWarps * m;
      set_smem(blr_trsm_lower_kernel, smem);
      blr_select_kernel<<<(cnt + 127) / 128, 128, 0, st>>>(tds, cnt, 0, drank_.p, nb, sel.p);
      blr_trsm_lower_kernel<<<dim3(cnt, (maxtile_ / 2 + kWarps - 1) / kWarps), kThreads, smem, st>>>(
          Aii, n, m, 1, gperm_.p + off_[i], dst.p + step_ptr[i], sel.p);
      blr_select_kernel<<<(cnt + 127) / 128, 128, 0, st>>>(tds, cnt, 1, drank_.p, nb, sel.p);
      blr_trsm_lower_kernel<<<dim3(cnt, (maxtile_ / 2 + kWarps - 1) / kWarps), kThreads, smem, st>>>(
          UT.p, m, m, 0, nullptr, dst.p + step_ptr[i], sel.p);
      // dense off-diagonal tiles (DenseTile in the reference): A_ij <- L^{-1} P A_ij, A_ji <- A_ji U^{-1}
      blr_select_kernel<<<(cnt + 127) / 128, 128, 0, st>>>(tds, cnt, 2, drank_.p, nb, sel.p);
      blr_trsm_lower_kernel<<<dim3(cnt, (maxtile_ + kWarps - 1) / kWarps), kThreads, smem, st>>>(
          Aii, n, m, 1, gperm_.p + off_[i], dstD.p + step_ptr[i], sel.p);
      blr_trsm_right_upper_dense_kernel<<<cnt, kThreads, 0, st>>>(A_.p, n, Aii, m, tds, drank_.p, nb);
      launches_ += 7;
    }
    // K13: trailing update (right-looking schedule)
    const int nrem = nb - i - 1;
    if (nrem > 0 && !ll) {
      const size_t smem = sizeof(double) * ((size_t)3 * ldp * KC + (size_t)(KC + 4) * KC);
      if (KC == 32) { set_smem(blr_schur_kernel<32>, smem);
        blr_schur_kernel<32><<<nrem * nrem, kThreads, smem, st>>>(A_.p, n, doff_.p, nb, i, lr_.p, dlroff_.p, drcap_.p, drank_.p, ldp, ident.p, maxtile_, 0);
      } else { set_smem(blr_schur_kernel<16>, smem);
        blr_schur_kernel<16><<<nrem * nrem, kThreads, smem, st>>>(A_.p, n, doff_.p, nb, i, lr_.p, dlroff_.p, drcap_.p, drank_.p, ldp, ident.p, maxtile_, 0);
      }
      launches_++;
    }
  }
  if (do_factor && ll && nsteps < nb) schur_launch(nsteps, 2, (nb - nsteps) * (nb - nsteps));   // A22 of a front
  SB200_CUDA(cudaGetLastError());
  SB200_CUDA(cudaStreamSynchronize(st));
  hrank_.resize((size_t)nb * nb);
  SB200_CUDA(cudaMemcpy(hrank_.data(), drank_.p, sizeof(int) * hrank_.size(), cudaMemcpyDeviceToHost));
  // rank -1 = a tile that does not compress to rank <= min(m,n)/2: kept dense
  // inside A (DenseTile, reference BLRMatrix.cpp:563-570)
  dense_tiles_ = 0;
  for (int j = 0; j < nb; j++)
    for (int i = 0; i < nb; i++)
      if (i != j && hrank_[i + (size_t)j * nb] < 0) dense_tiles_++;
  factored_ = do_factor;
}


// ------------------------------------------------------------------ solve
// x <- A^{-1} x : laswp(piv), trsm(L, unit lower BLR), trsm(U, upper BLR)
// (BLRMatrix::solve, BLRMatrix.hpp:118-122; left-looking block substitution,
// BLRMatrix.cpp:1683-1707)
void BLREngine::build_solve_tasks(int s, double* dB, int ldB, cudaStream_t st) {
  const int nb = nb_;
  // task lists (k,j) for every block row, built once.  Forward: the tiles left
  // of the diagonal that belong to eliminated block columns (all of them unless
  // the factorization is partial); backward: the tiles right of the diagonal.
  if (!tasks_built_) {
    std::vector<GemvTask> tl;
    fwd_ptr_.assign(nb + 1, 0);
    for (int i = 0; i < nb; i++) {
      for (int j = 0; j < std::min(i, nsteps_); j++) tl.push_back({i, j});
      fwd_ptr_[i + 1] = int(tl.size());
    }
    bwd_ptr_.assign(nb + 1, 0);
    const int base = int(tl.size());
    for (int i = 0; i < nb; i++) {
      for (int j = i + 1; j < nb; j++) tl.push_back({i, j});
      bwd_ptr_[i + 1] = int(tl.size()) - base;
    }
    bwd_base_ = base;
    gtasks_.alloc(tl.size() * sizeof(GemvTask) / sizeof(int) + 2);
    SB200_CUDA(cudaMemcpyAsync(gtasks_.p, tl.data(), tl.size() * sizeof(GemvTask), cudaMemcpyHostToDevice, st));
    tasks_built_ = true;
  }
  DevBuf<SolveTask>& sv = solve_task_;
  if (!sv.p) {
    sv.alloc(nb);
    solve_ldb_ = -1;
  }
  if (solve_ldb_ != ldB || solve_ptr_ != dB || solve_s_ != s) {
    std::vector<SolveTask> t(nb);
    for (int i = 0; i < nb; i++) { t[i].B = dB + off_[i]; t[i].ldb = ldB; t[i].ncols = s; }
    SB200_CUDA(cudaMemcpyAsync(sv.p, t.data(), sizeof(SolveTask) * nb, cudaMemcpyHostToDevice, st));
    SB200_CUDA(cudaStreamSynchronize(st));
    solve_ldb_ = ldB; solve_ptr_ = dB; solve_s_ = s;
  }
}

// forward: x_i <- L_ii^{-1} P_i (x_i - sum_{j<i} T_ij x_j) for the eliminated
// block rows; the remaining rows (partial factorization) only receive the
// update  x_i -= sum_j F21_ij x_j   (trsmLNU_gemm, BLRMatrix.cpp:1552-1608)
void BLREngine::forward_rows(int s, double* dB, int ldB, cudaStream_t st) {
  const int nb = nb_, n = n_;
  const GemvTask* gt = reinterpret_cast<const GemvTask*>(gtasks_.p);
  const size_t gsm = sizeof(double) * (size_t)(maxtile_ / 2 + 8);
  for (int i = 0; i < nb; i++) {
    const int m = off_[i + 1] - off_[i], cnt = fwd_ptr_[i + 1] - fwd_ptr_[i];
    if (cnt) {
      blr_lr_gemv_kernel<<<dim3(cnt, s), kThreads, gsm, st>>>(gt + fwd_ptr_[i], doff_.p, nb, lr_.p, dlroff_.p,
                                                             drcap_.p, drank_.p, dB, ldB, dB, ldB, -1., A_.p, n);
      launches_++;
    }
    if (i >= nsteps_) continue;
    const size_t smem = sizeof(double) * (size_t)kWarps * m;
    set_smem(blr_trsm_lower_kernel, smem);
    blr_trsm_lower_kernel<<<dim3(1, (s + kWarps - 1) / kWarps), kThreads, smem, st>>>(
        A_.p + off_[i] + (size_t)off_[i] * n, n, m, 1, gperm_.p + off_[i], solve_task_.p + i, nullptr);
    launches_++;
  }
}

// backward: x_i <- U_ii^{-1} (x_i - sum_{j>i} T_ij x_j) over the eliminated
// block rows; with a partial factorization the tiles j of the trailing block
// are F12, i.e. gemm_trsmUNN (BLRMatrix.cpp:1610-1665)
void BLREngine::backward_rows(int s, double* dB, int ldB, cudaStream_t st) {
  const int nb = nb_, n = n_;
  const GemvTask* gt = reinterpret_cast<const GemvTask*>(gtasks_.p);
  const size_t gsm = sizeof(double) * (size_t)(maxtile_ / 2 + 8);
  for (int i = nsteps_ - 1; i >= 0; i--) {
    const int m = off_[i + 1] - off_[i], cnt = bwd_ptr_[i + 1] - bwd_ptr_[i];
    if (cnt) {
      blr_lr_gemv_kernel<<<dim3(cnt, s), kThreads, gsm, st>>>(gt + bwd_base_ + bwd_ptr_[i], doff_.p, nb, lr_.p,
                                                             dlroff_.p, drcap_.p, drank_.p, dB, ldB, dB, ldB, -1., A_.p, n);
      launches_++;
    }
    const size_t smem = sizeof(double) * (size_t)kWarps * m;
    set_smem(blr_trsm_upper_kernel, smem);
    blr_trsm_upper_kernel<<<(s + kWarps - 1) / kWarps, kThreads, smem, st>>>(
        A_.p + off_[i] + (size_t)off_[i] * n, n, m, dB + off_[i], ldB, s);
    launches_++;
  }
}

void BLREngine::solve(int s, double* dB, int ldB, cudaStream_t st) {
  if (!factored_) throw std::logic_error("BLR solve called on an unfactored matrix");
  if (partial()) throw std::logic_error("BLR solve: the factorization is partial (use the partial forward / backward solves around the Schur system)");
  if (s <= 0) return;
  build_solve_tasks(s, dB, ldB, st);
  forward_rows(s, dB, ldB, st);
  backward_rows(s, dB, ldB, st);
  SB200_CUDA(cudaGetLastError());
}

void BLREngine::partial_forward(int s, double* dB, int ldB, cudaStream_t st) {
  if (!factored_) throw std::logic_error("BLR solve called on an unfactored matrix");
  if (s <= 0) return;
  build_solve_tasks(s, dB, ldB, st);
  forward_rows(s, dB, ldB, st);
  SB200_CUDA(cudaGetLastError());
}

void BLREngine::partial_backward(int s, double* dB, int ldB, cudaStream_t st) {
  if (!factored_) throw std::logic_error("BLR solve called on an unfactored matrix");
  if (s <= 0) return;
  build_solve_tasks(s, dB, ldB, st);
  backward_rows(s, dB, ldB, st);
  SB200_CUDA(cudaGetLastError());
}

// y = op(A) x on the compressed (unfactored) matrix  (BLRMatrix::mult ->
// gemv, BLRMatrix.cpp:1742-1763)
void BLREngine::mult(char trans, int s, const double* dB, int ldB, double* dC, int ldC,
                     cudaStream_t st) {
  if (factored_) throw std::logic_error("BLR mult: the tiles hold LU factors (use solve)");
  const bool T = !(trans == 'N' || trans == 'n');
  const int nb = nb_;
  if (!mtasks_.p) {
    std::vector<GemvTask> tl;
    for (int i = 0; i < nb; i++)
      for (int j = 0; j < nb; j++) if (i != j) tl.push_back({i, j});
    nmt_ = int(tl.size());
    mtasks_.alloc(tl.size() * sizeof(GemvTask) / sizeof(int) + 2);
    SB200_CUDA(cudaMemcpy(mtasks_.p, tl.data(), tl.size() * sizeof(GemvTask), cudaMemcpyHostToDevice));
  }
  blr_zero_kernel<<<256, 256, 0, st>>>(dC, ldC, n_, s);
  blr_diag_gemv_kernel<<<dim3(nb, s), kThreads, 0, st>>>(A_.p, n_, doff_.p, dB, ldB, dC, ldC, T ? 1 : 0);
  if (nmt_) {
    const size_t gsm = sizeof(double) * (size_t)(maxtile_ / 2 + 8);
    blr_lr_gemv_kernel<<<dim3(nmt_, s), kThreads, gsm, st>>>(reinterpret_cast<const GemvTask*>(mtasks_.p),
                                                            doff_.p, nb, lr_.p, dlroff_.p, drcap_.p, drank_.p,
                                                            dB, ldB, dC, ldC, 1., A_.p, n_, T ? 1 : 0);
  }
  launches_ += 3;
  SB200_CUDA(cudaGetLastError());
}

int BLREngine::max_rank() const {
  // BLRMatrix::rank() = max of BLRTile::maximum_rank(), which is 0 for a
  // DenseTile (reference BLRMatrix.cpp:290-294, DenseTile.hpp:93)
  int r = 0;
  for (int v : hrank_) r = std::max(r, v);
  return r;
}

long long BLREngine::nonzeros() const {
  long long nnz = 0;
  for (int j = 0; j < nb_; j++)
    for (int i = 0; i < nb_; i++) {
      const long long m = off_[i + 1] - off_[i], n = off_[j + 1] - off_[j];
      const int r = hrank_[i + (size_t)j * nb_];
      nnz += (i == j || r < 0) ? m * n : (long long)r * (m + n);
    }
  return nnz;
}

}  // namespace sb200
