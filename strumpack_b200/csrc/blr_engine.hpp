// strumpack_b200 -- device-resident BLR matrix (see blr_engine.cu).
#pragma once
#include <cuda_runtime.h>
#include <vector>

#include "sb200_common.cuh"

namespace sb200 {

struct BLROpts {            // reference BLROptions defaults, src/BLR/BLROptions.hpp:128-140
  double rel_tol = 1e-4, abs_tol = 1e-12;
  int leaf_size = 256;
  int max_rank = 5000;
  double pivot_threshold = -1.;
  // BLRFactorAlgorithm (BLROptions.hpp:65): 0 = RL (default), 1 = LL.  COMB / STAR
  // (LUAR accumulation with recompression) and COLWISE are mapped to RL.
  int factor_algorithm = 0;
  // admissibility of the tiles of the eliminated block (BLRMatrix adm_t,
  // BLRMatrix.hpp:78; strong admissibility marks neighbouring tiles
  // inadmissible = kept dense): nadm x nadm, column-major, nonzero = the tile
  // may be compressed.  Empty = weak admissibility (every off-diagonal tile).
  std::vector<int> admissible;
  int nadm = 0;
  // tile partition given by the caller (the reference passes tiles1 / tiles2 =
  // the leaf sizes of its separator / update cluster trees, FrontBLR.cpp:58-76);
  // empty: ClusterTree(n).refine(leaf_size)
  std::vector<int> tiles1, tiles2;
};

struct SolveTask { double* B; long long ldb; int ncols; };   // one column block of a batched trsm

class BLREngine {
 public:
  // do_factor = true : BLRMatrix::compress_and_factor(A, weak admissibility, RL)
  // do_factor = false: BLRMatrix::compress (all off-diagonal tiles low rank)
  BLREngine(int n, const double* A, int ldA, const BLROpts& o, bool do_factor,
            bool device_input = false);   // A: host pointer, or device pointer if device_input
  // BLRMatrix::construct_and_partial_factor(A11, A12, A21, A22, B11, B12, B21,
  // tiles1, tiles2, admissible, opts) (reference BLRMatrix.cpp:739-1037, RL):
  // the front [A11 A12; A21 A22] (n1 + n2 rows) is eliminated over the tiles of
  // A11 only; F11 = LU(A11) in BLR form, F12 = L^{-1} P A12 and F21 = A21 U^{-1}
  // compressed, and the trailing n2 x n2 block becomes the dense Schur
  // complement A22 - A21 A11^{-1} A12.  Tiles: ClusterTree(n1/n2).refine(leaf).
  BLREngine(int n1, int n2, const double* A11, int ld11, const double* A12, int ld12,
            const double* A21, int ld21, const double* A22, int ld22, const BLROpts& o,
            bool device_input = false);
  bool partial() const { return nsteps_ < nb_; }
  int sep_rows() const { return n1_; }
  int upd_rows() const { return n_ - n1_; }
  const double* schur() const { return A_.p + n1_ + (size_t)n1_ * n_; }   // device, n2 x n2, ld = rows()
  // forward: [b1; b2] <- [L11^{-1} P b1; b2 - F21 (L11^{-1} P b1)]  (laswp + trsmLNU_gemm,
  // BLRMatrix.cpp:1552-1608, FrontBLR.cpp:529-531); backward: b1 <- U11^{-1} (b1 - F12 b2)
  // (gemm_trsmUNN, BLRMatrix.cpp:1610-1665, FrontBLR.cpp:555)
  void partial_forward(int s, double* dB, int ldB, cudaStream_t st);
  void partial_backward(int s, double* dB, int ldB, cudaStream_t st);
  int rows() const { return n_; }
  int cols() const { return n_; }
  int tiles() const { return nb_; }
  int dense_tiles() const { return dense_tiles_; }   // off-diagonal tiles kept dense
  int max_rank() const;
  long long nonzeros() const;
  long long memory_bytes() const { return nonzeros() * (long long)sizeof(double); }
  bool factored() const { return factored_; }
  long long launches() const { return launches_; }
  void solve(int s, double* dB, int ldB, cudaStream_t st);
  void mult(char trans, int s, const double* dB, int ldB, double* dC, int ldC, cudaStream_t st);

 private:
  void setup(const std::vector<int>& tiles, bool do_factor);
  void run(bool do_factor);
  void build_solve_tasks(int s, double* dB, int ldB, cudaStream_t st);
  void forward_rows(int s, double* dB, int ldB, cudaStream_t st);
  void backward_rows(int s, double* dB, int ldB, cudaStream_t st);
  int n_ = 0, nb_ = 0, maxtile_ = 0, dense_tiles_ = 0;
  int n1_ = 0, nsteps_ = 0;   // eliminated rows / tile steps (= n_, nb_ unless partial)
  BLROpts opts_;
  std::vector<int> off_, rcap_, hrank_;
  std::vector<long long> lroff_;
  DevBuf<double> A_, lr_;
  DevBuf<int> doff_, drcap_, drank_, piv_, gperm_;
  DevBuf<long long> dlroff_;
  bool factored_ = false;
  long long launches_ = 0;
  // solve / mult task lists
  bool tasks_built_ = false;
  std::vector<int> fwd_ptr_, bwd_ptr_;
  int bwd_base_ = 0, nmt_ = 0;
  DevBuf<int> gtasks_, mtasks_;
  DevBuf<SolveTask> solve_task_;
  long long solve_ldb_ = -1;
  const double* solve_ptr_ = nullptr;
  int solve_s_ = 0;
};

}  // namespace sb200
