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
};

struct SolveTask { double* B; long long ldb; int ncols; };   // one column block of a batched trsm

class BLREngine {
 public:
  // do_factor = true : BLRMatrix::compress_and_factor(A, weak admissibility, RL)
  // do_factor = false: BLRMatrix::compress (all off-diagonal tiles low rank)
  BLREngine(int n, const double* A, int ldA, const BLROpts& o, bool do_factor,
            bool device_input = false);   // A: host pointer, or device pointer if device_input
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
  void run(bool do_factor);
  int n_ = 0, nb_ = 0, maxtile_ = 0, dense_tiles_ = 0;
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
