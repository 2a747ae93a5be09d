// strumpack_b200 -- device-resident HSS matrix: apply, ULV factor, ULV solve.
//
// Replaces (reference, CPU/OpenMP-task recursion, one BLAS call per block):
//   apply_HSS / apply_fwd / apply_bwd / applyT_*  src/HSS/HSSMatrix.apply.hpp:34-220
//   factor / factor_recursive                     src/HSS/HSSMatrix.factor.hpp:35-147
//   solve / solve_fwd / solve_bwd                 src/HSS/HSSMatrix.solve.hpp:35-238
// with per-height-class batched sm_100a kernels over a flat node table.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <vector>

#include "hss_tree.hpp"
#include "sb200_common.cuh"

namespace sb200 {

// One record per node, read by the CTA that owns the node.
struct DNode {
  int ch0, ch1, parent, leaf;
  int rows, cols, row_off, col_off;
  int u_rows, u_rank, v_rows, v_rank;
  long long D, Eu, Ev, B01, B10;  // offsets into the generator arena (-1: none)
  long long Pu, Pv;               // offsets into the permutation arena
  int w_off;                      // apply workspace: prefix of max(u_rank,v_rank)
  // ---- ULV ----
  int m, k;            // reduced-system size, eliminated unknowns (m - u_rank)
  int naug;            // k + v_rank + u_rank columns of the factor block
  long long F;         // factor block m x naug (ld = m) in the factor arena
  long long T;         // block-reflector T factors, nbq x k (ld = nbq)
  int nbq;             // panel width this node's QR runs with (its height class decides)
  int y_off, z_off, f_off, x_off;  // solve workspace prefixes (k, v_rank, u_rank, m)
};

// node lists of one sweep domain, grouped into height classes
struct NodeLists {
  std::vector<int> host, hptr;          // nodes sorted by height; class h = [hptr[h], hptr[h+1])
  std::vector<int> max_m, max_k;        // per class: launch configuration
  std::vector<long long> soff;          // per entry: Dfull scratch offset (inner nodes)
  long long smax = 0;
  DevBuf<int> list;
  DevBuf<long long> dsoff;
  int classes() const { return hptr.empty() ? 0 : int(hptr.size()) - 1; }
};

class HSSEngine {
 public:
  explicit HSSEngine(HSSHost&& host);
  ~HSSEngine();

  const HSSHost& host() const { return H_; }
  int rows() const { return H_.rows(); }
  int cols() const { return H_.cols(); }

  // C = op(H) B ; dB, dC device pointers (column-major)
  // beta != 0: C = op(H) B + beta C   (apply_HSS, reference HSSMatrix.cpp:419-435)
  void mult(char trans, int s, const double* dB, int ldB, double* dC, int ldC,
            cudaStream_t st, double beta = 0.);
  void factor(cudaStream_t st);
  void solve(int s, double* dB, int ldB, cudaStream_t st);
  // the two halves of solve (reference forward_solve / backward_solve): forward
  // reads B (a single-node tree also overwrites it), backward writes x into B
  void forward_solve(int s, double* dB, int ldB, cudaStream_t st);
  void backward_solve(int s, double* dB, int ldB, cudaStream_t st);
  void shift(double sigma, cudaStream_t st);
  // B (nI x nJ, device, ld ldB) = or += H(I, J); I, J host index lists
  // (HSSMatrix::extract / extract_add, reference HSSMatrix.extract.hpp:8-188)
  void extract(int nI, const int* I, int nJ, const int* J, double* dB, int ldB, bool add,
               cudaStream_t st);
  bool factored() const { return factored_; }

  // ---- subtree sharding over `nparts` GPUs: this rank owns the
  // subtree of cut node `part` at depth log2(nparts); the nparts-1 nodes above
  // the cut are replicated.  Each operation is split around the single small
  // exchange it needs: *_begin fills `send` (dist_sizes doubles), the caller
  // all-gathers send -> recv over NCCL, *_end consumes recv.
  void set_partition(int nparts, int part);
  int nparts() const { return nparts_; }
  void owned_range(int* lo, int* hi) const;
  void dist_sizes(int s, long long* out) const;   // apply, factor, solve (doubles per rank)
  void dist_mult_begin(char trans, int s, const double* dB, int ldB, double* send, cudaStream_t st);
  void dist_mult_end(char trans, int s, const double* dB, int ldB, double* dC, int ldC,
                     const double* recv, cudaStream_t st);
  void dist_factor_begin(double* send, cudaStream_t st);
  void dist_factor_end(const double* recv, cudaStream_t st);
  void dist_solve_begin(int s, double* dB, int ldB, double* send, cudaStream_t st);
  void dist_solve_end(int s, double* dB, int ldB, const double* recv, cudaStream_t st);

  long long factor_nonzeros() const { return fact_nnz_; }
  long long launches() const { return launches_; }
  // ULV factors to host (HSSMatrix::ULV()); null pointers: sizes only
  void export_ulv(double* factors, double* tfactors, long long* sizes);
  // optional live timing of the dominant kernel (leaf-class QR) with CUDA
  // events on the launching stream; ms of the last factor() call
  void set_profile(bool on);
  float qr_leaf_ms();
  // download the generator arena (after shift) for write_file / dense
  void sync_host_values();

 private:
  void build_tables();
  void ensure_apply_ws(int s);
  void ensure_solve_ws(int s);
  void make_lists(NodeLists& L, const std::vector<int>& nodes);
  int class_nb(int h, int max_m) const;   // QR panel width of a height class
  void run_up(const NodeLists& L, bool T, int s, const double* dB, int ldB, cudaStream_t st);
  void run_down(const NodeLists& L, bool T, int s, const double* dB, int ldB, double* dC,
                int ldC, bool leaves, cudaStream_t st, double beta = 0.);
  void factor_prepare();
  void factor_classes(const NodeLists& L, bool time_leaf, cudaStream_t st);
  void solve_fwd(const NodeLists& L, int s, double* dB, int ldB, cudaStream_t st);
  void solve_root(int s, double* dB, int ldB, cudaStream_t st);
  void solve_bwd(const NodeLists& L, int s, double* dB, int ldB, cudaStream_t st);

  HSSHost H_;
  std::vector<int> leaf_ids_;   // leaves in index order (extract)
  int max_depth_ = 0;
  std::vector<DNode> hn_;
  DevBuf<DNode> dn_;
  DevBuf<double> vals_;
  DevBuf<int32_t> perms_;
  NodeLists own_, top_;
  std::vector<int> cut_;
  int nparts_ = 1, part_ = 0;
  // apply workspace
  DevBuf<double> t1_, t2_;
  int ws_total_ = 0, apply_s_ = 0;
  // ULV
  DevBuf<double> fact_, tfac_, scratch_;
  DevBuf<int> rootpiv_;
  DevBuf<double> ysol_, zsol_, fsol_, xsol_;
  int solve_s_ = 0, fwd_s_ = 0;
  long long tot_k_ = 0, tot_rv_ = 0, tot_ru_ = 0, tot_m_ = 0;
  long long fact_nnz_ = 0;
  long long scratch_per_node_max_ = 0;
  bool factored_ = false;
  long long launches_ = 0;
  int nb_ = 32;
  int nsm_ = 148, qr_split_ = 0, qr_regpanel_ = 1, qr_skew_ = 0, qr_ll_ = 0, qr_variant_ = 1, qr_nowide_ = 0;   // switches; env SB200_QR_*
  int mm_min_ = 4;      // >= this many right-hand sides: GEMM-shaped (tensor pipe) apply kernels
  bool profile_ = false;
  cudaEvent_t ev_[2] = {nullptr, nullptr};
};

}  // namespace sb200
