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
#include <functional>
#include <map>
#include <string>
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
  // The same three operations with the exchange done by the engine itself: the
  // local sweep, ONE ncclAllGather on `st` and the replicated top + the sweep back
  // down are queued back to back (and replayed as one CUDA graph).  dist_init joins
  // the communicator (unique id from nccl_unique_id(), shared by the caller) and
  // sets the partition.
  static void nccl_unique_id(char* out128);
  void dist_init(int nparts, int part, const char* unique_id128);
  void dist_mult(char trans, int s, const double* dB, int ldB, double* dC, int ldC, cudaStream_t st);
  void dist_factor(cudaStream_t st);
  void dist_solve(int s, double* dB, int ldB, cudaStream_t st);

  // ---- Schur complement of the (0,0) block of H = [H00 H01; H10 H11] (root's
  // children), what the reference's HSS fronts use (reference
  // HSSMatrix.factor.hpp:44-50 partial_factor, HSSMatrix.Schur.hpp:35-215,
  // caller src/sparse/fronts/FrontHSS.cpp:385-410,150-222,440-500):
  //   S = H11 - Theta Vhat^H Phi^H,  Theta = U1big B10 (rows1 x rv0),
  //   DUB01 = D0^{-1} U0 B01 (m0 x rv1),  Phi = V1big DUB01^H (cols1 x m0),
  //   Vhat = Vh of child 0 (m0 x rv0), m0 = reduced size of child 0.
  // All matrix arguments are DEVICE pointers, column-major.
  void partial_factor(cudaStream_t st);
  bool partially_factored() const { return pf_ok_; }
  // out[0..6] = rows(ch1), cols(ch1), v_rank(ch0), m0, v_rank(ch1), u_rank(ch1), rows(ch0)
  void schur_sizes(int* out) const;
  void schur_update(double* dTheta, int ldT, double* dDUB01, int ldD, double* dPhi, int ldP,
                    cudaStream_t st);
  const double* vhat() const { return vhat_.p; }   // m0 x v_rank(ch0), ld = m0
  // Sr = S R, Sc = S^H R for R (rows1 x c)     (Schur_product_direct)
  void schur_product_direct(const double* dTheta, int ldT, const double* dDUB01, int ldD,
                            const double* dPhi, int ldP, int c, const double* dR, int ldR,
                            double* dSr, int ldSr, double* dSc, int ldSc, cudaStream_t st);
  // Sr = Sr1 - U1big (B10 V0big^H R0 + B10 Vhat^H DUB01 V1big^H R1),
  // Sc = Sc1 - V1big (B01^H U0big^H R0 + (B10 Vhat^H DUB01)^H U1big^H R1)
  //                                             (Schur_product_indirect)
  void schur_product_indirect(const double* dDUB01, int ldD, int c, const double* dR0, int ldR0,
                              const double* dR1, int ldR1, const double* dSr1, int ldSr1,
                              const double* dSc1, int ldSc1, double* dSr, int ldSr,
                              double* dSc, int ldSc, cudaStream_t st);
  // child(0)->forward_solve(w, b, partial = true) / child(0)->backward_solve(w, x)
  // (reference HSSMatrix.solve.hpp:52-66,133-152; FrontHSS.cpp:452-462,487-495):
  // forward eliminates b0 (rows0 x s, overwritten) in the subtree of child 0,
  // solves with D0 and returns reduced_rhs = Vhat^H x + V0^H [z0; z1] (rv0 x s);
  // the reduced solution x (m0 x s, kept inside the engine) can be updated in
  // place (partial_x) before backward expands it into x0 (rows0 x s).
  void partial_forward_solve(int s, double* dB0, int ldB, double* dRed, int ldRed, cudaStream_t st);
  double* partial_x(int s);   // device, m0 x s, ld = m0
  void partial_backward_solve(int s, double* dX0, int ldX, cudaStream_t st);

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
  void factor_prepare(bool whole = true);
  // lu_node: the node that is LU-factored instead of eliminated (the root; child
  // 0 for partial_factor, whose LU goes to lu_dst / lu_piv)
  void factor_classes(const NodeLists& L, bool time_leaf, cudaStream_t st, int lu_node = 0,
                      double* lu_dst = nullptr, int* lu_piv = nullptr);
  void schur_lists();
  // nclass >= 0: only the lowest nclass height classes of the list
  void solve_fwd(const NodeLists& L, int s, double* dB, int ldB, cudaStream_t st, int nclass = -1);
  void solve_root(int s, double* dB, int ldB, cudaStream_t st, int node = 0,
                  const double* lu = nullptr, const int* piv = nullptr);
  void solve_bwd(const NodeLists& L, int s, double* dB, int ldB, cudaStream_t st, int nclass = -1);

  HSSHost H_;
  std::vector<int> leaf_ids_;   // leaves in index order (extract)
  std::vector<int> class_gmm_;  // per height class: largest block of the whole tree
  int max_depth_ = 0;
  std::vector<DNode> hn_;
  DevBuf<DNode> dn_;
  DevBuf<double> vals_;
  DevBuf<int32_t> perms_;
  NodeLists own_, top_;
  std::vector<int> cut_;
  void* nccl_comm_ = nullptr;          // ncclComm_t of the sharded matrix (dist_init)
  DevBuf<double> xsend_[3], xrecv_[3]; // exchange buffers of dist_mult / dist_factor / dist_solve
  void all_gather(int which, long long count, cudaStream_t st);
  void dist_close();
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
  long long fact_nnz_ = 0, fact_len_ = 0;   // factor blocks + T factors; factor blocks alone
  long long scratch_per_node_max_ = 0;
  bool factored_ = false;
  long long launches_ = 0;
  // CUDA graphs of the per-class launch sequences (apply / factor / solve and the
  // begin / end halves of their sharded versions): the sweeps above the leaves
  // are ~90 latency-bound launches per step; replaying a captured graph removes
  // the per-launch gaps on the device.  Keyed by operation + operand addresses;
  // dropped whenever an arena or a node list changes.  SB200_GRAPH=0 disables.
  struct GraphEntry { cudaGraphExec_t exec = nullptr; long long launches = 0; };
  std::map<std::string, GraphEntry> graphs_;
  std::map<std::string, int> seen_;    // a sequence is captured the second time it is asked for
  int use_graph_ = 1;
  void drop_graphs();
  void run_graphed(const std::string& key, cudaStream_t st, const std::function<void()>& body);
  // Schur / partial factorization state
  NodeLists sub0_, sub1_;        // subtrees of the root's children (cut nodes included)
  bool sub_ok_ = false, pf_ok_ = false;
  int pfwd_s_ = 0;
  DevBuf<double> pf_lu_, vhat_, zeros_, stmp_[4];
  DevBuf<int> pf_piv_;
  int nb_ = 32;
  DevBuf<unsigned char> tmaps_;          // one CUtensorMap (128 B) per node: ulv_qr3_kernel's TMA feed
  const double* tmaps_for_ = nullptr;    // the factor arena those descriptors point into
  int qr3_ = 0;        // SB200_QR3=1: ulv_qr3.cuh (left-looking, TMA-fed, warp-specialised) for classes with m <= 256;
                       // measured slower than the right-looking kernel, kept as an option
  int nsm_ = 148, qr_regpanel_ = 1, qr_variant_ = 1, qr_nowide_ = 0;   // switches; env SB200_QR_*
  int elim_variant_ = 0;   // ulv_eliminate_kernel: 0 = 32-column tiles, 1 = 16-column tiles x 4 CTAs/SM, 2 = 16-column tiles prefetched (SB200_ELIM_VARIANT)
  int solve_pipe_ = 3;  // bit 0: ulv_bwd_pipe_kernel, bit 1: ulv_fwd_pipe_kernel (SB200_SOLVE_PIPE; 0 = the non-streamed kernels)
  int mm_min_ = 4;      // >= this many right-hand sides: GEMM-shaped (tensor pipe) apply kernels
  bool profile_ = false;
  cudaEvent_t ev_[2] = {nullptr, nullptr};
};

// live-measured fp64 tensor-pipe (mma.sync m8n8k4) peak of the current device, TFLOP/s
double measure_fp64_dmma_peak_tflops();

// test / microbenchmark hook for the leaf QR kernels (see hss_engine.cu)
void debug_qr_batch(int m, int k, int naug, int count, const double* hA, double* hOut, double* hT, int variant,
                    int reps, float* ms);

}  // namespace sb200
