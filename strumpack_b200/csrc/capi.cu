// strumpack_b200 -- C ABI (include/sb200_structured.h).
// Mirrors reference src/structured/StructuredMatrixC.cpp:39-119: opaque handle
// owning the matrix, try/catch around every call, "Operation failed: ..." on
// stderr and return code 1 on error.
#include "../../include/sb200_structured.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <vector>

#include "blr_engine.hpp"
#include "hss_compress.hpp"
#include "hss_engine.hpp"
#include "hss_tree.hpp"

using namespace sb200;

namespace {

// One handle = one matrix + its own CUDA stream + its own staging buffers.  The
// host-pointer entry points take the handle's lock, queue copies and kernels on
// the handle's stream and wait for that stream only: calls on DIFFERENT handles
// from different threads overlap on the device, calls on the SAME handle are
// serialised (re-entrant per object, thread-safe across objects).
// The stream is a blocking one, so the few legacy-default-stream operations of
// the set-up paths (allocations, memsets, table uploads) still order against it.
struct Mat {
  SP_STRUCTURED_TYPE type = SP_TYPE_HSS;
  std::unique_ptr<HSSEngine> hss;
  std::unique_ptr<BLREngine> blr;
  cudaStream_t st = nullptr;
  std::mutex mu;
  Mat() { if (cudaStreamCreate(&st) != cudaSuccess) st = nullptr; }
  ~Mat() { if (st) cudaStreamDestroy(st); }
  Mat(const Mat&) = delete;
  Mat& operator=(const Mat&) = delete;
  // staging buffers for the host-pointer entry points
  DevBuf<double> dB, dC;
  // device copies of Theta / DUB01 / Phi of the last Schur_update and staging
  // for the Schur products
  DevBuf<double> dTheta, dDUB01, dPhi, dS[6];
};

void require_gpu() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    throw std::runtime_error(
        "no CUDA device: strumpack_b200 has no CPU fallback (sm_100a only)");
}

template <typename F> int guarded(F&& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    std::cerr << "Operation failed: " << e.what() << std::endl;
    return 1;
  } catch (...) {
    std::cerr << "Operation failed: unknown exception" << std::endl;
    return 1;
  }
}

Mat* M(const CSPStructMat S) {
  if (!S) throw std::invalid_argument("null CSPStructMat");
  return static_cast<Mat*>(S);
}

HSSEngine& hss(const CSPStructMat S) {
  Mat* m = M(S);
  if (!m->hss) throw std::logic_error("operation not supported for this type");
  return *m->hss;
}

// copy a host column-major block to a packed device buffer and back
void h2d(DevBuf<double>& d, const double* h, int rows, int cols, int ld,
         cudaStream_t st) {
  d.ensure((size_t)rows * cols);
  SB200_CUDA(cudaMemcpy2DAsync(d.p, sizeof(double) * rows, h, sizeof(double) * ld,
                               sizeof(double) * rows, cols,
                               cudaMemcpyHostToDevice, st));
}
void d2h(double* h, const DevBuf<double>& d, int rows, int cols, int ld,
         cudaStream_t st) {
  SB200_CUDA(cudaMemcpy2DAsync(h, sizeof(double) * ld, d.p, sizeof(double) * rows,
                               sizeof(double) * rows, cols,
                               cudaMemcpyDeviceToHost, st));
}

// float <-> double at the SP_s_* boundary
std::vector<double> widen(const float* A, int rows, int cols, int ld) {
  std::vector<double> D((size_t)rows * cols);
  for (int j = 0; j < cols; j++)
    for (int i = 0; i < rows; i++) D[i + (size_t)j * rows] = A[i + (size_t)j * ld];
  return D;
}
void narrow(const std::vector<double>& D, float* A, int rows, int cols, int ld) {
  for (int j = 0; j < cols; j++)
    for (int i = 0; i < rows; i++) A[i + (size_t)j * ld] = (float)D[i + (size_t)j * rows];
}

// y[i, 0..3] = sum_j A[i, j] r[j, 0..3]; A packed column-major n x n on the device
__global__ void dense_matvec4_kernel(const double* __restrict__ A, int n,
                                     const double* __restrict__ R, double* __restrict__ Y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double y0 = 0., y1 = 0., y2 = 0., y3 = 0.;
  for (int j = 0; j < n; j++) {
    const double a = A[i + (size_t)j * n];
    y0 += a * R[j]; y1 += a * R[j + n]; y2 += a * R[j + 2 * (size_t)n]; y3 += a * R[j + 3 * (size_t)n];
  }
  Y[i] = y0; Y[i + n] = y1; Y[i + 2 * (size_t)n] = y2; Y[i + 3 * (size_t)n] = y3;
}

// || A R - H R ||_F / || A R ||_F for 4 random vectors, all on the device
double dense_check(HSSEngine& E, int n, const double* dA) {
  std::vector<double> R((size_t)n * 4), YA((size_t)n * 4), YH((size_t)n * 4);
  unsigned long long x = 0x9E3779B97F4A7C15ull;
  for (auto& v : R) {   // xorshift64*, uniform in [-1, 1)
    x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
    v = (double)((x * 0x2545F4914F6CDD1Dull) >> 11) / 4503599627370496.0 - 1.0;
  }
  DevBuf<double> dR, dYA((size_t)n * 4), dYH((size_t)n * 4);
  dR.upload(R.data(), R.size());
  dense_matvec4_kernel<<<(n + 127) / 128, 128>>>(dA, n, dR.p, dYA.p);
  SB200_CUDA(cudaGetLastError());
  E.mult('N', 4, dR.p, n, dYH.p, n, 0);
  SB200_CUDA(cudaMemcpy(YA.data(), dYA.p, sizeof(double) * YA.size(), cudaMemcpyDeviceToHost));
  SB200_CUDA(cudaMemcpy(YH.data(), dYH.p, sizeof(double) * YH.size(), cudaMemcpyDeviceToHost));
  double num = 0., den = 0.;
  for (size_t q = 0; q < YA.size(); q++) { num += (YA[q] - YH[q]) * (YA[q] - YH[q]); den += YA[q] * YA[q]; }
  return den > 0. ? std::sqrt(num / den) : std::sqrt(num);
}

// HSS from a dense matrix.  Up to n = 8192 every node is compressed against its
// whole complement (exact).  Above that the construction samples complement
// columns by index distance, which misses columns of a matrix that does not
// decay with |i-j| (round-1 advisor finding): the result is therefore checked a
// posteriori with random products against A (the role the reference's adaptive
// randomized sampling plays, HSSMatrix.compress_stable.hpp) and the sample is
// enlarged, then replaced by the whole complement, until the tolerance is met.
std::unique_ptr<HSSEngine> hss_from_dense_checked(int rows, int cols, const double* A, int ldA,
                                                  CompressOptions co, const GivenTree* tree = nullptr) {
  for (int attempt = 0;; attempt++) {
    DevBuf<double> dA;
    auto E = std::make_unique<HSSEngine>(compress_dense(rows, cols, A, ldA, co, &dA, tree));
    const bool sampled = co.full_complement == 0 || (co.full_complement < 0 && rows > 8192);
    if (!sampled || !dA.p) return E;
    const double err = dense_check(*E, rows, dA.p);
    const double bound = 20. * co.rel_tol;
    if (co.verbose)
      std::printf("# sb200 compress: sampled construction, a-posteriori ||AR - HR||/||AR|| = %.3e (bound %.1e)\n", err, bound);
    if (err <= bound) return E;
    if (attempt == 0) { co.sample_near *= 4; co.sample_far *= 4; }
    else if (attempt == 1 && rows <= 32768) co.full_complement = 1;
    else {
      std::cerr << "sb200 warning: HSS construction from sampled columns of a dense matrix reached relative error "
                << err << " > " << bound << " (rel_tol " << co.rel_tol
                << "); the matrix does not decay with |i-j|: reorder it or pass coordinates" << std::endl;
      return E;
    }
  }
}

}  // namespace

extern "C" {

const char* SB200_version(void) { return "strumpack_b200 0.1 sm_100a"; }

void SP_d_struct_default_options(CSPOptions* o) {
  // StructuredOptions defaults, reference StructuredOptions.hpp:106-162
  o->type = SP_TYPE_BLR;
  o->rel_tol = 1e-4;
  o->abs_tol = 1e-10;
  o->leaf_size = 128;
  o->max_rank = 5000;
  o->verbose = 1;   // StructuredOptions::verbose_ = true (StructuredOptions.hpp:160)
}

void SP_d_struct_destroy(CSPStructMat* S) {
  if (!S || !*S) return;
  delete static_cast<Mat*>(*S);
  *S = nullptr;
}

// queries never throw across the C boundary: a handle that holds neither kind
// of matrix (or a null one) answers 0
int SP_d_struct_rows(const CSPStructMat S) {
  if (!S) return 0;
  const Mat* m = static_cast<const Mat*>(S);
  return m->blr ? m->blr->rows() : m->hss ? m->hss->rows() : 0;
}
int SP_d_struct_cols(const CSPStructMat S) {
  if (!S) return 0;
  const Mat* m = static_cast<const Mat*>(S);
  return m->blr ? m->blr->cols() : m->hss ? m->hss->cols() : 0;
}
long long int SP_d_struct_memory(const CSPStructMat S) {
  if (!S) return 0;
  const Mat* m = static_cast<const Mat*>(S);
  return m->blr ? m->blr->memory_bytes() : m->hss ? m->hss->host().memory_bytes() : 0;
}
long long int SP_d_struct_nonzeros(const CSPStructMat S) {
  if (!S) return 0;
  const Mat* m = static_cast<const Mat*>(S);
  return m->blr ? m->blr->nonzeros() : m->hss ? m->hss->host().nonzeros() : 0;
}
int SP_d_struct_rank(const CSPStructMat S) {
  if (!S) return 0;
  const Mat* m = static_cast<const Mat*>(S);
  return m->blr ? m->blr->max_rank() : m->hss ? m->hss->host().max_rank() : 0;
}

int SP_d_struct_from_dense(CSPStructMat* S, int rows, int cols, const double* A,
                           int ldA, const CSPOptions* opts) {
  return guarded([&] {
    require_gpu();
    auto m = std::make_unique<Mat>();
    m->type = opts->type;
    if (opts->type == SP_TYPE_BLR) {
      // construct_from_dense, Type::BLR: compress only (StructuredMatrix.cpp:78-98)
      if (rows != cols) throw std::invalid_argument("BLR: only square matrices are supported");
      BLROpts bo;
      bo.rel_tol = opts->rel_tol; bo.abs_tol = opts->abs_tol;
      bo.leaf_size = opts->leaf_size; bo.max_rank = opts->max_rank;
      m->blr = std::make_unique<BLREngine>(rows, A, ldA, bo, false);
      *S = m.release();
      return;
    }
    if (opts->type != SP_TYPE_HSS)
      throw std::invalid_argument("structured type not supported (HSS and BLR only)");
    CompressOptions co;
    co.rel_tol = opts->rel_tol; co.abs_tol = opts->abs_tol;
    co.leaf_size = opts->leaf_size; co.max_rank = opts->max_rank;
    co.verbose = opts->verbose;
    m->hss = hss_from_dense_checked(rows, cols, A, ldA, co);
    *S = m.release();
  });
}

int SP_d_struct_from_elements(CSPStructMat* S, int rows, int cols,
                              double A(int i, int j), const CSPOptions* opts) {
  return guarded([&] {
    require_gpu();
    auto m = std::make_unique<Mat>();
    m->type = opts->type;
    if (opts->type == SP_TYPE_BLR) {
      // construct_from_elements, Type::BLR (StructuredMatrix.cpp:230-252): every
      // tile is filled through the callback, then compressed; the callback is a
      // host function, so the n^2 evaluations happen here and the tiles are
      // compressed on the device
      if (rows != cols) throw std::invalid_argument("BLR: only square matrices are supported");
      std::vector<double> Ad((size_t)rows * cols);
      for (int j = 0; j < cols; j++)
        for (int i = 0; i < rows; i++) Ad[i + (size_t)j * rows] = A(i, j);
      BLROpts bo;
      bo.rel_tol = opts->rel_tol; bo.abs_tol = opts->abs_tol;
      bo.leaf_size = opts->leaf_size; bo.max_rank = opts->max_rank;
      m->blr = std::make_unique<BLREngine>(rows, Ad.data(), rows, bo, false);
      *S = m.release();
      return;
    }
    if (opts->type != SP_TYPE_HSS)
      throw std::invalid_argument("structured type not supported (HSS and BLR only)");
    CompressOptions co;
    co.rel_tol = opts->rel_tol; co.abs_tol = opts->abs_tol;
    co.leaf_size = opts->leaf_size; co.max_rank = opts->max_rank;
    co.verbose = opts->verbose;
    if (rows != cols) throw std::invalid_argument("HSS: only square matrices are supported");
    if (rows > 16384)
      throw std::invalid_argument("from_elements: scalar host callbacks are limited to n <= 16384; use "
                                  "SB200_d_hss_from_element_blocks or SB200_d_hss_from_kernel for larger matrices");
    std::vector<double> Ad((size_t)rows * cols);
    for (int j = 0; j < cols; j++)
      for (int i = 0; i < rows; i++) Ad[i + (size_t)j * rows] = A(i, j);
    m->hss = hss_from_dense_checked(rows, cols, Ad.data(), rows, co);
    *S = m.release();
  });
}

int SB200_d_hss_from_element_blocks(CSPStructMat* S, int n, SB200ElemBlockFn elem, void* user,
                                    const CSPOptions* opts) {
  return guarded([&] {
    require_gpu();
    auto m = std::make_unique<Mat>();
    m->type = SP_TYPE_HSS;
    CompressOptions co;
    co.rel_tol = opts->rel_tol; co.abs_tol = opts->abs_tol;
    co.leaf_size = opts->leaf_size; co.max_rank = opts->max_rank;
    co.verbose = opts->verbose;
    m->hss = std::make_unique<HSSEngine>(compress_element_blocks(n, elem, user, co));
    *S = m.release();
  });
}

int SB200_d_hss_from_element_blocks_ex(CSPStructMat* S, int n, SB200ElemBlockFn elem, void* user,
                                       const CSPOptions* opts, int tree_nodes, const int* tree_sizes,
                                       const int* tree_nchild, int d, const double* coords) {
  return guarded([&] {
    require_gpu();
    auto m = std::make_unique<Mat>();
    m->type = SP_TYPE_HSS;
    CompressOptions co;
    co.rel_tol = opts->rel_tol; co.abs_tol = opts->abs_tol;
    co.leaf_size = opts->leaf_size; co.max_rank = opts->max_rank;
    co.verbose = opts->verbose;
    GivenTree t{tree_nodes, tree_sizes, tree_nchild};
    m->hss = std::make_unique<HSSEngine>(compress_element_blocks(n, elem, user, co, tree_nodes > 0 ? &t : nullptr, d, coords));
    *S = m.release();
  });
}

int SB200_d_hss_from_dense_tree(CSPStructMat* S, int n, const double* A, int ldA, const CSPOptions* opts,
                                int tree_nodes, const int* tree_sizes, const int* tree_nchild) {
  return guarded([&] {
    require_gpu();
    auto m = std::make_unique<Mat>();
    m->type = SP_TYPE_HSS;
    CompressOptions co;
    co.rel_tol = opts->rel_tol; co.abs_tol = opts->abs_tol;
    co.leaf_size = opts->leaf_size; co.max_rank = opts->max_rank;
    co.verbose = opts->verbose;
    GivenTree t{tree_nodes, tree_sizes, tree_nchild};
    m->hss = hss_from_dense_checked(n, n, A, ldA, co, tree_nodes > 0 ? &t : nullptr);
    *S = m.release();
  });
}

/* nodes x 10 (pre-order): parent, child 0, child 1, rows, cols, row offset, column offset, U rank, V rank,
 * height.  out == NULL: only the number of nodes is returned (-1 on error). */
int SB200_d_hss_node_table(const CSPStructMat S, long long int* out) {
  int count = -1;
  guarded([&] {
    const auto& nodes = hss(S).host().nodes;
    count = (int)nodes.size();
    if (!out) return;
    for (size_t i = 0; i < nodes.size(); i++) {
      const auto& n = nodes[i];
      long long* r = out + 10 * i;
      r[0] = n.parent; r[1] = n.ch0; r[2] = n.ch1; r[3] = n.rows; r[4] = n.cols;
      r[5] = n.row_off; r[6] = n.col_off; r[7] = n.u_rank; r[8] = n.v_rank; r[9] = n.height;
    }
  });
  return count;
}

int SB200_d_hss_from_kernel(CSPStructMat* S, int n, int d, double* pts,
                            int kernel_type, double h, double lambda,
                            const CSPOptions* opts, int* perm) {
  return SB200_d_hss_from_kernel_ex(S, n, d, pts, kernel_type, h, lambda, opts, perm, 2);
}

int SB200_d_hss_from_kernel_ex(CSPStructMat* S, int n, int d, double* pts,
                               int kernel_type, double h, double lambda,
                               const CSPOptions* opts, int* perm, int clustering) {
  return guarded([&] {
    require_gpu();
    auto m = std::make_unique<Mat>();
    m->type = SP_TYPE_HSS;
    CompressOptions co;
    co.rel_tol = opts->rel_tol; co.abs_tol = opts->abs_tol;
    co.leaf_size = opts->leaf_size; co.max_rank = opts->max_rank;
    co.verbose = opts->verbose;
    m->hss = std::make_unique<HSSEngine>(
        compress_kernel(n, d, pts, kernel_type, h, lambda, co, perm, clustering));
    *S = m.release();
  });
}

// BLRFactorAlgorithm of the reference (BLROptions.hpp:65): COLWISE 0, RL 1, LL 2,
// COMB 3, STAR 4.  The engine has the right-looking (the reference's default)
// and the left-looking schedule.  COLWISE, COMB and STAR (LUAR accumulation with
// recompression, BLRMatrix.cpp:216-235,1039-1187) are not built: asking for them
// is an error, not a silent substitution.
static BLROpts blr_opts(const CSPOptions* opts, const SB200BLRParams* p) {
  BLROpts bo;
  bo.rel_tol = opts->rel_tol; bo.abs_tol = opts->abs_tol;
  bo.leaf_size = opts->leaf_size; bo.max_rank = opts->max_rank;
  if (p) {
    if (p->factor_algorithm < 0 || p->factor_algorithm > 4)
      throw std::invalid_argument("unknown BLR factor algorithm");
    if (p->factor_algorithm != 1 && p->factor_algorithm != 2)
      throw std::invalid_argument(
          std::string("BLR factor algorithm ") +
          (p->factor_algorithm == 0 ? "COLWISE" : p->factor_algorithm == 3 ? "COMB" : "STAR") +
          " is not implemented by this engine (RL, the reference's default, and LL are)");
    bo.pivot_threshold = p->pivot_threshold;
    bo.factor_algorithm = p->factor_algorithm == 2 ? 1 : 0;
    if (p->tiles1 && p->n_tiles1 > 0) bo.tiles1.assign(p->tiles1, p->tiles1 + p->n_tiles1);
    if (p->tiles2 && p->n_tiles2 > 0) bo.tiles2.assign(p->tiles2, p->tiles2 + p->n_tiles2);
    if (p->admissible) {
      if (p->n_admissible <= 0) throw std::invalid_argument("BLR admissibility matrix without a size");
      bo.nadm = p->n_admissible;
      bo.admissible.assign(p->admissible, p->admissible + (size_t)bo.nadm * bo.nadm);
    }
  }
  return bo;
}

int SB200_d_blr_compress_and_factor_ex(CSPStructMat* S, int n, const double* A, int ldA,
                                       const CSPOptions* opts, const SB200BLRParams* params) {
  return guarded([&] {
    require_gpu();
    auto m = std::make_unique<Mat>();
    m->type = SP_TYPE_BLR;
    m->blr = std::make_unique<BLREngine>(n, A, ldA, blr_opts(opts, params), true);
    *S = m.release();
  });
}

int SB200_d_blr_compress_and_factor(CSPStructMat* S, int n, const double* A, int ldA,
                                    const CSPOptions* opts, double pivot_threshold) {
  SB200BLRParams p{pivot_threshold, 1, nullptr, 0, nullptr, 0, nullptr, 0};
  return SB200_d_blr_compress_and_factor_ex(S, n, A, ldA, opts, &p);
}

int SB200_d_blr_compress_and_factor_device(CSPStructMat* S, int n, const double* dA, int ldA,
                                           const CSPOptions* opts, double pivot_threshold) {
  return guarded([&] {
    require_gpu();
    auto m = std::make_unique<Mat>();
    m->type = SP_TYPE_BLR;
    BLROpts bo;
    bo.rel_tol = opts->rel_tol; bo.abs_tol = opts->abs_tol;
    bo.leaf_size = opts->leaf_size; bo.max_rank = opts->max_rank;
    bo.pivot_threshold = pivot_threshold;
    m->blr = std::make_unique<BLREngine>(n, dA, ldA, bo, true, true);
    *S = m.release();
  });
}

static void blr_partial_impl(CSPStructMat* S, int n1, int n2, const double* A11, int ld11,
                             const double* A12, int ld12, const double* A21, int ld21, double* A22,
                             int ld22, const CSPOptions* opts, const SB200BLRParams* params, bool device) {
  require_gpu();
  auto m = std::make_unique<Mat>();
  m->type = SP_TYPE_BLR;
  BLROpts bo = blr_opts(opts, params);
  m->blr = std::make_unique<BLREngine>(n1, n2, A11, ld11, A12, ld12, A21, ld21, A22, ld22, bo, device);
  if (n2 > 0)   // A22 <- A22 - A21 A11^{-1} A12, in place like the reference
    SB200_CUDA(cudaMemcpy2D(A22, sizeof(double) * ld22, m->blr->schur(), sizeof(double) * (n1 + n2),
                            sizeof(double) * n2, n2,
                            device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost));
  *S = m.release();
}

int SB200_d_blr_partial_factor(CSPStructMat* S, int n1, int n2, const double* A11, int ld11,
                               const double* A12, int ld12, const double* A21, int ld21,
                               double* A22, int ld22, const CSPOptions* opts,
                               double pivot_threshold) {
  return guarded([&] {
    SB200BLRParams p{pivot_threshold, 1, nullptr, 0, nullptr, 0, nullptr, 0};
    blr_partial_impl(S, n1, n2, A11, ld11, A12, ld12, A21, ld21, A22, ld22, opts, &p, false);
  });
}

int SB200_d_blr_partial_factor_device(CSPStructMat* S, int n1, int n2, const double* dA11, int ld11,
                                      const double* dA12, int ld12, const double* dA21, int ld21,
                                      double* dA22, int ld22, const CSPOptions* opts,
                                      double pivot_threshold) {
  return guarded([&] {
    SB200BLRParams p{pivot_threshold, 1, nullptr, 0, nullptr, 0, nullptr, 0};
    blr_partial_impl(S, n1, n2, dA11, ld11, dA12, ld12, dA21, ld21, dA22, ld22, opts, &p, true);
  });
}

int SB200_d_blr_partial_factor_ex(CSPStructMat* S, int n1, int n2, const double* A11, int ld11,
                                  const double* A12, int ld12, const double* A21, int ld21,
                                  double* A22, int ld22, const CSPOptions* opts,
                                  const SB200BLRParams* params) {
  return guarded([&] {
    blr_partial_impl(S, n1, n2, A11, ld11, A12, ld12, A21, ld21, A22, ld22, opts, params, false);
  });
}

// ClusterTree(n).refine(leaf) (reference src/structured/ClusterTree.hpp:104-114)
static void refine_tiles(std::vector<int>& t, int size, int leaf) {
  if (size >= 2 * leaf) { refine_tiles(t, size / 2, leaf); refine_tiles(t, size - size / 2, leaf); }
  else t.push_back(size);
}

// fill the dense n x n host matrix tile by tile through the block callback (the
// granularity at which the reference calls its extract_t, BLRMatrix.cpp:91-111)
static std::vector<double> fill_by_tiles(int n, const std::vector<int>& tiles, SB200ElemBlockFn elem, void* user) {
  if (!elem) throw std::invalid_argument("no element callback");
  std::vector<double> A((size_t)n * n);
  std::vector<int> idx(n);
  for (int i = 0; i < n; i++) idx[i] = i;
  std::vector<int> off(tiles.size() + 1, 0);
  for (size_t t = 0; t < tiles.size(); t++) off[t + 1] = off[t] + tiles[t];
  for (size_t j = 0; j < tiles.size(); j++)
    for (size_t i = 0; i < tiles.size(); i++)
      elem(tiles[i], idx.data() + off[i], tiles[j], idx.data() + off[j], A.data() + off[i] + (size_t)off[j] * n, n, user);
  return A;
}

int SB200_d_blr_from_element_blocks(CSPStructMat* S, int n, SB200ElemBlockFn elem, void* user,
                                    const CSPOptions* opts, const SB200BLRParams* params, int factor) {
  return guarded([&] {
    require_gpu();
    if (n <= 0) throw std::invalid_argument("empty matrix");
    std::vector<int> tiles;
    refine_tiles(tiles, n, std::max(1, opts->leaf_size));
    std::vector<double> A = fill_by_tiles(n, tiles, elem, user);
    auto m = std::make_unique<Mat>();
    m->type = SP_TYPE_BLR;
    m->blr = std::make_unique<BLREngine>(n, A.data(), n, blr_opts(opts, params), factor != 0);
    *S = m.release();
  });
}

int SB200_d_blr_partial_factor_element_blocks(CSPStructMat* S, int n1, int n2, SB200ElemBlockFn elem,
                                              void* user, double* A22, int ld22,
                                              const CSPOptions* opts, const SB200BLRParams* params) {
  return guarded([&] {
    if (n1 <= 0 || n2 < 0) throw std::invalid_argument("partial factorization needs n1 > 0, n2 >= 0");
    const int n = n1 + n2;
    std::vector<int> tiles;
    refine_tiles(tiles, n1, std::max(1, opts->leaf_size));
    if (n2 > 0) refine_tiles(tiles, n2, std::max(1, opts->leaf_size));
    std::vector<double> A = fill_by_tiles(n, tiles, elem, user);
    const double* a = A.data();
    // the callback defines the whole front; the Schur complement goes to the caller's A22
    std::vector<double> S22((size_t)std::max(n2, 1) * std::max(n2, 1));
    for (int j = 0; j < n2; j++)
      for (int i = 0; i < n2; i++) S22[i + (size_t)j * n2] = a[(n1 + i) + (size_t)(n1 + j) * n];
    blr_partial_impl(S, n1, n2, a, n, a + (size_t)n1 * n, n, a + n1, n, S22.data(), std::max(n2, 1), opts, params, false);
    if (A22)
      for (int j = 0; j < n2; j++)
        for (int i = 0; i < n2; i++) A22[i + (size_t)j * ld22] = S22[i + (size_t)j * n2];
  });
}

int SB200_d_blr_sep_rows(const CSPStructMat S) {
  return (S && M(S)->blr) ? M(S)->blr->sep_rows() : 0;
}

int SB200_d_blr_partial_forward_solve(const CSPStructMat S, int nrhs, double* B, int ldB) {
  return guarded([&] {
    Mat* mm = M(S);
    std::lock_guard<std::mutex> lk(mm->mu);
    if (!mm->blr) throw std::logic_error("operation not supported for this type");
    const int n = mm->blr->rows();
    h2d(mm->dB, B, n, nrhs, ldB, mm->st);
    mm->blr->partial_forward(nrhs, mm->dB.p, n, mm->st);
    d2h(B, mm->dB, n, nrhs, ldB, mm->st);
    SB200_CUDA(cudaStreamSynchronize(mm->st));
  });
}

int SB200_d_blr_partial_backward_solve(const CSPStructMat S, int nrhs, double* Y, int ldY) {
  return guarded([&] {
    Mat* mm = M(S);
    std::lock_guard<std::mutex> lk(mm->mu);
    if (!mm->blr) throw std::logic_error("operation not supported for this type");
    const int n = mm->blr->rows();
    h2d(mm->dB, Y, n, nrhs, ldY, mm->st);
    mm->blr->partial_backward(nrhs, mm->dB.p, n, mm->st);
    d2h(Y, mm->dB, n, nrhs, ldY, mm->st);
    SB200_CUDA(cudaStreamSynchronize(mm->st));
  });
}

int SB200_d_blr_tiles(const CSPStructMat S) {
  return (S && M(S)->blr) ? M(S)->blr->tiles() : 0;
}

int SB200_d_blr_dense_tiles(const CSPStructMat S) {
  return (S && M(S)->blr) ? M(S)->blr->dense_tiles() : 0;
}

int SB200_d_hss_read(CSPStructMat* S, const char* path) {
  return guarded([&] {
    require_gpu();
    auto m = std::make_unique<Mat>();
    m->hss = std::make_unique<HSSEngine>(HSSHost::read_file(path));
    *S = m.release();
  });
}

int SB200_d_hss_write(const CSPStructMat S, const char* path) {
  return guarded([&] {
    hss(S).sync_host_values();
    hss(S).host().write_file(path);
  });
}

int SB200_d_hss_from_generators(CSPStructMat* S, int n_nodes,
                                const int64_t* node_tab, const double* vals,
                                int64_t n_vals, const int32_t* perms,
                                int64_t n_perms) {
  return guarded([&] {
    require_gpu();
    auto m = std::make_unique<Mat>();
    m->hss = std::make_unique<HSSEngine>(
        HSSHost::from_flat(n_nodes, node_tab, vals, n_vals, perms, n_perms));
    *S = m.release();
  });
}

int SP_d_struct_mult(const CSPStructMat S, char trans, int m, const double* B,
                     int ldB, double* C, int ldC) {
  return guarded([&] {
    Mat* mm = M(S);
    std::lock_guard<std::mutex> lk(mm->mu);
    cudaStream_t st = mm->st;
    if (mm->blr) {
      const int n = mm->blr->rows();
      h2d(mm->dB, B, n, m, ldB, st);
      mm->dC.ensure((size_t)n * m);
      mm->blr->mult(trans, m, mm->dB.p, n, mm->dC.p, n, st);
      d2h(C, mm->dC, n, m, ldC, st);
      SB200_CUDA(cudaStreamSynchronize(st));
      return;
    }
    auto& H = hss(S);
    const bool T = !(trans == 'N' || trans == 'n');
    const int nb = T ? H.rows() : H.cols(), nc = T ? H.cols() : H.rows();
    h2d(mm->dB, B, nb, m, ldB, st);
    mm->dC.ensure((size_t)nc * m);
    H.mult(trans, m, mm->dB.p, nb, mm->dC.p, nc, st);
    d2h(C, mm->dC, nc, m, ldC, st);
    SB200_CUDA(cudaStreamSynchronize(st));
  });
}

int SP_d_struct_factor(CSPStructMat S) {
  return guarded([&] {
    if (M(S)->blr)   // as in the reference: BLRMatrix has no factor() (StructuredMatrix.cpp:1539-1589)
      throw std::logic_error("factor() is not supported for a compressed BLR matrix; "
                             "use SB200_d_blr_compress_and_factor");
    Mat* mm = M(S);
    std::lock_guard<std::mutex> lk(mm->mu);
    hss(S).factor(mm->st);
    SB200_CUDA(cudaStreamSynchronize(mm->st));
  });
}

int SP_d_struct_solve(const CSPStructMat S, int nrhs, double* B, int ldB) {
  return guarded([&] {
    Mat* mm = M(S);
    std::lock_guard<std::mutex> lk(mm->mu);
    cudaStream_t st = mm->st;
    if (mm->blr) {
      const int n = mm->blr->rows();
      h2d(mm->dB, B, n, nrhs, ldB, st);
      mm->blr->solve(nrhs, mm->dB.p, n, st);
      d2h(B, mm->dB, n, nrhs, ldB, st);
      SB200_CUDA(cudaStreamSynchronize(st));
      return;
    }
    auto& H = hss(S);
    h2d(mm->dB, B, H.rows(), nrhs, ldB, st);
    H.solve(nrhs, mm->dB.p, H.rows(), st);
    d2h(B, mm->dB, H.rows(), nrhs, ldB, st);
    SB200_CUDA(cudaStreamSynchronize(st));
  });
}

int SP_d_struct_shift(CSPStructMat S, double s) {
  return guarded([&] {
    Mat* mm = M(S);
    std::lock_guard<std::mutex> lk(mm->mu);
    hss(S).shift(s, mm->st);
    SB200_CUDA(cudaStreamSynchronize(mm->st));
  });
}

int SB200_d_struct_mult_device(const CSPStructMat S, char trans, int m,
                               const double* dB, int ldB, double* dC, int ldC,
                               void* stream) {
  return guarded([&] {
    hss(S).mult(trans, m, dB, ldB, dC, ldC, static_cast<cudaStream_t>(stream));
  });
}

int SB200_d_struct_factor_device(CSPStructMat S, void* stream) {
  return guarded([&] { hss(S).factor(static_cast<cudaStream_t>(stream)); });
}

int SB200_d_struct_solve_device(const CSPStructMat S, int nrhs, double* dB,
                                int ldB, void* stream) {
  return guarded([&] {
    hss(S).solve(nrhs, dB, ldB, static_cast<cudaStream_t>(stream));
  });
}

int SB200_d_hss_apply(const CSPStructMat S, char trans, int m, const double* B,
                      int ldB, double beta, double* C, int ldC) {
  return guarded([&] {
    Mat* mm = M(S);
    std::lock_guard<std::mutex> lk(mm->mu);
    auto& H = hss(S);
    cudaStream_t st = mm->st;
    const bool T = !(trans == 'N' || trans == 'n');
    const int nb = T ? H.rows() : H.cols(), nc = T ? H.cols() : H.rows();
    h2d(mm->dB, B, nb, m, ldB, st);
    if (beta != 0.) h2d(mm->dC, C, nc, m, ldC, st);
    else mm->dC.ensure((size_t)nc * m);
    H.mult(trans, m, mm->dB.p, nb, mm->dC.p, nc, st, beta);
    d2h(C, mm->dC, nc, m, ldC, st);
    SB200_CUDA(cudaStreamSynchronize(st));
  });
}

int SB200_d_hss_apply_device(const CSPStructMat S, char trans, int m, const double* dB,
                             int ldB, double beta, double* dC, int ldC, void* stream) {
  return guarded([&] {
    hss(S).mult(trans, m, dB, ldB, dC, ldC, static_cast<cudaStream_t>(stream), beta);
  });
}

int SB200_d_hss_extract(const CSPStructMat S, int nI, const int* I, int nJ, const int* J,
                        double* B, int ldB, int add) {
  return guarded([&] {
    if (nI <= 0 || nJ <= 0) return;
    Mat* mm = M(S);
    std::lock_guard<std::mutex> lk(mm->mu);
    auto& H = hss(S);
    if (add) h2d(mm->dC, B, nI, nJ, ldB, mm->st);
    else mm->dC.ensure((size_t)nI * nJ);
    H.extract(nI, I, nJ, J, mm->dC.p, nI, add != 0, mm->st);
    d2h(B, mm->dC, nI, nJ, ldB, mm->st);
    SB200_CUDA(cudaStreamSynchronize(mm->st));
  });
}

int SB200_d_hss_forward_solve(const CSPStructMat S, int nrhs, const double* B, int ldB) {
  return guarded([&] {
    Mat* mm = M(S);
    std::lock_guard<std::mutex> lk(mm->mu);
    auto& H = hss(S);
    h2d(mm->dB, B, H.rows(), nrhs, ldB, mm->st);
    H.forward_solve(nrhs, mm->dB.p, H.rows(), mm->st);
    SB200_CUDA(cudaStreamSynchronize(mm->st));
  });
}

int SB200_d_hss_backward_solve(const CSPStructMat S, int nrhs, double* X, int ldX) {
  return guarded([&] {
    Mat* mm = M(S);
    std::lock_guard<std::mutex> lk(mm->mu);
    auto& H = hss(S);
    mm->dB.ensure((size_t)H.rows() * nrhs);
    H.backward_solve(nrhs, mm->dB.p, H.rows(), mm->st);
    d2h(X, mm->dB, H.rows(), nrhs, ldX, mm->st);
    SB200_CUDA(cudaStreamSynchronize(mm->st));
  });
}

int SB200_d_hss_forward_solve_device(const CSPStructMat S, int nrhs, double* dB, int ldB,
                                     void* stream) {
  return guarded([&] { hss(S).forward_solve(nrhs, dB, ldB, static_cast<cudaStream_t>(stream)); });
}

int SB200_d_hss_backward_solve_device(const CSPStructMat S, int nrhs, double* dX, int ldX,
                                      void* stream) {
  return guarded([&] { hss(S).backward_solve(nrhs, dX, ldX, static_cast<cudaStream_t>(stream)); });
}

/* ---- Schur complement of the (0,0) block (HSS fronts) --------------------- */
int SB200_d_hss_partial_factor(CSPStructMat S) {
  return guarded([&] {
    Mat* mm = M(S);
    std::lock_guard<std::mutex> lk(mm->mu);
    hss(S).partial_factor(mm->st);
    SB200_CUDA(cudaStreamSynchronize(mm->st));
  });
}

/* device-resident variants: every matrix argument is a DEVICE pointer, the work
 * is queued on `stream` and not synchronised */
int SB200_d_hss_partial_factor_device(CSPStructMat S, void* stream) {
  return guarded([&] { hss(S).partial_factor(static_cast<cudaStream_t>(stream)); });
}

int SB200_d_hss_schur_update_device(const CSPStructMat S, double* dTheta, int ldT, double* dDUB01,
                                    int ldD, double* dPhi, int ldP, void* stream) {
  return guarded([&] {
    hss(S).schur_update(dTheta, ldT, dDUB01, ldD, dPhi, ldP, static_cast<cudaStream_t>(stream));
  });
}

int SB200_d_hss_schur_product_direct_device(const CSPStructMat S, const double* dTheta, int ldT,
                                            const double* dDUB01, int ldD, const double* dPhi,
                                            int ldP, int c, const double* dR, int ldR, double* dSr,
                                            int ldSr, double* dSc, int ldSc, void* stream) {
  return guarded([&] {
    hss(S).schur_product_direct(dTheta, ldT, dDUB01, ldD, dPhi, ldP, c, dR, ldR, dSr, ldSr, dSc, ldSc,
                                static_cast<cudaStream_t>(stream));
  });
}

int SB200_d_hss_schur_sizes(const CSPStructMat S, int* out) {
  return guarded([&] { hss(S).schur_sizes(out); });
}

int SB200_d_hss_schur_update(const CSPStructMat S, double* Theta, int ldT, double* DUB01,
                             int ldD, double* Phi, int ldP) {
  return guarded([&] {
    Mat* mm = M(S);
    std::lock_guard<std::mutex> lk(mm->mu);
    auto& H = hss(S);
    int z[7];
    H.schur_sizes(z);
    const int rows1 = z[0], cols1 = z[1], rv0 = z[2], m0 = z[3], rv1 = z[4];
    mm->dTheta.ensure((size_t)std::max(rows1, 1) * std::max(rv0, 1));
    mm->dDUB01.ensure((size_t)std::max(m0, 1) * std::max(rv1, 1));
    mm->dPhi.ensure((size_t)std::max(cols1, 1) * std::max(m0, 1));
    H.schur_update(mm->dTheta.p, std::max(rows1, 1), mm->dDUB01.p, std::max(m0, 1), mm->dPhi.p,
                   std::max(cols1, 1), mm->st);
    if (Theta && rows1 && rv0) d2h(Theta, mm->dTheta, rows1, rv0, ldT, mm->st);
    if (DUB01 && m0 && rv1) d2h(DUB01, mm->dDUB01, m0, rv1, ldD, mm->st);
    if (Phi && cols1 && m0) d2h(Phi, mm->dPhi, cols1, m0, ldP, mm->st);
    SB200_CUDA(cudaStreamSynchronize(mm->st));
  });
}

int SB200_d_hss_vhat(const CSPStructMat S, double* Vhat, int ldV) {
  return guarded([&] {
    auto& H = hss(S);
    if (!H.partially_factored()) throw std::logic_error("Vhat requested before partial_factor");
    int z[7];
    H.schur_sizes(z);
    if (z[3] && z[2])
      SB200_CUDA(cudaMemcpy2D(Vhat, sizeof(double) * ldV, H.vhat(), sizeof(double) * z[3],
                              sizeof(double) * z[3], z[2], cudaMemcpyDeviceToHost));
  });
}

int SB200_d_hss_schur_product_direct(const CSPStructMat S, const double* Theta, int ldT,
                                     const double* DUB01, int ldD, const double* Phi, int ldP,
                                     int c, const double* R, int ldR, double* Sr, int ldSr,
                                     double* Sc, int ldSc) {
  return guarded([&] {
    Mat* mm = M(S);
    std::lock_guard<std::mutex> lk(mm->mu);
    auto& H = hss(S);
    int z[7];
    H.schur_sizes(z);
    const int rows1 = z[0], cols1 = z[1], rv0 = z[2], m0 = z[3], rv1 = z[4];
    if (c <= 0) return;
    // NULL: use the device copies kept by the last SB200_d_hss_schur_update
    if (Theta && rows1 && rv0) h2d(mm->dTheta, Theta, rows1, rv0, ldT, mm->st);
    if (DUB01 && m0 && rv1) h2d(mm->dDUB01, DUB01, m0, rv1, ldD, mm->st);
    if (Phi && cols1 && m0) h2d(mm->dPhi, Phi, cols1, m0, ldP, mm->st);
    if (!mm->dTheta.p || !mm->dDUB01.p || !mm->dPhi.p)
      throw std::logic_error("Schur_product_direct: no Theta / DUB01 / Phi (call Schur_update first)");
    h2d(mm->dS[0], R, rows1, c, ldR, mm->st);
    mm->dS[1].ensure((size_t)rows1 * c);
    mm->dS[2].ensure((size_t)cols1 * c);
    H.schur_product_direct(mm->dTheta.p, std::max(rows1, 1), mm->dDUB01.p, std::max(m0, 1),
                           mm->dPhi.p, std::max(cols1, 1), c, mm->dS[0].p, rows1, mm->dS[1].p,
                           rows1, mm->dS[2].p, cols1, mm->st);
    d2h(Sr, mm->dS[1], rows1, c, ldSr, mm->st);
    d2h(Sc, mm->dS[2], cols1, c, ldSc, mm->st);
    SB200_CUDA(cudaStreamSynchronize(mm->st));
  });
}

int SB200_d_hss_schur_product_indirect(const CSPStructMat S, const double* DUB01, int ldD, int c,
                                       const double* R0, int ldR0, const double* R1, int ldR1,
                                       const double* Sr1, int ldSr1, const double* Sc1, int ldSc1,
                                       double* Sr, int ldSr, double* Sc, int ldSc) {
  return guarded([&] {
    Mat* mm = M(S);
    std::lock_guard<std::mutex> lk(mm->mu);
    auto& H = hss(S);
    int z[7];
    H.schur_sizes(z);
    const int rows1 = z[0], cols1 = z[1], m0 = z[3], rv1 = z[4], rows0 = z[6];
    if (c <= 0) return;
    if (DUB01 && m0 && rv1) h2d(mm->dDUB01, DUB01, m0, rv1, ldD, mm->st);
    if (!mm->dDUB01.p) throw std::logic_error("Schur_product_indirect: no DUB01 (call Schur_update first)");
    h2d(mm->dS[0], R0, rows0, c, ldR0, mm->st);
    h2d(mm->dS[3], R1, rows1, c, ldR1, mm->st);
    h2d(mm->dS[1], Sr1, rows1, c, ldSr1, mm->st);
    h2d(mm->dS[2], Sc1, cols1, c, ldSc1, mm->st);
    H.schur_product_indirect(mm->dDUB01.p, std::max(m0, 1), c, mm->dS[0].p, rows0, mm->dS[3].p, rows1,
                             mm->dS[1].p, rows1, mm->dS[2].p, cols1, mm->dS[1].p, rows1,
                             mm->dS[2].p, cols1, mm->st);
    d2h(Sr, mm->dS[1], rows1, c, ldSr, mm->st);
    d2h(Sc, mm->dS[2], cols1, c, ldSc, mm->st);
    SB200_CUDA(cudaStreamSynchronize(mm->st));
  });
}

int SB200_d_hss_partial_forward_solve(const CSPStructMat S, int nrhs, const double* B0, int ldB,
                                      double* reduced_rhs, int ldR) {
  return guarded([&] {
    Mat* mm = M(S);
    std::lock_guard<std::mutex> lk(mm->mu);
    auto& H = hss(S);
    int z[7];
    H.schur_sizes(z);
    const int rv0 = z[2], rows0 = z[6];
    if (nrhs <= 0) return;
    h2d(mm->dS[4], B0, rows0, nrhs, ldB, mm->st);
    mm->dS[5].ensure((size_t)std::max(rv0, 1) * nrhs);
    H.partial_forward_solve(nrhs, mm->dS[4].p, rows0, mm->dS[5].p, std::max(rv0, 1), mm->st);
    if (rv0 && reduced_rhs) d2h(reduced_rhs, mm->dS[5], rv0, nrhs, ldR, mm->st);
    SB200_CUDA(cudaStreamSynchronize(mm->st));
  });
}

int SB200_d_hss_partial_x(const CSPStructMat S, int nrhs, double* X, int ldX, int set) {
  return guarded([&] {
    auto& H = hss(S);
    int z[7];
    H.schur_sizes(z);
    const int m0 = z[3];
    double* dx = H.partial_x(nrhs);
    if (!m0) return;
    if (set)
      SB200_CUDA(cudaMemcpy2D(dx, sizeof(double) * m0, X, sizeof(double) * ldX, sizeof(double) * m0,
                              nrhs, cudaMemcpyHostToDevice));
    else
      SB200_CUDA(cudaMemcpy2D(X, sizeof(double) * ldX, dx, sizeof(double) * m0, sizeof(double) * m0,
                              nrhs, cudaMemcpyDeviceToHost));
  });
}

int SB200_d_hss_partial_backward_solve(const CSPStructMat S, int nrhs, double* X0, int ldX) {
  return guarded([&] {
    Mat* mm = M(S);
    std::lock_guard<std::mutex> lk(mm->mu);
    auto& H = hss(S);
    int z[7];
    H.schur_sizes(z);
    const int rows0 = z[6];
    if (nrhs <= 0) return;
    mm->dS[4].ensure((size_t)rows0 * nrhs);
    H.partial_backward_solve(nrhs, mm->dS[4].p, rows0, mm->st);
    d2h(X0, mm->dS[4], rows0, nrhs, ldX, mm->st);
    SB200_CUDA(cudaStreamSynchronize(mm->st));
  });
}

/* ---- sharded operations with the NCCL exchange inside the engine ---------- */
int SB200_nccl_unique_id(char* out128) {
  return guarded([&] { HSSEngine::nccl_unique_id(out128); });
}
int SB200_d_hss_dist_init(CSPStructMat S, int nparts, int part, const char* unique_id128) {
  return guarded([&] { require_gpu(); hss(S).dist_init(nparts, part, unique_id128); });
}
int SB200_d_hss_dist_mult(const CSPStructMat S, char trans, int m, const double* dB, int ldB, double* dC,
                          int ldC, void* stream) {
  return guarded([&] { hss(S).dist_mult(trans, m, dB, ldB, dC, ldC, static_cast<cudaStream_t>(stream)); });
}
int SB200_d_hss_dist_factor(CSPStructMat S, void* stream) {
  return guarded([&] { hss(S).dist_factor(static_cast<cudaStream_t>(stream)); });
}
int SB200_d_hss_dist_solve(const CSPStructMat S, int nrhs, double* dB, int ldB, void* stream) {
  return guarded([&] { hss(S).dist_solve(nrhs, dB, ldB, static_cast<cudaStream_t>(stream)); });
}

int SB200_d_hss_set_partition(CSPStructMat S, int nparts, int part) {
  return guarded([&] { hss(S).set_partition(nparts, part); });
}
int SB200_d_hss_owned_range(const CSPStructMat S, int* lo, int* hi) {
  return guarded([&] { hss(S).owned_range(lo, hi); });
}
int SB200_d_hss_dist_sizes(const CSPStructMat S, int nrhs, long long int* out) {
  return guarded([&] { hss(S).dist_sizes(nrhs, out); });
}
int SB200_d_hss_dist_mult_begin(const CSPStructMat S, char trans, int m,
                                const double* dB, int ldB, double* dSend, void* stream) {
  return guarded([&] { hss(S).dist_mult_begin(trans, m, dB, ldB, dSend, static_cast<cudaStream_t>(stream)); });
}
int SB200_d_hss_dist_mult_end(const CSPStructMat S, char trans, int m, const double* dB,
                              int ldB, double* dC, int ldC, const double* dRecv, void* stream) {
  return guarded([&] { hss(S).dist_mult_end(trans, m, dB, ldB, dC, ldC, dRecv, static_cast<cudaStream_t>(stream)); });
}
int SB200_d_hss_dist_factor_begin(CSPStructMat S, double* dSend, void* stream) {
  return guarded([&] { hss(S).dist_factor_begin(dSend, static_cast<cudaStream_t>(stream)); });
}
int SB200_d_hss_dist_factor_end(CSPStructMat S, const double* dRecv, void* stream) {
  return guarded([&] { hss(S).dist_factor_end(dRecv, static_cast<cudaStream_t>(stream)); });
}
int SB200_d_hss_dist_solve_begin(const CSPStructMat S, int nrhs, double* dB, int ldB,
                                 double* dSend, void* stream) {
  return guarded([&] { hss(S).dist_solve_begin(nrhs, dB, ldB, dSend, static_cast<cudaStream_t>(stream)); });
}
int SB200_d_hss_dist_solve_end(const CSPStructMat S, int nrhs, double* dB, int ldB,
                               const double* dRecv, void* stream) {
  return guarded([&] { hss(S).dist_solve_end(nrhs, dB, ldB, dRecv, static_cast<cudaStream_t>(stream)); });
}

int SB200_d_hss_file_info(const char* path, long long int* out) {
  return guarded([&] {
    HSSHost h = HSSHost::read_file(path);
    out[0] = h.rows(); out[1] = h.cols(); out[2] = (long long)h.nodes.size();
    out[3] = h.levels(); out[4] = h.max_rank(); out[5] = h.nonzeros();
    out[6] = h.apply_flops(); out[7] = h.factor_flops_ref();
    out[8] = h.solve_flops_ref(); out[9] = h.factor_flops_exec();
  });
}

double SB200_fp64_dmma_peak_tflops(void) {
  double v = 0.;
  guarded([&] { require_gpu(); v = measure_fp64_dmma_peak_tflops(); });
  return v;
}

int SB200_debug_qr_batch(int m, int k, int naug, int count, const double* A, double* out, double* T,
                         int variant, int reps, float* ms) {
  return guarded([&] {
    require_gpu();
    debug_qr_batch(m, k, naug, count, A, out, T, variant, reps, ms);
  });
}

int SB200_d_hss_file_copy(const char* in_path, const char* out_path) {
  return guarded([&] { HSSHost::read_file(in_path).write_file(out_path); });
}

// HSS-only statistics: 0 for a BLR handle (queries never throw across the C boundary)
int SB200_d_struct_levels(const CSPStructMat S) {
  const Mat* m = static_cast<const Mat*>(S);
  return m && m->hss ? m->hss->host().levels() : 0;
}
long long int SB200_d_struct_factor_nonzeros(const CSPStructMat S) {
  const Mat* m = static_cast<const Mat*>(S);
  return m && m->hss ? m->hss->factor_nonzeros() : 0;
}
int SB200_d_struct_ulv_data(const CSPStructMat S, double* factors, double* tfactors,
                            long long int* sizes) {
  return guarded([&] {
    if (!S) throw std::invalid_argument("null handle");
    hss(S).export_ulv(factors, tfactors, sizes);
  });
}
long long int SB200_d_struct_flops(const CSPStructMat S, int which) {
  const Mat* m = static_cast<const Mat*>(S);
  if (!m || !m->hss) return 0;
  const auto& h = m->hss->host();
  switch (which) {
    case 0: return h.apply_flops();
    case 1: return h.factor_flops_ref();
    case 2: return h.solve_flops_ref();
    case 3: return h.factor_flops_exec();
    case 4: return h.qr_class_flops_ref(0);
    case 5: return h.qr_class_flops_exec(0);
  }
  return 0;
}
int SB200_d_struct_set_profile(CSPStructMat S, int on) {
  return guarded([&] { hss(S).set_profile(on != 0); });
}
double SB200_d_struct_kernel_ms(const CSPStructMat S, int which) {
  double ms = 0.;
  guarded([&] { if (which == 0) ms = hss(S).qr_leaf_ms(); });
  return ms;
}
long long int SB200_d_struct_launches(const CSPStructMat S) {
  const Mat* m = static_cast<const Mat*>(S);
  if (!m) return 0;
  return m->blr ? m->blr->launches() : m->hss ? m->hss->launches() : 0;
}
int SB200_d_struct_print_info(const CSPStructMat S) {
  return guarded([&] { hss(S).host().print_info(); });
}
int SB200_d_struct_dense(const CSPStructMat S, double* A, int ldA) {
  return guarded([&] {
    if (M(S)->blr) {     // BLRMatrix::dense() (BLRMatrix.cpp:296-305): B * I in slabs, compressed matrices only
      auto& B = *M(S)->blr;
      const int n = B.rows(), sb = 128;
      DevBuf<double> I((size_t)n * sb), Y((size_t)n * sb);
      std::vector<double> hI((size_t)n * sb);
      for (int c0 = 0; c0 < n; c0 += sb) {
        const int nc = std::min(sb, n - c0);
        std::fill(hI.begin(), hI.end(), 0.);
        for (int c = 0; c < nc; c++) hI[(size_t)c * n + c0 + c] = 1.;
        SB200_CUDA(cudaMemcpy(I.p, hI.data(), sizeof(double) * (size_t)n * nc, cudaMemcpyHostToDevice));
        B.mult('N', nc, I.p, n, Y.p, n, 0);
        SB200_CUDA(cudaMemcpy2D(A + (size_t)c0 * ldA, sizeof(double) * ldA, Y.p, sizeof(double) * n,
                                sizeof(double) * n, nc, cudaMemcpyDeviceToHost));
      }
      return;
    }
    auto& H = hss(S);
    const int n = H.cols(), m = H.rows();
    // H * I, in slabs of 256 columns
    DevBuf<double> I, Y;
    const int sb = 256;
    I.alloc((size_t)n * sb);
    Y.alloc((size_t)m * sb);
    std::vector<double> hI((size_t)n * sb);
    for (int c0 = 0; c0 < n; c0 += sb) {
      const int nc = std::min(sb, n - c0);
      std::fill(hI.begin(), hI.end(), 0.);
      for (int c = 0; c < nc; c++) hI[(size_t)c * n + c0 + c] = 1.;
      SB200_CUDA(cudaMemcpy(I.p, hI.data(), sizeof(double) * (size_t)n * nc,
                            cudaMemcpyHostToDevice));
      H.mult('N', nc, I.p, n, Y.p, m, 0);
      SB200_CUDA(cudaMemcpy2D(A + (size_t)c0 * ldA, sizeof(double) * ldA, Y.p,
                              sizeof(double) * m, sizeof(double) * m, nc,
                              cudaMemcpyDeviceToHost));
    }
  });
}

/* ---- single precision interface (reference StructuredMatrix.h SP_s_*) -------
 * float at the boundary, fp64 inside: operands are widened on the way in and
 * rounded on the way out, every kernel is the double precision one.  (B200's
 * fp64 tensor pipe is what the engine is built on; the results are at least
 * as accurate as a float implementation's.) */
void SP_s_struct_default_options(CSPOptions* o) {
  SP_d_struct_default_options(o);
  // StructuredOptions<float>: default_structured_rel_tol / abs_tol specialisations
  // (reference StructuredOptions.hpp:49-54)
  o->rel_tol = 1e-2;
  o->abs_tol = 1e-5;
}
void SP_s_struct_destroy(CSPStructMat* S) { SP_d_struct_destroy(S); }
int SP_s_struct_rows(const CSPStructMat S) { return SP_d_struct_rows(S); }
int SP_s_struct_cols(const CSPStructMat S) { return SP_d_struct_cols(S); }
/* bytes / nonzeros of the object as it is stored (fp64) */
long long int SP_s_struct_memory(const CSPStructMat S) { return SP_d_struct_memory(S); }
long long int SP_s_struct_nonzeros(const CSPStructMat S) { return SP_d_struct_nonzeros(S); }
int SP_s_struct_rank(const CSPStructMat S) { return SP_d_struct_rank(S); }

int SP_s_struct_from_dense(CSPStructMat* S, int rows, int cols, const float* A, int ldA,
                           const CSPOptions* opts) {
  if (rows < 0 || cols < 0 || !A) return guarded([] { throw std::invalid_argument("from_dense: bad arguments"); });
  std::vector<double> D = widen(A, rows, cols, ldA);
  return SP_d_struct_from_dense(S, rows, cols, D.data(), std::max(rows, 1), opts);
}

int SP_s_struct_from_elements(CSPStructMat* S, int rows, int cols, float A(int i, int j),
                              const CSPOptions* opts) {
  if (rows < 0 || cols < 0 || !A) return guarded([] { throw std::invalid_argument("from_elements: bad arguments"); });
  if ((long long)rows * cols > (1LL << 28))
    return guarded([] { throw std::invalid_argument("from_elements: host callbacks are limited to n <= 16384"); });
  std::vector<double> D((size_t)rows * cols);
  for (int j = 0; j < cols; j++)
    for (int i = 0; i < rows; i++) D[i + (size_t)j * rows] = A(i, j);
  return SP_d_struct_from_dense(S, rows, cols, D.data(), std::max(rows, 1), opts);
}

int SP_s_struct_mult(const CSPStructMat S, char trans, int m, const float* B, int ldB, float* C,
                     int ldC) {
  if (!S || m <= 0) return S ? 0 : guarded([] { throw std::invalid_argument("null CSPStructMat"); });
  const bool T = !(trans == 'N' || trans == 'n');
  const int r = SP_d_struct_rows(S), c = SP_d_struct_cols(S);
  const int nb = T ? r : c, nc = T ? c : r;
  std::vector<double> Bd = widen(B, nb, m, ldB), Cd((size_t)nc * m);
  const int rc = SP_d_struct_mult(S, trans, m, Bd.data(), std::max(nb, 1), Cd.data(), std::max(nc, 1));
  if (!rc) narrow(Cd, C, nc, m, ldC);
  return rc;
}

int SP_s_struct_factor(CSPStructMat S) { return SP_d_struct_factor(S); }

int SP_s_struct_solve(const CSPStructMat S, int nrhs, float* B, int ldB) {
  if (!S || nrhs <= 0) return S ? 0 : guarded([] { throw std::invalid_argument("null CSPStructMat"); });
  const int n = SP_d_struct_rows(S);
  std::vector<double> Bd = widen(B, n, nrhs, ldB);
  const int rc = SP_d_struct_solve(S, nrhs, Bd.data(), std::max(n, 1));
  if (!rc) narrow(Bd, B, n, nrhs, ldB);
  return rc;
}

int SP_s_struct_shift(CSPStructMat S, float s) { return SP_d_struct_shift(S, (double)s); }

}  // extern "C"
