/*
 * ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product.
 *
 * Thin C-ABI driver around the UNMODIFIED reference classes, compiled from
 * the sources where they lie under /root/reference (see oracle/Makefile).
 * It lets tests/ and bench.py's cpu_baseline / --impl reference legs
 *   - build an HSS matrix with the reference's own compression
 *     (HSSMatrix(A,opts)            reference src/HSS/HSSMatrix.cpp:49-54,
 *      HSSMatrix(Kernel&,opts)      reference src/HSS/HSSMatrix.cpp:88-106),
 *   - dump it with the reference's own HSSMatrix::write
 *                                   (reference src/HSS/HSSMatrix.cpp:438-486),
 *   - run mult / factor / solve     (reference src/HSS/HSSMatrix.apply.hpp:34,
 *                                    .factor.hpp:35, .solve.hpp:35),
 *   - read the reference's flop counters (src/StrumpackParameters.hpp:85-98),
 *   - run BLR compress_and_factor + solve (src/BLR/BLRMatrix.cpp:113, .hpp:118),
 *   - run partial_factor / Schur_update / Schur_product_direct and the partial
 *     forward / backward solves of child(0) the way FrontHSS does
 *     (src/HSS/HSSMatrix.factor.hpp:44, HSSMatrix.Schur.hpp:40,73,
 *      src/sparse/fronts/FrontHSS.cpp:385-410,452-462,487-495).
 * Nothing here restates an algorithm: every numeric result comes from the
 * reference code itself.
 */
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <random>
#include <sstream>
#include <string>
#include <vector>
#include <omp.h>

#include "HSS/HSSMatrix.hpp"
#include "BLR/BLRMatrix.hpp"
#include "kernel/Kernel.hpp"
#include "structured/ClusterTree.hpp"
#include "StrumpackParameters.hpp"

using namespace strumpack;
using HSS::HSSMatrix;
using HSS::HSSOptions;
using DenseD = DenseMatrix<double>;
using DenseW = DenseMatrixWrapper<double>;

namespace {
  struct HSSHandle {
    std::unique_ptr<HSSMatrix<double>> H;
    // state of a partial factorization, as FrontHSS keeps it
    DenseD Theta, DUB01, Phi, TV;
    std::unique_ptr<HSS::WorkSolve<double>> w;
  };
  void copy_out(const DenseD& A, double* dst) {
    if (!dst) return;
    for (std::size_t j=0; j<A.cols(); j++)
      for (std::size_t i=0; i<A.rows(); i++)
        dst[i + j*A.rows()] = A(i, j);
  }
  struct BLRHandle {
    std::unique_ptr<BLR::BLRMatrix<double>> B;
    int n = 0;
    // partially factored front (construct_and_partial_factor)
    BLR::BLRMatrix<double> B11, B12, B21;
    int n1 = 0, n2 = 0;
  };

  // "--hss_leaf_size 128 --hss_rel_tol 1e-4" -> argc/argv for the reference's
  // own getopt parser (HSSOptions.cpp:67-240)
  template<typename Opts> void parse(Opts& o, const char* args) {
    std::vector<std::string> tok{"oracle"};
    std::istringstream is(args ? args : "");
    for (std::string t; is >> t;) tok.push_back(t);
    std::vector<char*> argv;
    for (auto& t : tok) argv.push_back(const_cast<char*>(t.c_str()));
    argv.push_back(nullptr);
    o.set_from_command_line(int(tok.size()), argv.data());
  }
}

extern "C" {

  void ref_set_num_threads(int t) { omp_set_num_threads(t); }
  int ref_get_max_threads() { return omp_get_max_threads(); }

  void ref_flops_reset() {
    params::flops = 0;
    params::ULV_factor_flops = 0;
    params::hss_solve_flops = 0;
  }
  long long ref_flops_total() { return params::flops.load(); }
  long long ref_flops_ulv_factor() { return params::ULV_factor_flops.load(); }
  long long ref_flops_hss_solve() { return params::hss_solve_flops.load(); }

  /* kind: 'T' Toeplitz A(i,j)=1/(1+|i-j|), diag 1 (test_HSS_seq.cpp:69-79)
   *       'U' upper triangular Toeplitz            (test_HSS_seq.cpp:80-91) */
  void* ref_hss_toeplitz(int n, int kind, const char* hss_args) {
    DenseD A(n, n);
    for (int j=0; j<n; j++)
      for (int i=0; i<n; i++) {
        double v = (i==j) ? 1. : 1./(1+std::abs(i-j));
        if (kind == 'U' && i > j) v = 0.;
        A(i, j) = v;
      }
    HSSOptions<double> opts;
    opts.set_verbose(false);
    parse(opts, hss_args);
    auto h = new HSSHandle;
    h->H.reset(new HSSMatrix<double>(A, opts));
    return h;
  }

  void* ref_hss_dense(int m, int n, const double* A, int lda,
                      const char* hss_args) {
    DenseD Ad(m, n, A, lda);
    HSSOptions<double> opts;
    opts.set_verbose(false);
    parse(opts, hss_args);
    auto h = new HSSHandle;
    h->H.reset(new HSSMatrix<double>(Ad, opts));
    return h;
  }

  /* Gaussian kernel exp(-|x-y|^2/(2h^2)) + lambda*I on n points of dimension
   * d (pts: d x n column-major, PERMUTED IN PLACE by the clustering,
   * HSSMatrix.cpp:93-95). perm1 (length n) receives the reference's 1-based
   * permutation vector (Kernel::permutation()). */
  void* ref_hss_gauss(int n, int d, double* pts, double h, double lambda,
                      const char* hss_args, int* perm1) {
    DenseW data(d, n, pts, d);
    kernel::GaussKernel<double> K(data, h, lambda);
    HSSOptions<double> opts;
    opts.set_verbose(false);
    parse(opts, hss_args);
    auto hd = new HSSHandle;
    hd->H.reset(new HSSMatrix<double>(K, opts));
    if (perm1)
      for (int i=0; i<n; i++) perm1[i] = K.permutation()[i];
    return hd;
  }

  void* ref_hss_read(const char* path) {
    auto h = new HSSHandle;
    h->H.reset(new HSSMatrix<double>(HSSMatrix<double>::read(path)));
    return h;
  }

  int ref_hss_write(void* hv, const char* path) {
    static_cast<HSSHandle*>(hv)->H->write(std::string(path));
    return 0;
  }

  /* out[0..7] = rows, cols, rank, levels, nonzeros, factor_nonzeros, memory,
   *             is_compressed */
  void ref_hss_info(void* hv, long long* out) {
    auto& H = *static_cast<HSSHandle*>(hv)->H;
    out[0] = H.rows(); out[1] = H.cols(); out[2] = H.rank();
    out[3] = H.levels(); out[4] = H.nonzeros();
    out[5] = H.factor_nonzeros(); out[6] = H.memory();
    out[7] = H.is_compressed();
  }

  void ref_hss_print_info(void* hv) {
    static_cast<HSSHandle*>(hv)->H->print_info();
  }

  /* y = op(H) x ; trans: 0 = N, 1 = C */
  void ref_hss_mult(void* hv, int trans, int s, const double* x, int ldx,
                    double* y, int ldy) {
    auto& H = *static_cast<HSSHandle*>(hv)->H;
    int nx = trans ? H.rows() : H.cols(), ny = trans ? H.cols() : H.rows();
    DenseW X(nx, s, const_cast<double*>(x), ldx), Y(ny, s, y, ldy);
    H.mult(trans ? Trans::C : Trans::N, X, Y);
  }

  void ref_hss_factor(void* hv) { static_cast<HSSHandle*>(hv)->H->factor(); }

  void ref_hss_solve(void* hv, int s, double* b, int ldb) {
    auto& H = *static_cast<HSSHandle*>(hv)->H;
    DenseW B(H.rows(), s, b, ldb);
    H.solve(B);
  }

  void ref_hss_shift(void* hv, double sigma) {
    static_cast<HSSHandle*>(hv)->H->shift(sigma);
  }

  void ref_hss_dense_out(void* hv, double* A, int lda) {
    auto& H = *static_cast<HSSHandle*>(hv)->H;
    auto D = H.dense();
    for (std::size_t j=0; j<D.cols(); j++)
      std::memcpy(A + j*std::size_t(lda), D.ptr(0, j), sizeof(double)*D.rows());
  }

  double ref_hss_get(void* hv, int i, int j) {
    return static_cast<HSSHandle*>(hv)->H->get(i, j);
  }

  void ref_hss_destroy(void* hv) { delete static_cast<HSSHandle*>(hv); }

  /* ---- Schur complement of the (0,0) block, exactly the call sequence of
   *      FrontHSS::multifrontal_factorization (FrontHSS.cpp:385-410) */
  void ref_hss_partial_factor(void* hv) {
    auto* h = static_cast<HSSHandle*>(hv);
    auto& H = *h->H;
    H.partial_factor();
    H.Schur_update(h->Theta, h->DUB01, h->Phi);
    const DenseD& Vhat = H.child(0)->ULV().Vhat();
    if (h->Theta.cols() < h->Phi.cols()) {
      h->TV = DenseD(Vhat.cols(), h->Phi.rows());
      gemm(Trans::C, Trans::C, 1., Vhat, h->Phi, 0., h->TV);
    } else {
      h->TV = DenseD(h->Theta.rows(), Vhat.rows());
      gemm(Trans::N, Trans::C, 1., h->Theta, Vhat, 0., h->TV);
    }
  }

  /* out[0..7] = Theta rows, cols, DUB01 rows, cols, Phi rows, cols, Vhat rows, cols */
  void ref_hss_schur_sizes(void* hv, long long* out) {
    auto* h = static_cast<HSSHandle*>(hv);
    const DenseD& Vhat = h->H->child(0)->ULV().Vhat();
    out[0] = h->Theta.rows(); out[1] = h->Theta.cols();
    out[2] = h->DUB01.rows(); out[3] = h->DUB01.cols();
    out[4] = h->Phi.rows(); out[5] = h->Phi.cols();
    out[6] = Vhat.rows(); out[7] = Vhat.cols();
  }

  /* packed column-major copies; any pointer may be null */
  void ref_hss_schur_get(void* hv, double* Theta, double* DUB01, double* Phi,
                         double* Vhat) {
    auto* h = static_cast<HSSHandle*>(hv);
    copy_out(h->Theta, Theta);
    copy_out(h->DUB01, DUB01);
    copy_out(h->Phi, Phi);
    copy_out(h->H->child(0)->ULV().Vhat(), Vhat);
  }

  /* FrontHSS::sample_CB_direct (FrontHSS.cpp:211-221) */
  void ref_hss_schur_product_direct(void* hv, int c, const double* R, int ldR,
                                    double* Sr, int ldSr, double* Sc, int ldSc) {
    auto* h = static_cast<HSSHandle*>(hv);
    auto n1 = h->H->child(1)->rows();
    DenseD Rd(n1, c, R, ldR), Srd(n1, c), Scd(n1, c);
    h->H->Schur_product_direct(h->Theta, h->DUB01, h->Phi, h->TV, Rd, Srd, Scd);
    for (int j=0; j<c; j++)
      for (std::size_t i=0; i<n1; i++) {
        Sr[i + std::size_t(j)*ldSr] = Srd(i, j);
        Sc[i + std::size_t(j)*ldSc] = Scd(i, j);
      }
  }

  /* FrontHSS::fwd_solve_node (FrontHSS.cpp:452-462): b0 is rows(child 0) x s;
   * red receives reduced_rhs (V_rank(child 0) x s, packed) */
  void ref_hss_partial_forward(void* hv, int s, const double* b0, int ldb,
                               double* red) {
    auto* h = static_cast<HSSHandle*>(hv);
    auto n0 = h->H->child(0)->rows();
    DenseD B(n0, s, b0, ldb);
    h->w.reset(new HSS::WorkSolve<double>());
    h->H->child(0)->forward_solve(*h->w, B, true);
    copy_out(h->w->reduced_rhs, red);
  }

  /* w.x (rows = size of child 0's reduced block): get (set=0) or overwrite */
  int ref_hss_partial_x_rows(void* hv) {
    auto* h = static_cast<HSSHandle*>(hv);
    return h->w ? int(h->w->x.rows()) : 0;
  }
  void ref_hss_partial_x(void* hv, double* x, int set) {
    auto* h = static_cast<HSSHandle*>(hv);
    auto& X = h->w->x;
    for (std::size_t j=0; j<X.cols(); j++)
      for (std::size_t i=0; i<X.rows(); i++)
        if (set) X(i, j) = x[i + j*X.rows()];
        else x[i + j*X.rows()] = X(i, j);
  }

  /* FrontHSS::bwd_solve_node (FrontHSS.cpp:487-495) */
  void ref_hss_partial_backward(void* hv, int s, double* x0, int ldx) {
    auto* h = static_cast<HSSHandle*>(hv);
    auto n0 = h->H->child(0)->rows();
    DenseD X(n0, s);
    X.zero();
    h->H->child(0)->backward_solve(*h->w, X);
    for (int j=0; j<s; j++)
      for (std::size_t i=0; i<n0; i++) x0[i + std::size_t(j)*ldx] = X(i, j);
  }

  /* ---- BLR: compress_and_factor on a dense matrix, weak admissibility,
   *      tiles from ClusterTree(n).refine(leaf)   (test_BLR_seq.cpp:136-156) */
  void* ref_blr_factor_dense(int n, const double* A, int lda,
                             const char* blr_args, int* ntiles_out,
                             int* tiles_out /* may be null; >= n entries */) {
    DenseD Ad(n, n, A, lda);
    BLR::BLROptions<double> opts;
    opts.set_verbose(false);
    parse(opts, blr_args);
    structured::ClusterTree tree(n);
    tree.refine(opts.leaf_size());
    auto tiles = tree.template leaf_sizes<std::size_t>();
    int nt = tiles.size();
    DenseMatrix<bool> adm(nt, nt);
    adm.fill(true);
    for (int t=0; t<nt; t++) adm(t, t) = false;
    auto h = new BLRHandle;
    h->n = n;
    h->B.reset(new BLR::BLRMatrix<double>(n, tiles, n, tiles));
    h->B->compress_and_factor(Ad, adm, opts);
    if (ntiles_out) *ntiles_out = nt;
    if (tiles_out) for (int t=0; t<nt; t++) tiles_out[t] = int(tiles[t]);
    return h;
  }

  void ref_blr_solve(void* hv, int s, double* b, int ldb) {
    auto* h = static_cast<BLRHandle*>(hv);
    DenseW B(h->n, s, b, ldb);
    h->B->solve(B);
  }

  /* out[0..3] = rows, cols, rank, nonzeros */
  void ref_blr_info(void* hv, long long* out) {
    auto* h = static_cast<BLRHandle*>(hv);
    out[0] = h->B->rows(); out[1] = h->B->cols();
    out[2] = h->B->rank(); out[3] = h->B->nonzeros();
  }

  void ref_blr_destroy(void* hv) { delete static_cast<BLRHandle*>(hv); }

  /* ---- BLRMatrix::construct_and_partial_factor on the front [A11 A12; A21 A22]
   *      (BLRMatrix.cpp:739-1037; caller FrontBLR.cpp:429-433), weak
   *      admissibility, tiles from ClusterTree(n1/n2).refine(leaf).  A22 (n2 x n2,
   *      ld22) is overwritten with the Schur complement, as in the reference. */
  void* ref_blr_partial_factor(int n1, int n2, const double* A11, int ld11,
                               const double* A12, int ld12, const double* A21,
                               int ld21, double* A22, int ld22,
                               const char* blr_args) {
    DenseD a11(n1, n1, A11, ld11), a12(n1, n2, A12, ld12),
      a21(n2, n1, A21, ld21), a22(n2, n2, A22, ld22);
    BLR::BLROptions<double> opts;
    opts.set_verbose(false);
    parse(opts, blr_args);
    structured::ClusterTree t1(n1), t2(n2);
    t1.refine(opts.leaf_size());
    t2.refine(opts.leaf_size());
    auto tiles1 = t1.template leaf_sizes<std::size_t>();
    auto tiles2 = t2.template leaf_sizes<std::size_t>();
    int nt = tiles1.size();
    DenseMatrix<bool> adm(nt, nt);
    adm.fill(true);
    for (int t=0; t<nt; t++) adm(t, t) = false;
    auto h = new BLRHandle;
    h->n = n1 + n2; h->n1 = n1; h->n2 = n2;
    BLR::BLRMatrix<double>::construct_and_partial_factor
      (a11, a12, a21, a22, h->B11, h->B12, h->B21, tiles1, tiles2, adm, opts);
    for (int j=0; j<n2; j++)
      for (int i=0; i<n2; i++) A22[i + std::size_t(j)*ld22] = a22(i, j);
    return h;
  }

  /* out[0..2] = rank(F11), rank(F12), rank(F21) */
  void ref_blr_partial_info(void* hv, long long* out) {
    auto* h = static_cast<BLRHandle*>(hv);
    out[0] = h->B11.rank(); out[1] = h->B12.rank(); out[2] = h->B21.rank();
  }

  /* FrontBLR::fwd_solve_phase2 (FrontBLR.cpp:525-531) on b = [bloc; bupd] */
  void ref_blr_partial_forward(void* hv, int s, double* b, int ldb) {
    auto* h = static_cast<BLRHandle*>(hv);
    DenseW bloc(h->n1, s, b, ldb), bupd(h->n2, s, b + h->n1, ldb);
    bloc.laswp(h->B11.piv(), true);
    BLR::BLRMatrix<double>::trsmLNU_gemm(h->B11, h->B21, bloc, bupd, 0);
  }

  /* FrontBLR::bwd_solve_phase1 (FrontBLR.cpp:551-555) on y = [yloc; yupd] */
  void ref_blr_partial_backward(void* hv, int s, double* y, int ldy) {
    auto* h = static_cast<BLRHandle*>(hv);
    DenseW yloc(h->n1, s, y, ldy), yupd(h->n2, s, y + h->n1, ldy);
    BLR::BLRMatrix<double>::gemm_trsmUNN(h->B11, h->B12, yloc, yupd, 0);
  }
}
