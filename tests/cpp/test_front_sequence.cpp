// The call sequence of the reference's HSS front against the mirror, statement by
// statement with the front's member names:
//   FrontHSS<T>::multifrontal_factorization   src/sparse/fronts/FrontHSS.cpp:371-412
//   FrontHSS<T>::fwd_solve_node               :445-470
//   FrontHSS<T>::bwd_solve_node               :478-496
// i.e. HSSMatrix(ClusterTree, opts), compress(mult, elem, opts), partial_factor,
// Schur_update, child(0)->ULV().Vhat(), gemm, child(0)->forward_solve(w, b, true)
// with w.reduced_rhs / w.x, child(0)->backward_solve(w, y), delete_trailing_block,
// reset.  What the parent front does with the Schur complement is a dense solve
// here; the result must be the solution of the whole front's system.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <sstream>
#include <vector>

#include "strumpack_b200/StructuredMatrix.hpp"

using namespace strumpack;
using scalar_t = double;
using DenseM_t = DenseMatrix<scalar_t>;
using DenseMW_t = DenseMatrixWrapper<scalar_t>;
#define CHECK(cond, msg) do { if (!(cond)) { std::printf("ERROR: %s (line %d)\n", msg, __LINE__); return 1; } } while (0)

static double rel(DenseM_t A, const DenseM_t& B) { return A.sub(B).normF() / B.normF(); }

// dense Gaussian elimination with partial pivoting (test plumbing)
static void dense_solve(DenseM_t A, DenseM_t& B) {
  const std::size_t n = A.rows();
  for (std::size_t j = 0; j < n; j++) {
    std::size_t p = j;
    for (std::size_t i = j + 1; i < n; i++) if (std::abs(A(i, j)) > std::abs(A(p, j))) p = i;
    for (std::size_t c = 0; c < n; c++) std::swap(A(j, c), A(p, c));
    for (std::size_t c = 0; c < B.cols(); c++) std::swap(B(j, c), B(p, c));
    for (std::size_t i = j + 1; i < n; i++) {
      const double l = A(i, j) / A(j, j);
      for (std::size_t c = j + 1; c < n; c++) A(i, c) -= l * A(j, c);
      for (std::size_t c = 0; c < B.cols(); c++) B(i, c) -= l * B(j, c);
    }
  }
  for (std::size_t jj = n; jj-- > 0;)
    for (std::size_t c = 0; c < B.cols(); c++) {
      B(jj, c) /= A(jj, jj);
      for (std::size_t i = 0; i < jj; i++) B(i, c) -= A(i, jj) * B(jj, c);
    }
}

int main(int argc, char* argv[]) {
  const int dim_sep_ = argc > 1 ? std::atoi(argv[1]) : 384, dim_upd_ = argc > 2 ? std::atoi(argv[2]) : 256;
  const int dim_blk = dim_sep_ + dim_upd_;
  auto dim_sep = [&] { return std::size_t(dim_sep_); };
  auto dim_upd = [&] { return std::size_t(dim_upd_); };
  const int sep_begin_ = 0, etree_level = 1, task_depth = 0;
  // the front [F11 F12; F21 F22] (nonsymmetric, diagonally dominant)
  DenseM_t F(dim_blk, dim_blk);
  for (int j = 0; j < dim_blk; j++)
    for (int i = 0; i < dim_blk; i++)
      F(i, j) = (i == j) ? 4. : (1. + 0.3 * (i > j)) / (1 + std::abs(i - j));

  HSS::HSSOptions<scalar_t> HSSopts;
  HSSopts.set_rel_tol(1e-9);
  HSSopts.set_abs_tol(1e-12);
  HSSopts.set_leaf_size(64);

  // FrontHSS constructor: the separator tree with the update part as the root's second child
  structured::ClusterTree sep_tree(dim_sep_), upd_tree(dim_upd_), tree(dim_blk);
  sep_tree.refine(HSSopts.leaf_size());
  upd_tree.refine(HSSopts.leaf_size());
  tree.c = {sep_tree, upd_tree};
  HSS::HSSMatrix<scalar_t> H_(tree, HSSopts);
  CHECK(H_.rows() == std::size_t(dim_blk) && H_.is_untouched() && !H_.is_compressed(), "HSSMatrix(ClusterTree, opts)");
  H_.set_openmp_task_depth(task_depth);

  // ---- multifrontal_factorization ------------------------------------------------------
  auto mult = [&](DenseM_t&, DenseM_t&, DenseM_t&, DenseM_t&) {};
  auto elem = [&](const std::vector<std::size_t>& I, const std::vector<std::size_t>& J, DenseM_t& B) {
    for (std::size_t j = 0; j < J.size(); j++)
      for (std::size_t i = 0; i < I.size(); i++) B(i, j) = F(I[i], J[j]);
  };
  DenseM_t Theta_, DUB01_, Phi_, ThetaVhatC_or_VhatCPhiC_;
  H_.compress(mult, elem, HSSopts);
  CHECK(H_.is_compressed() && rel(H_.dense(), F) < 1e-7, "compress(mult, elem, opts) on the given tree");
  CHECK(H_.child(0)->rows() == dim_sep() && H_.child(1)->rows() == dim_upd(), "the root splits at the separator");
  if (dim_sep()) {
    if (etree_level > 0) {
      H_.partial_factor();
      H_.Schur_update(Theta_, DUB01_, Phi_);
      const DenseM_t& Vhat = H_.child(0)->ULV().Vhat();
      if (Theta_.cols() < Phi_.cols()) {
        ThetaVhatC_or_VhatCPhiC_ = DenseM_t(Vhat.cols(), Phi_.rows());
        gemm(Trans::C, Trans::C, scalar_t(1.), Vhat, Phi_, scalar_t(0.), ThetaVhatC_or_VhatCPhiC_, task_depth);
      } else {
        ThetaVhatC_or_VhatCPhiC_ = DenseM_t(Theta_.rows(), Vhat.rows());
        gemm(Trans::N, Trans::C, scalar_t(1.), Theta_, Vhat, scalar_t(0.), ThetaVhatC_or_VhatCPhiC_, task_depth);
      }
    } else {
      H_.factor();
    }
  }
  // the Schur complement the parent front receives: S = F22 - Theta Vhat^H Phi^H   (extend-add by sampling
  // in the reference, Schur_product_direct; dense here)
  DenseM_t S(dim_upd_, dim_upd_);
  {
    DenseM_t R(dim_upd_, dim_upd_), Sr, Sc;
    R.zero();
    for (int i = 0; i < dim_upd_; i++) R(i, i) = 1.;
    H_.Schur_product_direct(Theta_, DUB01_, Phi_, ThetaVhatC_or_VhatCPhiC_, R, Sr, Sc);
    S = Sr;
  }
  H_.delete_trailing_block();

  // ---- fwd_solve_node ------------------------------------------------------------------
  const int nrhs = 3;
  DenseM_t b(dim_blk, nrhs), bref;
  b.random();
  bref = b;
  DenseM_t bupd(dim_upd_, nrhs);
  for (int c = 0; c < nrhs; c++)
    for (int i = 0; i < dim_upd_; i++) bupd(i, c) = b(dim_sep_ + i, c);
  std::unique_ptr<HSS::WorkSolve<scalar_t>> ULVwork_;
  if (etree_level) {
    if (Theta_.cols() && Phi_.cols()) {
      DenseMW_t bloc(dim_sep(), b.cols(), b, sep_begin_, 0);
      ULVwork_ = std::unique_ptr<HSS::WorkSolve<scalar_t>>(new HSS::WorkSolve<scalar_t>());
      H_.child(0)->forward_solve(*ULVwork_, bloc, true);
      if (dim_upd())
        gemm(Trans::N, Trans::N, scalar_t(-1.), Theta_, ULVwork_->reduced_rhs, scalar_t(1.), bupd, task_depth);
      ULVwork_->reduced_rhs.clear();
    }
  }
  // ---- the parent: solve with the Schur complement
  DenseM_t yupd(bupd);
  dense_solve(S, yupd);
  // ---- bwd_solve_node ------------------------------------------------------------------
  DenseM_t y(dim_blk, nrhs);
  y.zero();
  if (etree_level) {
    if (Phi_.cols() && Theta_.cols()) {
      if (dim_upd())
        gemm(Trans::C, Trans::N, scalar_t(-1.), Phi_, yupd, scalar_t(1.), ULVwork_->x, task_depth);
      DenseMW_t yloc(dim_sep(), y.cols(), y, sep_begin_, 0);
      H_.child(0)->backward_solve(*ULVwork_, yloc);
      ULVwork_.reset();
    }
  }
  for (int c = 0; c < nrhs; c++)
    for (int i = 0; i < dim_upd_; i++) y(dim_sep_ + i, c) = yupd(i, c);
  // the whole front's system, densely
  DenseM_t xref(bref);
  dense_solve(F, xref);
  std::printf("# front %d + %d: ||y - F^-1 b|| / ||F^-1 b|| = %.3e\n", dim_sep_, dim_upd_, rel(y, xref));
  CHECK(rel(y, xref) < 1e-6, "front solve through child(0) partial solves + Schur complement");

  // whole-matrix operations are refused after delete_trailing_block
  bool thrown = false;
  try { H_.apply(b); } catch (const std::logic_error&) { thrown = true; }
  CHECK(thrown, "apply after delete_trailing_block");
  // draw: one rectangle per leaf block and per off-diagonal block
  std::ostringstream os;
  H_.draw(os);
  CHECK(os.str().find("set obj rect") != std::string::npos && os.str().find("B01") != std::string::npos, "draw");
  // reset: back to an uncompressed matrix on the same partition; compress(A, opts) uses it again
  H_.reset();
  CHECK(H_.is_untouched() && H_.rows() == std::size_t(dim_blk), "reset");
  H_.compress(F, HSSopts);
  CHECK(H_.child(0)->rows() == dim_sep() && rel(H_.dense(), F) < 1e-7, "compress(A, opts) after reset keeps the partition");
  // compress_with_coordinates: 1-D coordinates = the index (same clustering, geometric sampling)
  {
    DenseM_t coords(1, dim_blk);
    for (int i = 0; i < dim_blk; i++) coords(0, i) = i;
    HSS::HSSMatrix<scalar_t> Hc(tree, HSSopts);
    Hc.compress_with_coordinates(coords, elem, HSSopts);
    CHECK(rel(Hc.dense(), F) < 1e-7, "compress_with_coordinates");
  }
  // ---- the BLR front, statement by statement as FrontBLR does it --------------------------
  //   FrontBLR::multifrontal_factorization   src/sparse/fronts/FrontBLR.cpp:422-433
  //   fwd_solve_phase2 / bwd_solve_phase1     :525-568
  {
    using BLRM_t = BLR::BLRMatrix<scalar_t>;
    BLR::BLROptions<scalar_t> blr_opts;
    blr_opts.set_rel_tol(1e-9);
    blr_opts.set_leaf_size(96);
    structured::ClusterTree st(dim_sep_), ut(dim_upd_);
    st.refine(blr_opts.leaf_size());
    ut.refine(blr_opts.leaf_size());
    auto leaves = [](const structured::ClusterTree& t) { auto v = t.leaf_sizes(); return std::vector<std::size_t>(v.begin(), v.end()); };
    const std::vector<std::size_t> sep_tiles_ = leaves(st), upd_tiles_ = leaves(ut);
    DenseMatrix<bool> admissibility_(sep_tiles_.size(), sep_tiles_.size());
    for (std::size_t j = 0; j < sep_tiles_.size(); j++)
      for (std::size_t i = 0; i < sep_tiles_.size(); i++) admissibility_(i, j) = i != j;   // weak admissibility
    DenseM_t F11(dim_sep_, dim_sep_), F12(dim_sep_, dim_upd_), F21(dim_upd_, dim_sep_), F22_(dim_upd_, dim_upd_);
    for (int j = 0; j < dim_blk; j++)
      for (int i = 0; i < dim_blk; i++) {
        const double v = F(i, j);
        if (i < dim_sep_ && j < dim_sep_) F11(i, j) = v;
        else if (i < dim_sep_) F12(i, j - dim_sep_) = v;
        else if (j < dim_sep_) F21(i - dim_sep_, j) = v;
        else F22_(i - dim_sep_, j - dim_sep_) = v;
      }
    BLRM_t F11blr_, F12blr_, F21blr_;
    {
      auto nF11 = F11.normF();
      auto nF12 = F12.normF();
      auto nF21 = F21.normF();
      auto nF = std::sqrt(nF11*nF11 + nF12*nF12 + nF21*nF21);
      auto lopts = blr_opts;
      lopts.set_abs_tol(lopts.abs_tol() * nF);
      BLRM_t::construct_and_partial_factor
        (F11, F12, F21, F22_, F11blr_, F12blr_, F21blr_,
         sep_tiles_, upd_tiles_, admissibility_, lopts);
    }
    CHECK(F11blr_.rowblocks() == sep_tiles_.size() + upd_tiles_.size(), "the engine used the caller's tiles");
    // F22_ now holds the Schur complement; solve the whole front with the two half solves
    DenseM_t b(dim_blk, nrhs), y(dim_blk, nrhs);
    b.random();
    DenseM_t bcopy(b);
    DenseM_t bupd(dim_upd_, nrhs);
    for (int c = 0; c < nrhs; c++)
      for (int i = 0; i < dim_upd_; i++) bupd(i, c) = b(dim_sep_ + i, c);
    {   // fwd_solve_phase2
      DenseMW_t bloc(dim_sep(), b.cols(), b, sep_begin_, 0);
      bloc.laswp(F11blr_.piv(), true);
      BLRM_t::trsmLNU_gemm(F11blr_, F21blr_, bloc, bupd, task_depth);
    }
    DenseM_t yupd(bupd);
    dense_solve(F22_, yupd);         // the parent front
    for (int c = 0; c < nrhs; c++)
      for (int i = 0; i < dim_sep_; i++) y(i, c) = b(i, c);
    {   // bwd_solve_phase1
      DenseMW_t yloc(dim_sep(), y.cols(), y, sep_begin_, 0);
      BLRM_t::gemm_trsmUNN(F11blr_, F12blr_, yloc, yupd, task_depth);
    }
    for (int c = 0; c < nrhs; c++)
      for (int i = 0; i < dim_upd_; i++) y(dim_sep_ + i, c) = yupd(i, c);
    DenseM_t xref(bcopy);
    dense_solve(F, xref);
    std::printf("# BLR front %d + %d: ||y - F^-1 b|| / ||F^-1 b|| = %.3e\n", dim_sep_, dim_upd_, rel(y, xref));
    CHECK(rel(y, xref) < 1e-6, "BLR front solve through the reference's static calls");
  }
  std::printf("# exiting\n");
  return 0;
}
