// Exercises the C++ mirror of the reference's class surface the way the
// reference's own programs use it:
//   test/test_HSS_seq.cpp:93-250   HSSOptions from the command line, HSSMatrix(A, opts),
//                                  dense(), extract(I, J), get(i, j), factor/solve, shift
//   test/test_BLR_seq.cpp:120-196  BLROptions, compress_and_factor, solve
//   src/sparse/fronts/FrontHSS.cpp:385-410,452-495 and FrontBLR.cpp:429-433,525-568
//                                  partial factorizations + Schur complements of a front
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "strumpack_b200/StructuredMatrix.hpp"

using namespace strumpack;
using DenseM = DenseMatrix<double>;

static DenseM gemm(const DenseM& A, const DenseM& B, bool tA = false, bool tB = false) {
  const std::size_t m = tA ? A.cols() : A.rows(), k = tA ? A.rows() : A.cols(), n = tB ? B.rows() : B.cols();
  DenseM C(m, n);
  C.zero();
  for (std::size_t j = 0; j < n; j++)
    for (std::size_t l = 0; l < k; l++) {
      const double b = tB ? B(j, l) : B(l, j);
      for (std::size_t i = 0; i < m; i++) C(i, j) += (tA ? A(l, i) : A(i, l)) * b;
    }
  return C;
}
static double rel(DenseM A, const DenseM& B) { return A.sub(B).normF() / B.normF(); }
#define CHECK(cond, msg) do { if (!(cond)) { std::printf("ERROR: %s\n", msg); return 1; } } while (0)

int main(int argc, char* argv[]) {
  const int m = argc > 1 ? std::atoi(argv[1]) : 600;
  DenseM A(m, m);
  for (int j = 0; j < m; j++)
    for (int i = 0; i < m; i++) A(i, j) = (i == j) ? 3. : 1. / (1 + std::abs(i - j));

  // ---- HSS, as test_HSS_seq.cpp
  HSS::HSSOptions<double> hopts;
  CHECK(hopts.rel_tol() == 1e-2 && hopts.leaf_size() == 512 && hopts.d0() == 128, "HSSOptions defaults");
  const char* hargv[] = {"prog", "--hss_rel_tol", "1e-8", "--hss_leaf_size", "64", "--hss_d0", "96", nullptr};
  hopts.set_from_command_line(7, hargv);
  CHECK(hopts.rel_tol() == 1e-8 && hopts.leaf_size() == 64 && hopts.d0() == 96, "HSSOptions command line");
  HSS::HSSMatrix<double> H(A, hopts);
  std::printf("# H: %zu x %zu, rank %zu, levels %zu, %.3f MB\n", H.rows(), H.cols(), H.rank(), H.levels(),
              H.memory() / 1e6);
  CHECK(rel(H.dense(), A) < 1e2 * 1e-8, "compression error too big");
  std::vector<std::size_t> I = {0, 5, std::size_t(m / 2), std::size_t(m - 1)}, J = {1, std::size_t(m / 3), std::size_t(m - 2)};
  DenseM sub = H.extract(I, J);
  for (std::size_t a = 0; a < I.size(); a++)
    for (std::size_t b = 0; b < J.size(); b++)
      CHECK(std::abs(sub(a, b) - A(I[a], J[b])) < 1e-6, "extract");
  CHECK(std::abs(H.get(7, 300 % m) - A(7, 300 % m)) < 1e-6, "get");
  DenseM X(m, 2), B(m, 2);
  X.random();
  B = H.apply(X);
  DenseM C(B);
  HSS::apply_HSS(Trans::N, H, X, 2., C);        // C = H X + 2 C = 3 B
  for (int i = 0; i < m; i++) CHECK(std::abs(C(i, 0) - 3 * B(i, 0)) < 1e-10 * (1 + std::abs(B(i, 0))), "apply_HSS beta");
  H.factor();
  DenseM S(B);
  H.solve(S);
  CHECK(rel(S, X) < 1e-9, "ULV solve");
  HSS::WorkSolve<double> w;
  H.forward_solve(w, B);
  DenseM S2(m, 2);
  H.backward_solve(w, S2);
  CHECK(rel(S2, S) < 1e-13, "forward/backward solve");

  // ---- compress(Amult, Aelem, opts) through the block-extraction callback (FrontHSS.cpp:385)
  {
    HSS::HSSMatrix<double> He(m, m, hopts);
    int calls = 0;
    HSS::HSSMatrix<double>::elem_t Aelem = [&](const std::vector<std::size_t>& I, const std::vector<std::size_t>& J,
                                               DenseM& Bk) {
      calls++;
      for (std::size_t b = 0; b < J.size(); b++)
        for (std::size_t a = 0; a < I.size(); a++) Bk(a, b) = A(I[a], J[b]);
    };
    HSS::HSSMatrix<double>::mult_t Amult;      // not needed by the sampled ID
    He.compress(Amult, Aelem, hopts);
    CHECK(calls > 0 && He.rows() == std::size_t(m), "element callback");
    CHECK(rel(He.dense(), A) < 1e2 * 1e-8, "compress(Amult, Aelem) error too big");
  }

  // ---- the HSS front: partial_factor + Schur_update + Schur_product_direct
  H.partial_factor();
  DenseM Theta, DUB01, Phi;
  H.Schur_update(Theta, DUB01, Phi);
  DenseM Vhat = H.Vhat();
  const int n1 = int(Theta.rows()), n0 = m - n1;
  DenseM R(n1, 3), Sr, Sc;
  R.random();
  H.Schur_product_direct(Theta, DUB01, Phi, DenseM(), R, Sr, Sc);
  // S R = H11 R - Theta (Vhat^T (Phi^T R))
  DenseM A11(n1, n1, A.ptr(n0, n0), A.ld());
  DenseM ref = gemm(A11, R);
  ref.sub(gemm(Theta, gemm(Vhat, gemm(Phi, R, true, false), true, false)));
  CHECK(rel(Sr, ref) < 1e-6, "Schur_product_direct");
  DenseM b0(n0, 2);
  b0.random();
  DenseM red = H.child0_forward_solve(w, b0);
  DenseM x0(n0, 2);
  H.child0_backward_solve(w, x0);
  DenseM A00(n0, n0, A.ptr(0, 0), A.ld());
  CHECK(rel(gemm(A00, x0), b0) < 1e-6, "partial solve of block (0,0)");
  CHECK(red.rows() == Vhat.cols(), "reduced rhs shape");
  H.shift(1.5);
  H.factor();

  // ---- BLR, as test_BLR_seq.cpp
  BLR::BLROptions<double> bopts;
  CHECK(bopts.rel_tol() == 1e-4 && bopts.leaf_size() == 256 &&
        bopts.BLR_factor_algorithm() == BLR::BLRFactorAlgorithm::RL, "BLROptions defaults");
  const char* bargv[] = {"prog", "--blr_rel_tol", "1e-6", "--blr_leaf_size", "128", nullptr};
  bopts.set_from_command_line(5, bargv);
  BLR::BLRMatrix<double> Bm;
  Bm.compress_and_factor(A, bopts);
  DenseM Y(B);
  {
    DenseM AX = gemm(A, X);
    Bm.solve(AX);
    CHECK(rel(AX, X) < 1e2 * 1e-6, "BLR solve");
  }
  // ---- the BLR front: construct_and_partial_factor + half solves
  const int s1 = m / 2, s2 = m - s1;
  DenseM F11(s1, s1, A.ptr(0, 0), A.ld()), F12(s1, s2, A.ptr(0, s1), A.ld()), F21(s2, s1, A.ptr(s1, 0), A.ld()),
      F22(s2, s2, A.ptr(s1, s1), A.ld());
  auto F = BLR::BLRMatrix<double>::construct_and_partial_factor(F11, F12, F21, F22, bopts);
  CHECK(F.sep_rows() == std::size_t(s1) && F11.rows() == 0, "partial factor bookkeeping");
  DenseM rhs = gemm(A, X);
  F.trsmLNU_gemm(rhs);
  {   // dense solve of the Schur system by the engine itself (BLR with one tile = plain LU)
    BLR::BLROptions<double> dopts;
    dopts.set_leaf_size(s2 + 1);
    BLR::BLRMatrix<double> Sd;
    Sd.compress_and_factor(F22, dopts);
    DenseM yupd(s2, 2, rhs.ptr(s1, 0), rhs.ld());
    Sd.solve(yupd);
    for (int j = 0; j < 2; j++)
      for (int i = 0; i < s2; i++) rhs(s1 + i, j) = yupd(i, j);
  }
  F.gemm_trsmUNN(rhs);
  CHECK(rel(rhs, X) < 1e2 * 1e-6, "BLR front solve");
  std::printf("# exiting\n");
  return 0;
}
