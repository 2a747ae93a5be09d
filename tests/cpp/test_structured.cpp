// Mirrors the ULV check of the reference's test/test_HSS_seq.cpp:69-79,235-250
// (Toeplitz matrix, ||B - H (H\B)||_F / ||B||_F <= 1e-12) and the compression
// check (:143-152) through the C++ mirror of the reference interface.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>

#include "strumpack_b200/StructuredMatrix.hpp"

using namespace strumpack;

int main(int argc, char* argv[]) {
  int m = argc > 1 ? std::atoi(argv[1]) : 1000;
  DenseMatrix<double> A(m, m);
  for (int j = 0; j < m; j++)
    for (int i = 0; i < m; i++) A(i, j) = (i == j) ? 1. : 1. / (1 + std::abs(i - j));
  structured::StructuredOptions<double> opts;
  opts.set_type(structured::Type::HSS);
  opts.set_rel_tol(1e-6);
  opts.set_leaf_size(64);
  auto H = structured::construct_from_dense(A, opts);
  std::printf("# created H matrix of dimension %zu x %zu, rank %zu, memory %.3f MB\n", H->rows(),
              H->cols(), H->rank(), H->memory() / 1e6);
  DenseMatrix<double> X(m, 3), B(m, 3), R(m, 3);
  std::mt19937 g(1);
  std::normal_distribution<double> nd;
  for (int j = 0; j < 3; j++)
    for (int i = 0; i < m; i++) X(i, j) = nd(g);
  H->mult(Trans::N, X, B);
  // compression error against the dense product
  double err = 0, nrm = 0;
  for (int j = 0; j < 3; j++)
    for (int i = 0; i < m; i++) {
      double s = 0;
      for (int l = 0; l < m; l++) s += A(i, l) * X(l, j);
      err += (s - B(i, j)) * (s - B(i, j));
      nrm += s * s;
    }
  std::printf("# relative error = ||H*X-A*X||_F/||A*X||_F = %g\n", std::sqrt(err / nrm));
  if (std::sqrt(err / nrm) > 1e2 * 1e-6) { std::printf("ERROR: compression error too big!!\n"); return 1; }
  H->factor();
  for (int j = 0; j < 3; j++)
    for (int i = 0; i < m; i++) R(i, j) = B(i, j);
  H->solve(R);            // R = H \ B
  DenseMatrix<double> B2(m, 3);
  H->mult(Trans::N, R, B2);
  err = nrm = 0;
  for (int j = 0; j < 3; j++)
    for (int i = 0; i < m; i++) { err += (B(i, j) - B2(i, j)) * (B(i, j) - B2(i, j)); nrm += B(i, j) * B(i, j); }
  std::printf("# relative error = ||B-H*(H\\B)||_F/||B||_F = %g\n", std::sqrt(err / nrm));
  if (std::sqrt(err / nrm) > 1e-12) { std::printf("ERROR: ULV solve relative error too big!!\n"); return 1; }
  std::printf("# exiting\n");
  return 0;
}
