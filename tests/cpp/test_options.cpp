// CPU-only part of the C++ mirror: the option classes (defaults and command
// line parsing, reference StructuredOptions.hpp:106-162, HSSOptions.hpp:465-490,
// BLROptions.hpp:128-140) and the DenseMatrix helpers.  No GPU call is made.
#include <cstdio>
#include <cstring>

#include "strumpack_b200/StructuredMatrix.hpp"

using namespace strumpack;
#define CHECK(cond) do { if (!(cond)) { std::printf("FAILED: %s (line %d)\n", #cond, __LINE__); return 1; } } while (0)

int main() {
  structured::StructuredOptions<double> so;
  CHECK(so.type() == structured::Type::BLR && so.rel_tol() == 1e-4 && so.abs_tol() == 1e-10);
  CHECK(so.leaf_size() == 128 && so.max_rank() == 5000 && so.verbose());
  const char* a1[] = {"prog", "--structured_type", "HSS", "--structured_rel_tol", "1e-7", "--structured_leaf_size", "200",
                      "--structured_max_rank", "77", "--structured_abs_tol", "1e-13", "--structured_verbose", nullptr};
  so.set_from_command_line(12, a1);
  CHECK(so.type() == structured::Type::HSS && so.rel_tol() == 1e-7 && so.abs_tol() == 1e-13);
  CHECK(so.leaf_size() == 200 && so.max_rank() == 77 && so.verbose());
  CHECK(structured::get_name(structured::Type::HODLR) == "HODLR");

  HSS::HSSOptions<double> ho;
  CHECK(ho.type() == structured::Type::HSS && ho.rel_tol() == 1e-2 && ho.abs_tol() == 1e-8);
  CHECK(ho.leaf_size() == 512 && ho.max_rank() == 50000 && ho.d0() == 128 && ho.dd() == 64 && ho.p() == 10);
  CHECK(ho.compression_algorithm() == HSS::CompressionAlgorithm::STABLE);
  const char* a2[] = {"prog", "--hss_rel_tol", "1e-5", "--hss_dd", "32", "--hss_p", "5", "--unknown", "1", nullptr};
  ho.set_from_command_line(9, a2);
  CHECK(ho.rel_tol() == 1e-5 && ho.dd() == 32 && ho.p() == 5 && ho.leaf_size() == 512);
  HSS::HSSOptions<double> ho2(so);     // from StructuredOptions: keeps the tolerances, forces the type
  CHECK(ho2.type() == structured::Type::HSS && ho2.rel_tol() == 1e-7 && ho2.leaf_size() == 200);

  BLR::BLROptions<double> bo;
  CHECK(bo.type() == structured::Type::BLR && bo.rel_tol() == 1e-4 && bo.abs_tol() == 1e-12 && bo.leaf_size() == 256);
  CHECK(bo.low_rank_algorithm() == BLR::LowRankAlgorithm::RRQR && bo.admissibility() == BLR::Admissibility::WEAK);
  CHECK(bo.BLR_factor_algorithm() == BLR::BLRFactorAlgorithm::RL && int(BLR::BLRFactorAlgorithm::LL) == 2);
  const char* a3[] = {"prog", "--blr_factor_algorithm", "LL", "--blr_pivot_threshold", "1e-9", "--blr_leaf_size", "64", nullptr};
  bo.set_from_command_line(7, a3);
  CHECK(bo.BLR_factor_algorithm() == BLR::BLRFactorAlgorithm::LL && bo.pivot_threshold() == 1e-9 && bo.leaf_size() == 64);
  CHECK(BLR::get_name(BLR::BLRFactorAlgorithm::STAR) == "Star");

  DenseMatrix<double> A(3, 2);
  A.fill(2.);
  A(2, 1) = 5.;
  DenseMatrix<double> B(A), C;
  C = A;
  B.sub(A);
  CHECK(B.normF() == 0. && C(2, 1) == 5. && C.ld() == 3 && A.nonzeros() == 6);
  DenseMatrixWrapper<double> W(2, 1, A, 1, 1);     // view of A(1:3, 1)
  CHECK(W(1, 0) == 5. && W.ld() == 3);
  W(0, 0) = -1.;
  CHECK(A(1, 1) == -1.);
  DenseMatrix<double> M(std::move(C));
  CHECK(M.rows() == 3 && C.rows() == 0 && M(2, 1) == 5.);
  A.clear();
  CHECK(A.rows() == 0 && A.data() == nullptr);

  // laswp with LAPACK's 1-based pivot vector: forward applies the interchanges in order, backward undoes them
  // (FrontBLR.cpp:527, 566 call it around trsmLNU_gemm / gemm_trsmUNN)
  DenseMatrix<double> L(4, 2);
  for (int i = 0; i < 4; i++) { L(i, 0) = i; L(i, 1) = 10 + i; }
  const std::vector<int> piv = {3, 3, 4, 4};
  L.laswp(piv, true);
  CHECK(L(0, 0) == 2. && L(1, 0) == 0. && L(2, 0) == 3. && L(3, 0) == 1. && L(0, 1) == 12.);
  L.laswp(piv, false);
  for (int i = 0; i < 4; i++) CHECK(L(i, 0) == i && L(i, 1) == 10 + i);

  // the admissibility matrix type of the BLR constructors (adm_t = DenseMatrix<bool>, BLRMatrix.hpp:78)
  DenseMatrix<bool> adm(3, 3);
  adm.fill(true);
  for (int i = 0; i < 3; i++) adm(i, i) = false;
  CHECK(adm(0, 1) && !adm(2, 2) && adm.rows() == 3 && adm.ld() == 3);

  // ClusterTree: recursive bisection down to leaf_size (ClusterTree.hpp:104-114), leaves in order, pre-order export
  structured::ClusterTree tree(1000);
  tree.refine(128);
  const std::vector<int> leaves = tree.leaf_sizes();
  int sum = 0;
  for (int v : leaves) { sum += v; CHECK(v >= 128 && v < 256); }
  CHECK(sum == 1000 && leaves.size() == 4 && leaves[0] == 250);
  std::vector<int> sizes, nchild;
  tree.serialize(sizes, nchild);
  CHECK(sizes.size() == 7 && sizes[0] == 1000 && nchild[0] == 2 && sizes[1] == 500 && nchild[2] == 0);
  structured::ClusterTree flat(100);
  flat.refine(128);
  CHECK(flat.leaf_sizes().size() == 1 && flat.leaf_sizes()[0] == 100);
  std::printf("options ok\n");
  return 0;
}
