// strumpack_b200 -- C++ host-side mirror of the reference's structured-matrix
// interface, header-only on top of the C ABI (include/sb200_structured.h).
//
// Same class / function names, argument meaning and error behaviour as
//   strumpack::structured::StructuredMatrix<T>   reference src/structured/StructuredMatrix.hpp:209-418
//   strumpack::structured::construct_from_dense  reference src/structured/StructuredMatrix.cpp:53-127
//   strumpack::structured::StructuredOptions<T>  reference src/structured/StructuredOptions.hpp:106-162
//   strumpack::HSS::HSSOptions<T>                reference src/HSS/HSSOptions.hpp:151-491
//   strumpack::HSS::HSSMatrix<T>                 reference src/HSS/HSSMatrix.hpp:95-511, free apply_HSS :705-713
//   strumpack::BLR::BLROptions<T>                reference src/BLR/BLROptions.hpp:81-142
//   strumpack::BLR::BLRMatrix<T>                 reference src/BLR/BLRMatrix.hpp:68-291
//   strumpack::DenseMatrix<T> / DenseMatrixWrapper<T>  reference src/dense/DenseMatrix.hpp:139-146
// so that code written against the reference (examples/dense/testStructured.cpp,
// test/test_HSS_seq.cpp, test/test_BLR_seq.cpp, the calls FrontHSS / FrontBLR make)
// compiles against this header with only the include changed.  Everything
// numeric happens in libstrumpack_b200.so on the GPU; there is no CPU path
// here (a call without a CUDA device throws).  Only T = double is implemented.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "../sb200_structured.h"

namespace strumpack {

enum class Trans : char { N = 'N', C = 'C', T = 'T' };

// column-major owning matrix (reference DenseMatrix: data_, rows_, cols_, ld_)
template <typename scalar_t> class DenseMatrix {
 public:
  DenseMatrix() = default;
  DenseMatrix(std::size_t m, std::size_t n) : rows_(m), cols_(n), ld_(m ? m : 1), own_(m * n) {
    data_ = reinterpret_cast<scalar_t*>(own_.data());
  }
  // copy of a column-major block (reference DenseMatrix(m, n, D, ld))
  DenseMatrix(std::size_t m, std::size_t n, const scalar_t* D, std::size_t ld) : DenseMatrix(m, n) {
    for (std::size_t j = 0; j < n; j++)
      for (std::size_t i = 0; i < m; i++) data_[i + j * ld_] = D[i + j * ld];
  }
  DenseMatrix(const DenseMatrix& o) : DenseMatrix(o.rows_, o.cols_, o.data_, o.ld_) {}
  DenseMatrix(DenseMatrix&& o) noexcept { swap_in(std::move(o)); }
  DenseMatrix& operator=(const DenseMatrix& o) {
    if (this != &o) { DenseMatrix t(o); swap_in(std::move(t)); }
    return *this;
  }
  DenseMatrix& operator=(DenseMatrix&& o) noexcept {
    if (this != &o) swap_in(std::move(o));
    return *this;
  }
  virtual ~DenseMatrix() = default;
  std::size_t rows() const { return rows_; }
  std::size_t cols() const { return cols_; }
  std::size_t ld() const { return ld_; }
  scalar_t* data() { return data_; }
  const scalar_t* data() const { return data_; }
  scalar_t* ptr(std::size_t i, std::size_t j) { return data_ + i + j * ld_; }
  const scalar_t* ptr(std::size_t i, std::size_t j) const { return data_ + i + j * ld_; }
  scalar_t& operator()(std::size_t i, std::size_t j) { return data_[i + j * ld_]; }
  const scalar_t& operator()(std::size_t i, std::size_t j) const { return data_[i + j * ld_]; }
  void fill(scalar_t v) {
    for (std::size_t j = 0; j < cols_; j++)
      for (std::size_t i = 0; i < rows_; i++) data_[i + j * ld_] = v;
  }
  void zero() { fill(scalar_t(0)); }
  void eye() {
    zero();
    for (std::size_t i = 0; i < std::min(rows_, cols_); i++) data_[i + i * ld_] = scalar_t(1);
  }
  void random() {   // reference DenseMatrix::random(): standard normal entries
    std::mt19937 g(0);
    std::normal_distribution<double> d;
    for (std::size_t j = 0; j < cols_; j++)
      for (std::size_t i = 0; i < rows_; i++) data_[i + j * ld_] = scalar_t(d(g));
  }
  double normF() const {
    double s = 0;
    for (std::size_t j = 0; j < cols_; j++)
      for (std::size_t i = 0; i < rows_; i++) s += double(data_[i + j * ld_]) * double(data_[i + j * ld_]);
    return std::sqrt(s);
  }
  double norm() const { return normF(); }
  // this <- this - B     (reference DenseMatrix::sub)
  DenseMatrix& sub(const DenseMatrix& B) {
    for (std::size_t j = 0; j < cols_; j++)
      for (std::size_t i = 0; i < rows_; i++) data_[i + j * ld_] -= B(i, j);
    return *this;
  }
  DenseMatrix& add(const DenseMatrix& B) {
    for (std::size_t j = 0; j < cols_; j++)
      for (std::size_t i = 0; i < rows_; i++) data_[i + j * ld_] += B(i, j);
    return *this;
  }
  std::size_t nonzeros() const { return rows_ * cols_; }
  void clear() { own_.clear(); own_.shrink_to_fit(); data_ = nullptr; rows_ = cols_ = 0; ld_ = 1; }
  // DenseMatrix::laswp(P, fwd) (reference DenseMatrix.cpp:288-297): LAPACK row
  // interchanges, P 1-based.  An EMPTY P is a no-op: BLRMatrix::piv() of this
  // engine is empty because the interchanges are applied inside its solves.
  void laswp(const std::vector<int>& P, bool fwd) {
    if (P.empty()) return;
    const int n = int(P.size());
    for (int q = 0; q < n; q++) {
      const int i = fwd ? q : n - 1 - q, p = P[i] - 1;
      if (p != i)
        for (std::size_t c = 0; c < cols_; c++) std::swap((*this)(i, c), (*this)(p, c));
    }
  }

 protected:
  void swap_in(DenseMatrix&& o) {
    const bool owned = !o.own_.empty() || o.data_ == nullptr;
    own_ = std::move(o.own_);
    rows_ = o.rows_; cols_ = o.cols_; ld_ = o.ld_;
    data_ = owned ? reinterpret_cast<scalar_t*>(own_.data()) : o.data_;
    o.data_ = nullptr; o.rows_ = o.cols_ = 0; o.ld_ = 1;
  }
  scalar_t* data_ = nullptr;
  std::size_t rows_ = 0, cols_ = 0, ld_ = 1;
  // (bool is stored as bytes: std::vector<bool> has no contiguous storage; adm_t = DenseMatrix<bool>)
  using store_t = typename std::conditional<std::is_same<scalar_t, bool>::value, unsigned char, scalar_t>::type;
  std::vector<store_t> own_;
};

// non-owning view (reference DenseMatrixWrapper)
template <typename scalar_t> class DenseMatrixWrapper : public DenseMatrix<scalar_t> {
 public:
  DenseMatrixWrapper(std::size_t m, std::size_t n, scalar_t* D, std::size_t ld) {
    this->data_ = D; this->rows_ = m; this->cols_ = n; this->ld_ = ld ? ld : 1;
  }
  // sub-block view (reference DenseMatrixWrapper(m, n, D, i, j))
  DenseMatrixWrapper(std::size_t m, std::size_t n, DenseMatrix<scalar_t>& D, std::size_t i, std::size_t j)
      : DenseMatrixWrapper(m, n, D.ptr(i, j), D.ld()) {}
};

namespace detail {
inline void check(int rc, const char* what) {
  if (rc) throw std::logic_error(std::string("strumpack_b200: ") + what + " failed");
}
// "--prefix_name value" command-line scanning shared by the option classes
// (the reference uses getopt_long, e.g. StructuredOptions.cpp:62-140; unknown
// options are ignored there as well)
inline const char* arg_value(int argc, const char* const* argv, const std::string& name) {
  for (int i = 1; i + 1 < argc; i++)
    if (argv[i] && name == argv[i]) return argv[i + 1];
  return nullptr;
}
inline bool arg_flag(int argc, const char* const* argv, const std::string& name) {
  for (int i = 1; i < argc; i++)
    if (argv[i] && name == argv[i]) return true;
  return false;
}
}  // namespace detail

namespace structured {

// reference StructuredOptions.hpp:49-74
enum class Type : int {
  HSS = SP_TYPE_HSS, BLR = SP_TYPE_BLR, HODLR = SP_TYPE_HODLR, HODBF = SP_TYPE_HODBF,
  BUTTERFLY = SP_TYPE_BUTTERFLY, LR = SP_TYPE_LR, LOSSY = SP_TYPE_LOSSY, LOSSLESS = SP_TYPE_LOSSLESS
};
inline std::string get_name(Type a) {
  switch (a) {
    case Type::HSS: return "HSS";
    case Type::BLR: return "BLR";
    case Type::HODLR: return "HODLR";
    case Type::HODBF: return "HODBF";
    case Type::BUTTERFLY: return "BUTTERFLY";
    case Type::LR: return "LR";
    case Type::LOSSY: return "LOSSY";
    case Type::LOSSLESS: return "LOSSLESS";
  }
  return "UNKNOWN";
}

// reference StructuredOptions.hpp:106-162 (same defaults: BLR, 1e-4, 1e-10, 128, 5000)
template <typename scalar_t> class StructuredOptions {
 public:
  using real_t = double;
  StructuredOptions() { SP_d_struct_default_options(&o_); }
  explicit StructuredOptions(Type t) : StructuredOptions() { set_type(t); }
  virtual ~StructuredOptions() = default;
  void set_type(Type t) { o_.type = static_cast<SP_STRUCTURED_TYPE>(t); }
  void set_rel_tol(double t) { o_.rel_tol = t; }
  void set_abs_tol(double t) { o_.abs_tol = t; }
  void set_leaf_size(int s) { o_.leaf_size = s; }
  void set_max_rank(int r) { o_.max_rank = r; }
  void set_verbose(bool v) { o_.verbose = v; }
  Type type() const { return static_cast<Type>(o_.type); }
  double rel_tol() const { return o_.rel_tol; }
  double abs_tol() const { return o_.abs_tol; }
  int leaf_size() const { return o_.leaf_size; }
  int max_rank() const { return o_.max_rank; }
  bool verbose() const { return o_.verbose; }
  const CSPOptions* c() const { return &o_; }
  // --structured_{rel_tol,abs_tol,leaf_size,max_rank,type,verbose,quiet}
  // (reference StructuredOptions.cpp:62-140)
  virtual void set_from_command_line(int argc, const char* const* argv) { scan(argc, argv, "--structured_"); }
  virtual void describe_options() const {
    std::cout << "# Structured Options:\n#   --structured_rel_tol real (default " << rel_tol()
              << ")\n#   --structured_abs_tol real (default " << abs_tol()
              << ")\n#   --structured_leaf_size int (default " << leaf_size()
              << ")\n#   --structured_max_rank int (default " << max_rank()
              << ")\n#   --structured_type [HSS|BLR] (default " << get_name(type())
              << ")\n#   --structured_verbose or -v / --structured_quiet or -q\n";
  }

 protected:
  void scan(int argc, const char* const* argv, const std::string& pre) {
    if (auto v = detail::arg_value(argc, argv, pre + "rel_tol")) set_rel_tol(std::atof(v));
    if (auto v = detail::arg_value(argc, argv, pre + "abs_tol")) set_abs_tol(std::atof(v));
    if (auto v = detail::arg_value(argc, argv, pre + "leaf_size")) set_leaf_size(std::atoi(v));
    if (auto v = detail::arg_value(argc, argv, pre + "max_rank")) set_max_rank(std::atoi(v));
    if (auto v = detail::arg_value(argc, argv, pre + "type")) {
      const std::string s(v);
      if (s == "HSS") set_type(Type::HSS);
      else if (s == "BLR") set_type(Type::BLR);
      else std::cerr << "# WARNING: structured type " << s << " not supported by strumpack_b200" << std::endl;
    }
    if (detail::arg_flag(argc, argv, pre + "verbose")) set_verbose(true);
    if (detail::arg_flag(argc, argv, pre + "quiet")) set_verbose(false);
  }
  CSPOptions o_;
};

// reference StructuredMatrix.hpp:209-418: unsupported operations throw
template <typename scalar_t> class StructuredMatrix {
  static_assert(sizeof(scalar_t) == sizeof(double), "strumpack_b200 implements double only");

 public:
  StructuredMatrix() = default;
  explicit StructuredMatrix(CSPStructMat h) : h_(h) {}
  StructuredMatrix(const StructuredMatrix&) = delete;
  StructuredMatrix& operator=(const StructuredMatrix&) = delete;
  StructuredMatrix(StructuredMatrix&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
  StructuredMatrix& operator=(StructuredMatrix&& o) noexcept {
    if (this != &o) { SP_d_struct_destroy(&h_); h_ = o.h_; o.h_ = nullptr; }
    return *this;
  }
  virtual ~StructuredMatrix() { SP_d_struct_destroy(&h_); }

  virtual std::size_t rows() const { return SP_d_struct_rows(h_); }
  virtual std::size_t cols() const { return SP_d_struct_cols(h_); }
  virtual std::size_t memory() const { return SP_d_struct_memory(h_); }
  virtual std::size_t nonzeros() const { return SP_d_struct_nonzeros(h_); }
  virtual std::size_t rank() const { return SP_d_struct_rank(h_); }

  // y = op(A) x                                  (StructuredMatrix.hpp:280-300)
  virtual void mult(Trans op, const DenseMatrix<scalar_t>& x, DenseMatrix<scalar_t>& y) const {
    mult(op, int(x.cols()), x.data(), int(x.ld()), y.data(), int(y.ld()));
  }
  virtual void mult(Trans op, int m, const scalar_t* x, int ldx, scalar_t* y, int ldy) const {
    detail::check(SP_d_struct_mult(h_, char(op), m, x, ldx, y, ldy), "mult");
  }
  virtual void factor() { detail::check(SP_d_struct_factor(h_), "factor"); }
  // b <- A^{-1} b                                (StructuredMatrix.hpp:340-360)
  virtual void solve(DenseMatrix<scalar_t>& b) const { solve(int(b.cols()), b.data(), int(b.ld())); }
  virtual void solve(int nrhs, scalar_t* b, int ldb) const {
    detail::check(SP_d_struct_solve(h_, nrhs, b, ldb), "solve");
  }
  virtual void shift(scalar_t s) { detail::check(SP_d_struct_shift(h_, s), "shift"); }
  CSPStructMat handle() const { return h_; }

 protected:
  static void check(int rc, const char* what) { detail::check(rc, what); }
  CSPStructMat h_ = nullptr;
};

// reference StructuredMatrix.cpp:53-127
template <typename scalar_t>
std::unique_ptr<StructuredMatrix<scalar_t>> construct_from_dense(
    int rows, int cols, const scalar_t* A, int ldA, const StructuredOptions<scalar_t>& opts) {
  CSPStructMat h = nullptr;
  if (SP_d_struct_from_dense(&h, rows, cols, A, ldA, opts.c()))
    throw std::invalid_argument("construct_from_dense failed");
  return std::unique_ptr<StructuredMatrix<scalar_t>>(new StructuredMatrix<scalar_t>(h));
}
template <typename scalar_t>
std::unique_ptr<StructuredMatrix<scalar_t>> construct_from_dense(
    const DenseMatrix<scalar_t>& A, const StructuredOptions<scalar_t>& opts) {
  return construct_from_dense(int(A.rows()), int(A.cols()), A.data(), int(A.ld()), opts);
}
// reference StructuredMatrix.cpp:193-312: the callback is a plain function
// pointer at the C boundary (StructuredMatrix.h:244)
template <typename scalar_t>
std::unique_ptr<StructuredMatrix<scalar_t>> construct_from_elements(
    int rows, int cols, scalar_t (*A)(int, int), const StructuredOptions<scalar_t>& opts) {
  CSPStructMat h = nullptr;
  if (SP_d_struct_from_elements(&h, rows, cols, A, opts.c()))
    throw std::invalid_argument("construct_from_elements failed");
  return std::unique_ptr<StructuredMatrix<scalar_t>>(new StructuredMatrix<scalar_t>(h));
}

}  // namespace structured

// C = alpha op(A) op(B) + beta C on host matrices: the small dense products the
// fronts do around the structured calls (reference DenseMatrix.hpp free gemm,
// src/dense/DenseMatrix.cpp:936-...).  Plain loops: glue, not a kernel.
template <typename scalar_t>
void gemm(Trans ta, Trans tb, scalar_t alpha, const DenseMatrix<scalar_t>& A, const DenseMatrix<scalar_t>& B,
          scalar_t beta, DenseMatrix<scalar_t>& C, int /*task_depth*/ = 0) {
  const bool tA = ta != Trans::N, tB = tb != Trans::N;
  const std::size_t M = C.rows(), N = C.cols(), K = tA ? A.rows() : A.cols();
  if ((tA ? A.cols() : A.rows()) != M || (tB ? B.rows() : B.cols()) != N || (tB ? B.cols() : B.rows()) != K)
    throw std::invalid_argument("gemm: dimension mismatch");
  for (std::size_t j = 0; j < N; j++)
    for (std::size_t i = 0; i < M; i++) {
      scalar_t acc = 0;
      for (std::size_t k = 0; k < K; k++) acc += (tA ? A(k, i) : A(i, k)) * (tB ? B(j, k) : B(k, j));
      C(i, j) = alpha * acc + (beta == scalar_t(0) ? scalar_t(0) : beta * C(i, j));
    }
}

namespace structured {
// reference structured::ClusterTree (src/structured/ClusterTree.hpp:50-170): a
// recursive partition of an index range
class ClusterTree {
 public:
  int size = 0;
  std::vector<ClusterTree> c;
  ClusterTree() = default;
  explicit ClusterTree(int n) : size(n) {}
  // recursive bisection down to leaf_size                      (ClusterTree.hpp:104-114)
  const ClusterTree& refine(int leaf_size) {
    if (c.empty()) {
      if (size >= 2 * leaf_size) {
        c.resize(2);
        c[0].size = size / 2;
        c[1].size = size - size / 2;
        c[0].refine(leaf_size);
        c[1].refine(leaf_size);
      }
    } else
      for (auto& ch : c) ch.refine(leaf_size);
    return *this;
  }
  int levels() const {
    int l = 0;
    for (auto& ch : c) l = std::max(l, ch.levels());
    return l + 1;
  }
  std::vector<int> leaf_sizes() const {
    std::vector<int> out;
    collect_leaves(out);
    return out;
  }
  // pre-order (size, number of children) arrays: the form the C ABI takes
  void serialize(std::vector<int>& sizes, std::vector<int>& nchild) const {
    sizes.push_back(size);
    nchild.push_back(int(c.size()));
    for (auto& ch : c) ch.serialize(sizes, nchild);
  }
 private:
  void collect_leaves(std::vector<int>& out) const {
    if (c.empty()) out.push_back(size);
    else for (auto& ch : c) ch.collect_leaves(out);
  }
};
}  // namespace structured

namespace HSS {

// reference HSSOptions.hpp:59-148
enum class CompressionAlgorithm { ORIGINAL, STABLE, HARD_RESTART };
enum class CompressionSketch { GAUSSIAN, SJLT };
enum class ClusteringAlgorithm { NATURAL, TWO_MEANS, KD_TREE, PCA, COBBLE };

// reference HSSOptions.hpp:151-491, defaults :465-490 (rel 1e-2, abs 1e-8, leaf
// 512, max_rank 50000, d0 128, dd 64, p 10).  The engine's GPU compressor is a
// sampled interpolative decomposition: it honours rel_tol,
// abs_tol, leaf_size and max_rank; the sampling parameters d0/dd/p and the
// algorithm enums are kept so that option-setting code compiles unchanged.
template <typename scalar_t> class HSSOptions : public structured::StructuredOptions<scalar_t> {
 public:
  HSSOptions() : structured::StructuredOptions<scalar_t>(structured::Type::HSS) {
    this->set_rel_tol(1e-2); this->set_abs_tol(1e-8);
    this->set_leaf_size(512); this->set_max_rank(50000);
  }
  HSSOptions(const structured::StructuredOptions<scalar_t>& s) : structured::StructuredOptions<scalar_t>(s) {
    this->set_type(structured::Type::HSS);
  }
  void set_d0(int v) { d0_ = v; }
  void set_dd(int v) { dd_ = v; }
  void set_p(int v) { p_ = v; }
  void set_compression_algorithm(CompressionAlgorithm a) { alg_ = a; }
  void set_compression_sketch(CompressionSketch a) { sketch_ = a; }
  void set_clustering_algorithm(ClusteringAlgorithm a) { clus_ = a; }
  void set_approximate_neighbors(int v) { ann_ = v; }
  void set_ann_iterations(int v) { ann_it_ = v; }
  void set_user_defined_random(bool v) { user_rand_ = v; }
  void set_synchronized_compression(bool v) { sync_ = v; }
  void set_log_ranks(bool v) { log_ranks_ = v; }
  int d0() const { return d0_; }
  int dd() const { return dd_; }
  int p() const { return p_; }
  CompressionAlgorithm compression_algorithm() const { return alg_; }
  CompressionSketch compression_sketch() const { return sketch_; }
  ClusteringAlgorithm clustering_algorithm() const { return clus_; }
  int approximate_neighbors() const { return ann_; }
  int ann_iterations() const { return ann_it_; }
  bool user_defined_random() const { return user_rand_; }
  bool synchronized_compression() const { return sync_; }
  bool log_ranks() const { return log_ranks_; }
  // --hss_{rel_tol,abs_tol,leaf_size,max_rank,d0,dd,p,verbose,quiet}  (HSSOptions.cpp:67-240)
  void set_from_command_line(int argc, const char* const* argv) override {
    this->scan(argc, argv, "--hss_");
    if (auto v = detail::arg_value(argc, argv, "--hss_d0")) set_d0(std::atoi(v));
    if (auto v = detail::arg_value(argc, argv, "--hss_dd")) set_dd(std::atoi(v));
    if (auto v = detail::arg_value(argc, argv, "--hss_p")) set_p(std::atoi(v));
  }

 private:
  int d0_ = 128, dd_ = 64, p_ = 10, ann_ = 64, ann_it_ = 5;
  CompressionAlgorithm alg_ = CompressionAlgorithm::STABLE;
  CompressionSketch sketch_ = CompressionSketch::GAUSSIAN;
  ClusteringAlgorithm clus_ = ClusteringAlgorithm::TWO_MEANS;
  bool user_rand_ = false, sync_ = true, log_ranks_ = false;
};

// state handed from forward_solve to backward_solve (reference WorkSolve,
// HSSExtra.hpp:216-226); here the intermediate vectors stay on the device
// inside the matrix object, the struct only remembers the shape
template <typename scalar_t> struct WorkSolve {
  // what FrontHSS reads and writes between the two partial solves
  // (FrontHSS.cpp:452-462, 487-495): V-hat^H x of the eliminated block, and the
  // reduced unknowns of child 0 (updated by the caller before backward_solve)
  DenseMatrix<scalar_t> reduced_rhs, x;
  int nrhs = 0;
  bool partial = false;
};

// reference HSSFactors<T> (HSSExtra.hpp:161-213): the part of it callers use
template <typename scalar_t> class HSSFactors {
 public:
  const DenseMatrix<scalar_t>& Vhat() const { return Vhat_; }
  DenseMatrix<scalar_t>& Vhat() { return Vhat_; }
 private:
  DenseMatrix<scalar_t> Vhat_;
  template <typename T> friend class HSSMatrix;
};

// reference HSSMatrix<T> (src/HSS/HSSMatrix.hpp:95-511)
template <typename scalar_t> class HSSMatrix : public structured::StructuredMatrix<scalar_t> {
  using base = structured::StructuredMatrix<scalar_t>;
  using DenseM_t = DenseMatrix<scalar_t>;

 public:
  using opts_t = HSSOptions<scalar_t>;
  HSSMatrix() = default;
  explicit HSSMatrix(CSPStructMat h) : base(h) {}
  // HSSMatrix(const DenseM_t& A, const opts_t& opts)          (HSSMatrix.cpp:49-54)
  HSSMatrix(const DenseM_t& A, const opts_t& opts) { compress(A, opts); }
  // HSSMatrix(m, n, opts): an uncompressed m x n matrix, to be filled by compress(...)
  //                                                             (HSSMatrix.cpp:56-58)
  HSSMatrix(std::size_t m, std::size_t n, const opts_t& /*opts*/) : m0_(m), n0_(n) {}
  // HSSMatrix(const structured::ClusterTree& t, opts): uncompressed, the partition is t
  //                                                             (HSSMatrix.cpp:71-82)
  HSSMatrix(const structured::ClusterTree& t, const opts_t& /*opts*/) : m0_(t.size), n0_(t.size) {
    t.serialize(tree_sizes_, tree_nchild_);
  }
  std::size_t rows() const override { return this->h_ ? base::rows() : m0_; }
  std::size_t cols() const override { return this->h_ ? base::cols() : n0_; }
  HSSMatrix(HSSMatrix&& o) noexcept = default;
  HSSMatrix& operator=(HSSMatrix&& o) noexcept = default;
  // HSSMatrix(kernel::Kernel&, opts) (HSSMatrix.cpp:88-106): the d x n points
  // are reordered in place into the cluster ordering; perm (0-based, may be
  // null) receives the permutation.
  static HSSMatrix from_kernel(int n, int d, scalar_t* pts, SB200_KERNEL_TYPE kernel, double h,
                               double lambda, const opts_t& opts, int* perm = nullptr) {
    CSPStructMat s = nullptr;
    // opts.clustering_algorithm() is honoured (TWO_MEANS by default, as in the reference);
    // PCA / COBBLE are refused by the library
    if (SB200_d_hss_from_kernel_ex(&s, n, d, pts, kernel, h, lambda, opts.c(), perm, int(opts.clustering_algorithm())))
      throw std::invalid_argument("HSSMatrix(kernel) failed");
    return HSSMatrix(s);
  }
  // HSSMatrix::compress(A, opts)                               (HSSMatrix.cpp:158-171)
  void compress(const DenseM_t& A, const opts_t& opts) {
    opts_t o(opts);
    CSPStructMat s = nullptr;
    const int rc = tree_sizes_.empty()
        ? SP_d_struct_from_dense(&s, int(A.rows()), int(A.cols()), A.data(), int(A.ld()), o.c())
        : SB200_d_hss_from_dense_tree(&s, int(A.rows()), A.data(), int(A.ld()), o.c(), int(tree_sizes_.size()),
                                      tree_sizes_.data(), tree_nchild_.data());
    if (rc) throw std::invalid_argument("HSSMatrix::compress failed");
    adopt(s);
  }
  // HSSMatrix::compress(Amult, Aelem, opts) (HSSMatrix.cpp:173-186).  The
  // engine's sampled interpolative decomposition needs element access only:
  // Amult is accepted for source compatibility and not called.
  using elem_t = std::function<void(const std::vector<std::size_t>& I, const std::vector<std::size_t>& J, DenseM_t& B)>;
  using mult_t = std::function<void(DenseM_t& Rr, DenseM_t& Rc, DenseM_t& Sr, DenseM_t& Sc)>;
  void compress(const mult_t& Amult, const elem_t& Aelem, const opts_t& opts) {
    if (rows() != cols() || rows() == 0) throw std::invalid_argument("HSSMatrix::compress: square matrices only");
    compress(Amult, Aelem, rows(), opts);
  }
  void compress(const mult_t& /*Amult*/, const elem_t& Aelem, std::size_t n, const opts_t& opts) {
    compress_elements(Aelem, n, opts, 0, nullptr);
  }
  // HSSMatrix::compress_with_coordinates(coords, Aelem, opts) (HSSMatrix.hpp:302):
  // coords is d x n, one point per column; the sampled columns of every node are
  // its geometric neighbours
  void compress_with_coordinates(const DenseM_t& coords, const elem_t& Aelem, const opts_t& opts) {
    if (coords.cols() != rows()) throw std::invalid_argument("compress_with_coordinates: one point per row of the matrix");
    DenseM_t packed(coords.rows(), coords.cols());
    for (std::size_t j = 0; j < coords.cols(); j++)
      for (std::size_t i = 0; i < coords.rows(); i++) packed(i, j) = coords(i, j);
    compress_elements(Aelem, rows(), opts, int(coords.rows()), packed.data());
  }
  // HSSMatrix::reset(): back to the uncompressed state, same dimensions and partition
  //                                                             (HSSMatrix.cpp:138-146)
  void reset() {
    if (this->h_) { m0_ = base::rows(); n0_ = base::cols(); }
    SP_d_struct_destroy(&this->h_);
    ulv_ = HSSFactors<scalar_t>();
    trailing_deleted_ = false;
  }
  bool is_compressed() const { return this->h_ != nullptr; }
  bool is_untouched() const { return this->h_ == nullptr; }
  bool active() const { return true; }
  bool leaf() const { return levels() <= 1; }
  void set_openmp_task_depth(int) {}
  // HSSMatrix::delete_trailing_block() (HSSMatrix.cpp:358-368): after the Schur
  // complement has been taken, only the eliminated (0,0) block is used again
  // (child(0) solves).  The engine keeps its arenas (they are shared by both
  // halves); what changes is that whole-matrix operations are refused from here on.
  void delete_trailing_block() { trailing_deleted_ = true; }
  // child(c): the view FrontHSS uses -- child(0)->forward_solve(w, b, true),
  // child(0)->backward_solve(w, x), child(0)->ULV().Vhat(), rows / cols
  class Child {
   public:
    std::size_t rows() const { return rows_; }
    std::size_t cols() const { return cols_; }
    const HSSFactors<scalar_t>& ULV() const {
      if (c_ != 0) throw std::logic_error("child(1)->ULV(): only the eliminated (0,0) block has factors");
      H_->ulv_.Vhat() = H_->Vhat();
      return H_->ulv_;
    }
    void forward_solve(WorkSolve<scalar_t>& w, const DenseM_t& b, bool partial) const {
      if (c_ != 0 || !partial) throw std::logic_error("child(c)->forward_solve: child 0 with partial = true (as FrontHSS calls it)");
      w.reduced_rhs = H_->child0_forward_solve(w, b);
      w.x = H_->child0_x(w);
    }
    void backward_solve(WorkSolve<scalar_t>& w, DenseM_t& x) const {
      if (c_ != 0) throw std::logic_error("child(c)->backward_solve: child 0 only");
      H_->set_child0_x(w, w.x);
      H_->child0_backward_solve(w, x);
    }
   private:
    const HSSMatrix* H_ = nullptr;
    int c_ = 0;
    std::size_t rows_ = 0, cols_ = 0;
    friend class HSSMatrix;
  };
  const Child* child(int c) const {
    if (c < 0 || c > 1 || !this->h_) throw std::out_of_range("HSSMatrix::child");
    int z[7];
    base::check(SB200_d_hss_schur_sizes(this->h_, z), "schur_sizes");
    Child& ch = child_[c];
    ch.H_ = this; ch.c_ = c;
    ch.rows_ = c == 0 ? this->rows() - z[0] : z[0];
    ch.cols_ = c == 0 ? this->cols() - z[1] : z[1];
    return &ch;
  }
  const HSSFactors<scalar_t>& ULV() const { return ulv_; }
  // HSSMatrix::draw(of, rlo, clo) (HSSMatrix.cpp:370-405): gnuplot rectangles,
  // one per leaf diagonal block and per low-rank off-diagonal block, labelled
  // with the ranks
  void draw(std::ostream& of, std::size_t rlo = 0, std::size_t clo = 0) const {
    const int N = SB200_d_hss_node_table(this->h_, nullptr);
    if (N <= 0) return;
    std::vector<long long> t(std::size_t(N) * 10);
    SB200_d_hss_node_table(this->h_, t.data());
    const long long n = t[3];
    for (int i = 0; i < N; i++) {
      const long long* r = &t[std::size_t(i) * 10];
      const long long ro = rlo + r[5], co = clo + r[6];
      if (r[1] < 0) {
        of << "set obj rect from " << co << ", " << n - ro << " to " << co + r[4] << ", " << n - (ro + r[3])
           << " fc rgb 'red'" << std::endl;
      } else {
        const long long* a = &t[std::size_t(r[1]) * 10];
        const long long* b = &t[std::size_t(r[2]) * 10];
        of << "set obj rect from " << co + a[4] << ", " << n - ro << " to " << co + r[4] << ", " << n - (ro + a[3])
           << " fc rgb 'green' # B01 " << a[7] << " x " << b[8] << std::endl;
        of << "set obj rect from " << co << ", " << n - (ro + a[3]) << " to " << co + a[4] << ", " << n - (ro + r[3])
           << " fc rgb 'green' # B10 " << b[7] << " x " << a[8] << std::endl;
      }
    }
  }
  // HSSMatrix::read(fname) / write(fname)                      (HSSMatrix.cpp:438-510)
  static HSSMatrix read(const std::string& fname) {
    CSPStructMat h = nullptr;
    if (SB200_d_hss_read(&h, fname.c_str())) throw std::runtime_error("HSSMatrix::read failed");
    return HSSMatrix(h);
  }
  void write(const std::string& fname) const { base::check(SB200_d_hss_write(this->h_, fname.c_str()), "write"); }

  // apply / applyC / mult                                      (HSSMatrix.apply.hpp:34-53)
  DenseM_t apply(const DenseM_t& b) const {
    if (trailing_deleted_) throw std::logic_error("HSSMatrix::apply after delete_trailing_block()");
    DenseM_t c(this->rows(), b.cols());
    this->mult(Trans::N, b, c);
    return c;
  }
  DenseM_t applyC(const DenseM_t& b) const {
    DenseM_t c(this->cols(), b.cols());
    this->mult(Trans::C, b, c);
    return c;
  }
  // ULV: factor (base), partial_factor, solve (base), forward / backward solve
  //                                    (HSSMatrix.factor.hpp:35-50, solve.hpp:35-66)
  void partial_factor() { base::check(SB200_d_hss_partial_factor(this->h_), "partial_factor"); }
  void forward_solve(WorkSolve<scalar_t>& w, const DenseM_t& b, bool partial = false) const {
    if (partial) throw std::logic_error("forward_solve(partial = true) is a member of child(0): use child0_forward_solve");
    base::check(SB200_d_hss_forward_solve(this->h_, int(b.cols()), b.data(), int(b.ld())), "forward_solve");
    w.nrhs = int(b.cols()); w.partial = false;
  }
  void backward_solve(WorkSolve<scalar_t>& w, DenseM_t& x) const {
    base::check(SB200_d_hss_backward_solve(this->h_, w.nrhs, x.data(), int(x.ld())), "backward_solve");
  }
  // child(0)->forward_solve(w, b0, true) / child(0)->backward_solve(w, x0) as
  // FrontHSS calls them (FrontHSS.cpp:452-462, 487-495); returns reduced_rhs
  DenseM_t child0_forward_solve(WorkSolve<scalar_t>& w, const DenseM_t& b0) const {
    int z[7];
    base::check(SB200_d_hss_schur_sizes(this->h_, z), "schur_sizes");
    DenseM_t red(z[2], b0.cols());
    base::check(SB200_d_hss_partial_forward_solve(this->h_, int(b0.cols()), b0.data(), int(b0.ld()),
                                                  red.data(), int(red.ld())), "partial forward_solve");
    w.nrhs = int(b0.cols()); w.partial = true;
    return red;
  }
  DenseM_t child0_x(const WorkSolve<scalar_t>& w) const {          // w.x of the reference
    int z[7];
    base::check(SB200_d_hss_schur_sizes(this->h_, z), "schur_sizes");
    DenseM_t x(z[3], w.nrhs);
    base::check(SB200_d_hss_partial_x(this->h_, w.nrhs, x.data(), int(x.ld()), 0), "partial_x");
    return x;
  }
  void set_child0_x(const WorkSolve<scalar_t>& w, DenseM_t& x) const {
    base::check(SB200_d_hss_partial_x(this->h_, w.nrhs, x.data(), int(x.ld()), 1), "partial_x");
  }
  void child0_backward_solve(WorkSolve<scalar_t>& w, DenseM_t& x0) const {
    base::check(SB200_d_hss_partial_backward_solve(this->h_, w.nrhs, x0.data(), int(x0.ld())),
                "partial backward_solve");
  }
  // Schur complement of the (0,0) block                       (HSSMatrix.Schur.hpp:40-215)
  void Schur_update(DenseM_t& Theta, DenseM_t& DUB01, DenseM_t& Phi) const {
    int z[7];
    base::check(SB200_d_hss_schur_sizes(this->h_, z), "schur_sizes");
    Theta = DenseM_t(z[0], z[2]); DUB01 = DenseM_t(z[3], z[4]); Phi = DenseM_t(z[1], z[3]);
    base::check(SB200_d_hss_schur_update(this->h_, Theta.data(), int(Theta.ld()), DUB01.data(),
                                         int(DUB01.ld()), Phi.data(), int(Phi.ld())), "Schur_update");
  }
  DenseM_t Vhat() const {                                        // child(0)->ULV().Vhat()
    int z[7];
    base::check(SB200_d_hss_schur_sizes(this->h_, z), "schur_sizes");
    DenseM_t V(z[3], z[2]);
    base::check(SB200_d_hss_vhat(this->h_, V.data(), int(V.ld())), "Vhat");
    return V;
  }
  // the reference's 4th argument (Theta Vhat^H or Vhat^H Phi^H) is a cached
  // product of the others; the engine recomputes nothing from it
  void Schur_product_direct(const DenseM_t& Theta, const DenseM_t& DUB01, const DenseM_t& Phi,
                            const DenseM_t& /*ThetaVhatC_or_VhatCPhiC*/, const DenseM_t& R,
                            DenseM_t& Sr, DenseM_t& Sc) const {
    if (Sr.rows() != R.rows() || Sr.cols() != R.cols()) Sr = DenseM_t(R.rows(), R.cols());
    if (Sc.rows() != R.rows() || Sc.cols() != R.cols()) Sc = DenseM_t(R.rows(), R.cols());
    base::check(SB200_d_hss_schur_product_direct(
                    this->h_, Theta.data(), int(Theta.ld()), DUB01.data(), int(DUB01.ld()), Phi.data(),
                    int(Phi.ld()), int(R.cols()), R.data(), int(R.ld()), Sr.data(), int(Sr.ld()),
                    Sc.data(), int(Sc.ld())), "Schur_product_direct");
  }
  void Schur_product_indirect(const DenseM_t& DUB01, const DenseM_t& R0, const DenseM_t& R1,
                              const DenseM_t& Sr1, const DenseM_t& Sc1, DenseM_t& Sr, DenseM_t& Sc) const {
    Sr = DenseM_t(R1.rows(), R1.cols()); Sc = DenseM_t(R1.rows(), R1.cols());
    base::check(SB200_d_hss_schur_product_indirect(
                    this->h_, DUB01.data(), int(DUB01.ld()), int(R1.cols()), R0.data(), int(R0.ld()),
                    R1.data(), int(R1.ld()), Sr1.data(), int(Sr1.ld()), Sc1.data(), int(Sc1.ld()),
                    Sr.data(), int(Sr.ld()), Sc.data(), int(Sc.ld())), "Schur_product_indirect");
  }
  // element extraction                                        (HSSMatrix.extract.hpp:8-188)
  scalar_t get(std::size_t i, std::size_t j) const {
    const int I = int(i), J = int(j);
    scalar_t v = 0;
    base::check(SB200_d_hss_extract(this->h_, 1, &I, 1, &J, &v, 1, 0), "get");
    return v;
  }
  DenseM_t extract(const std::vector<std::size_t>& I, const std::vector<std::size_t>& J) const {
    DenseM_t B(I.size(), J.size());
    extract_impl(I, J, B, 0);
    return B;
  }
  void extract_add(const std::vector<std::size_t>& I, const std::vector<std::size_t>& J, DenseM_t& B) const {
    extract_impl(I, J, B, 1);
  }
  // statistics                                                (HSSMatrix.cpp:290-356)
  std::size_t levels() const { return SB200_d_struct_levels(this->h_); }
  std::size_t factor_nonzeros() const { return SB200_d_struct_factor_nonzeros(this->h_); }
  void print_info() const { SB200_d_struct_print_info(this->h_); }
  DenseM_t dense() const {
    DenseM_t A(this->rows(), this->cols());
    base::check(SB200_d_struct_dense(this->h_, A.data(), int(A.ld())), "dense");
    return A;
  }

 private:
  std::size_t m0_ = 0, n0_ = 0;     // dimensions before compression
  std::vector<int> tree_sizes_, tree_nchild_;   // partition given at construction (ClusterTree ctor), pre-order
  bool trailing_deleted_ = false;
  mutable HSSFactors<scalar_t> ulv_;
  mutable Child child_[2];
  void adopt(CSPStructMat s) {
    SP_d_struct_destroy(&this->h_);
    this->h_ = s;
    trailing_deleted_ = false;
  }
  void compress_elements(const elem_t& Aelem, std::size_t n, const opts_t& opts, int d, const scalar_t* coords) {
    opts_t o(opts);
    CSPStructMat s = nullptr;
    if (SB200_d_hss_from_element_blocks_ex(&s, int(n), &elem_trampoline, const_cast<elem_t*>(&Aelem), o.c(),
                                           int(tree_sizes_.size()), tree_sizes_.data(), tree_nchild_.data(), d, coords))
      throw std::invalid_argument("HSSMatrix::compress(Amult, Aelem) failed");
    adopt(s);
  }
  static void elem_trampoline(int nI, const int* I, int nJ, const int* J, double* B, int ldB, void* user) {
    std::vector<std::size_t> Iv(I, I + nI), Jv(J, J + nJ);
    DenseMatrixWrapper<scalar_t> Bw(nI, nJ, B, ldB);
    (*static_cast<const elem_t*>(user))(Iv, Jv, Bw);
  }
  void extract_impl(const std::vector<std::size_t>& I, const std::vector<std::size_t>& J, DenseM_t& B,
                    int add) const {
    if (I.empty() || J.empty()) return;
    std::vector<int> i(I.begin(), I.end()), j(J.begin(), J.end());
    base::check(SB200_d_hss_extract(this->h_, int(i.size()), i.data(), int(j.size()), j.data(), B.data(),
                                    int(B.ld()), add), "extract");
  }
};

// free function apply_HSS(op, A, B, beta, C): C = op(A) B + beta C   (HSSMatrix.hpp:705-713)
// free function draw(H, name): writes name.gnuplot                    (HSSMatrix.hpp:705-706)
template <typename scalar_t> void draw(const HSSMatrix<scalar_t>& H, const std::string& name) {
  std::ofstream of("plot" + name + ".gnuplot");
  of << "set terminal pdf enhanced color size 5,4" << std::endl << "set output '" << name << ".pdf'" << std::endl;
  H.draw(of);
  of << "set xrange [0:" << H.cols() << "]" << std::endl << "set yrange [0:" << H.rows() << "]" << std::endl
     << "plot x lt -1 notitle" << std::endl;
}

template <typename scalar_t>
void apply_HSS(Trans op, const HSSMatrix<scalar_t>& A, const DenseMatrix<scalar_t>& B, scalar_t beta,
               DenseMatrix<scalar_t>& C) {
  detail::check(SB200_d_hss_apply(A.handle(), char(op), int(B.cols()), B.data(), int(B.ld()), beta, C.data(),
                                  int(C.ld())), "apply_HSS");
}

}  // namespace HSS

namespace BLR {

// reference BLROptions.hpp:59-69
enum class LowRankAlgorithm { RRQR, ACA, BACA };
enum class Admissibility { STRONG, WEAK };
enum class BLRFactorAlgorithm { COLWISE, RL, LL, COMB, STAR };
enum class CompressionKernel { HALF, FULL };
inline std::string get_name(LowRankAlgorithm a) {
  return a == LowRankAlgorithm::RRQR ? "RRQR" : a == LowRankAlgorithm::ACA ? "ACA" : "BACA";
}
inline std::string get_name(Admissibility a) { return a == Admissibility::STRONG ? "strong" : "weak"; }
inline std::string get_name(BLRFactorAlgorithm a) {
  switch (a) {
    case BLRFactorAlgorithm::COLWISE: return "Colwise";
    case BLRFactorAlgorithm::RL: return "RL";
    case BLRFactorAlgorithm::LL: return "LL";
    case BLRFactorAlgorithm::COMB: return "Comb";
    case BLRFactorAlgorithm::STAR: return "Star";
  }
  return "UNKNOWN";
}

// reference BLROptions.hpp:81-142 (defaults :128-140: rel 1e-4, abs 1e-12, leaf
// 256, max_rank 5000, RRQR, WEAK, RL, HALF).  The engine implements RRQR
// compression, weak admissibility and the RL and LL schedules of the
// factorization; the other enum values are accepted and mapped onto those
// (COLWISE / COMB / STAR run as RL).
template <typename scalar_t> class BLROptions : public structured::StructuredOptions<scalar_t> {
 public:
  BLROptions() : structured::StructuredOptions<scalar_t>(structured::Type::BLR) {
    this->set_rel_tol(1e-4); this->set_abs_tol(1e-12);
    this->set_leaf_size(256); this->set_max_rank(5000);
  }
  BLROptions(const structured::StructuredOptions<scalar_t>& s) : structured::StructuredOptions<scalar_t>(s) {
    this->set_type(structured::Type::BLR);
  }
  void set_low_rank_algorithm(LowRankAlgorithm a) { lr_ = a; }
  void set_admissibility(Admissibility a) { adm_ = a; }
  void set_BACA_blocksize(int b) { baca_ = b; }
  void set_BLR_factor_algorithm(BLRFactorAlgorithm a) { alg_ = a; }
  void set_compression_kernel(CompressionKernel a) { ck_ = a; }
  void set_pivot_threshold(double t) { pivot_ = t; }
  LowRankAlgorithm low_rank_algorithm() const { return lr_; }
  Admissibility admissibility() const { return adm_; }
  int BACA_blocksize() const { return baca_; }
  BLRFactorAlgorithm BLR_factor_algorithm() const { return alg_; }
  CompressionKernel compression_kernel() const { return ck_; }
  double pivot_threshold() const { return pivot_; }
  // --blr_{rel_tol,abs_tol,leaf_size,max_rank,verbose,quiet}   (BLROptions.cpp:58-200)
  void set_from_command_line(int argc, const char* const* argv) override {
    this->scan(argc, argv, "--blr_");
    if (auto v = detail::arg_value(argc, argv, "--blr_pivot_threshold")) set_pivot_threshold(std::atof(v));
    if (auto v = detail::arg_value(argc, argv, "--blr_factor_algorithm")) {
      const std::string s(v);
      if (s == "RL") alg_ = BLRFactorAlgorithm::RL;
      else if (s == "LL") alg_ = BLRFactorAlgorithm::LL;
      else if (s == "Comb") alg_ = BLRFactorAlgorithm::COMB;
      else if (s == "Star") alg_ = BLRFactorAlgorithm::STAR;
      else if (s == "Colwise") alg_ = BLRFactorAlgorithm::COLWISE;
    }
  }

 private:
  LowRankAlgorithm lr_ = LowRankAlgorithm::RRQR;
  Admissibility adm_ = Admissibility::WEAK;
  BLRFactorAlgorithm alg_ = BLRFactorAlgorithm::RL;
  CompressionKernel ck_ = CompressionKernel::HALF;
  int baca_ = 4;
  double pivot_ = -1.;
};

// reference BLRMatrix<T> (src/BLR/BLRMatrix.hpp:68-291).  Tiles come from
// ClusterTree(n).refine(opts.leaf_size()) as in test/test_BLR_seq.cpp:136-145;
// admissibility is weak (every off-diagonal tile is a compression candidate).
template <typename scalar_t> class BLRMatrix : public structured::StructuredMatrix<scalar_t> {
  using base = structured::StructuredMatrix<scalar_t>;
  using DenseM_t = DenseMatrix<scalar_t>;

 public:
  using Opts_t = BLROptions<scalar_t>;
  using adm_t = DenseMatrix<bool>;
  BLRMatrix() = default;
  explicit BLRMatrix(CSPStructMat h) : base(h) {}
  BLRMatrix(BLRMatrix&& o) noexcept : base(std::move(o)), view_(o.view_) { o.view_ = false; }
  BLRMatrix& operator=(BLRMatrix&& o) noexcept {
    if (this != &o) {
      if (view_) this->h_ = nullptr;
      base::operator=(std::move(o));
      view_ = o.view_;
      o.view_ = false;
    }
    return *this;
  }
  // BLRMatrix::compress(A, admissible, opts)                   (BLRMatrix.cpp:91-111)
  void compress(const DenseM_t& A, const Opts_t& opts) {
    CSPStructMat s = nullptr;
    if (SP_d_struct_from_dense(&s, int(A.rows()), int(A.cols()), A.data(), int(A.ld()), opts.c()))
      throw std::invalid_argument("BLRMatrix::compress failed");
    reset(s);
  }
  // what the engine does not build is refused, never replaced by something else:
  // ACA / BACA tile compression (LRTile.cpp:77-115); COLWISE / COMB / STAR are
  // refused by the library itself (SB200_d_blr_*_ex)
  static void check_supported(const Opts_t& opts) {
    if (opts.low_rank_algorithm() != LowRankAlgorithm::RRQR)
      throw std::invalid_argument("BLR: low-rank algorithm " + get_name(opts.low_rank_algorithm()) +
                                  " is not implemented by this engine (RRQR, the reference's default, is)");
  }
  // BLRMatrix::compress_and_factor(A, admissible, opts)        (BLRMatrix.cpp:113-241)
  void compress_and_factor(const DenseM_t& A, const Opts_t& opts) {
    if (A.rows() != A.cols()) throw std::invalid_argument("BLR: only square matrices are supported");
    CSPStructMat s = nullptr;
    check_supported(opts);
    SB200BLRParams p{opts.pivot_threshold(), int(opts.BLR_factor_algorithm()), nullptr, 0, nullptr, 0, nullptr, 0};
    if (SB200_d_blr_compress_and_factor_ex(&s, int(A.rows()), A.data(), int(A.ld()), opts.c(), &p))
      throw std::invalid_argument("BLRMatrix::compress_and_factor failed");
    reset(s);
  }
  // admissible: (number of tiles)^2, true = the tile may be compressed
  void compress_and_factor(const DenseM_t& A, const adm_t& admissible, const Opts_t& opts) {
    if (A.rows() != A.cols()) throw std::invalid_argument("BLR: only square matrices are supported");
    std::vector<int> adm(admissible.rows() * admissible.cols());
    for (std::size_t j = 0; j < admissible.cols(); j++)
      for (std::size_t i = 0; i < admissible.rows(); i++) adm[i + j * admissible.rows()] = admissible(i, j) ? 1 : 0;
    check_supported(opts);
    SB200BLRParams p{opts.pivot_threshold(), int(opts.BLR_factor_algorithm()), adm.data(), int(admissible.rows()), nullptr, 0, nullptr, 0};
    CSPStructMat s = nullptr;
    if (SB200_d_blr_compress_and_factor_ex(&s, int(A.rows()), A.data(), int(A.ld()), opts.c(), &p))
      throw std::invalid_argument("BLRMatrix::compress_and_factor failed");
    reset(s);
  }
  // BLRMatrix::construct_and_partial_factor(A11, A12, A21, A22, B11, B12, B21,
  // tiles1, tiles2, admissible, opts) (BLRMatrix.cpp:739-1037): one object
  // holds F11, F12 and F21; A22 is overwritten with the Schur complement and
  // A11, A12, A21 are cleared, as in the reference.
  static BLRMatrix construct_and_partial_factor(DenseM_t& A11, DenseM_t& A12, DenseM_t& A21, DenseM_t& A22,
                                                const Opts_t& opts) {
    CSPStructMat s = nullptr;
    check_supported(opts);
    SB200BLRParams p{opts.pivot_threshold(), int(opts.BLR_factor_algorithm()), nullptr, 0, nullptr, 0, nullptr, 0};
    if (SB200_d_blr_partial_factor_ex(&s, int(A11.rows()), int(A22.rows()), A11.data(), int(A11.ld()), A12.data(),
                                      int(A12.ld()), A21.data(), int(A21.ld()), A22.data(), int(A22.ld()),
                                      opts.c(), &p))
      throw std::invalid_argument("BLRMatrix::construct_and_partial_factor failed");
    A11.clear(); A12.clear(); A21.clear();
    return BLRMatrix(s);
  }
  // The reference's own signature (BLRMatrix.hpp:184-191; FrontBLR.cpp:429-433):
  // three result matrices and the tile partitions of the separator / update
  // blocks.  B11 owns the factored front; B12 and B21 are views of it (the
  // engine keeps F11, F12 and F21 in one object), so they must not outlive B11.
  static void construct_and_partial_factor(DenseM_t& A11, DenseM_t& A12, DenseM_t& A21, DenseM_t& A22,
                                           BLRMatrix& B11, BLRMatrix& B12, BLRMatrix& B21,
                                           const std::vector<std::size_t>& tiles1,
                                           const std::vector<std::size_t>& tiles2, const adm_t& admissible,
                                           const Opts_t& opts) {
    check_supported(opts);
    std::vector<int> t1(tiles1.begin(), tiles1.end()), t2(tiles2.begin(), tiles2.end());
    std::vector<int> adm(admissible.rows() * admissible.cols());
    for (std::size_t j = 0; j < admissible.cols(); j++)
      for (std::size_t i = 0; i < admissible.rows(); i++) adm[i + j * admissible.rows()] = admissible(i, j) ? 1 : 0;
    SB200BLRParams p{opts.pivot_threshold(), int(opts.BLR_factor_algorithm()), adm.empty() ? nullptr : adm.data(),
                     int(admissible.rows()), t1.data(), int(t1.size()), t2.empty() ? nullptr : t2.data(), int(t2.size())};
    CSPStructMat s = nullptr;
    if (SB200_d_blr_partial_factor_ex(&s, int(A11.rows()), int(A22.rows()), A11.data(), int(A11.ld()), A12.data(),
                                      int(A12.ld()), A21.data(), int(A21.ld()), A22.data(), int(A22.ld()),
                                      opts.c(), &p))
      throw std::invalid_argument("BLRMatrix::construct_and_partial_factor failed");
    A11.clear(); A12.clear(); A21.clear();
    B11.reset(s);
    B12.view(s);
    B21.view(s);
  }
  // pivots of the factored block: empty -- the row interchanges are applied inside
  // trsmLNU_gemm (DenseMatrix::laswp with an empty vector is a no-op), so the
  // reference's `bloc.laswp(F11.piv(), true); trsmLNU_gemm(F11, F21, bloc, bupd, d)`
  // (FrontBLR.cpp:529-531) computes the same thing unchanged
  const std::vector<int>& piv() const { static const std::vector<int> none; return none; }
  // B1 <- L11^{-1} P B1, B2 <- B2 - F21 B1                     (BLRMatrix.cpp:1552-1608)
  static void trsmLNU_gemm(const BLRMatrix& F1, const BLRMatrix& /*F2*/, DenseM_t& B1, DenseM_t& B2,
                           int /*task_depth*/) {
    DenseM_t b = stack(B1, B2);
    F1.trsmLNU_gemm(b);
    unstack(b, B1, B2);
  }
  // B1 <- U11^{-1} (B1 - F12 B2)                               (BLRMatrix.cpp:1610-1665)
  static void gemm_trsmUNN(const BLRMatrix& F1, const BLRMatrix& /*F2*/, DenseM_t& B1, DenseM_t& B2,
                           int /*task_depth*/) {
    DenseM_t b = stack(B1, B2);
    F1.gemm_trsmUNN(b);
    unstack(b, B1, B2);
  }
  // the two halves of the front solve on b = [b_sep; b_upd]
  // (laswp + trsmLNU_gemm, gemm_trsmUNN; BLRMatrix.cpp:1552-1665)
  void trsmLNU_gemm(DenseM_t& b) const {
    base::check(SB200_d_blr_partial_forward_solve(this->h_, int(b.cols()), b.data(), int(b.ld())), "trsmLNU_gemm");
  }
  void gemm_trsmUNN(DenseM_t& y) const {
    base::check(SB200_d_blr_partial_backward_solve(this->h_, int(y.cols()), y.data(), int(y.ld())), "gemm_trsmUNN");
  }
  std::size_t sep_rows() const { return SB200_d_blr_sep_rows(this->h_); }
  std::size_t rowblocks() const { return SB200_d_blr_tiles(this->h_); }
  std::size_t colblocks() const { return SB200_d_blr_tiles(this->h_); }
  std::size_t dense_tiles() const { return SB200_d_blr_dense_tiles(this->h_); }

  ~BLRMatrix() override { if (view_) this->h_ = nullptr; }   // a view does not own the engine object

 private:
  bool view_ = false;
  void reset(CSPStructMat s) { if (view_) this->h_ = nullptr; SP_d_struct_destroy(&this->h_); this->h_ = s; view_ = false; }
  void view(CSPStructMat s) { reset(nullptr); this->h_ = s; view_ = true; }
  static DenseM_t stack(const DenseM_t& B1, const DenseM_t& B2) {
    DenseM_t b(B1.rows() + B2.rows(), B1.cols());
    for (std::size_t c = 0; c < B1.cols(); c++) {
      for (std::size_t i = 0; i < B1.rows(); i++) b(i, c) = B1(i, c);
      for (std::size_t i = 0; i < B2.rows(); i++) b(B1.rows() + i, c) = B2(i, c);
    }
    return b;
  }
  static void unstack(const DenseM_t& b, DenseM_t& B1, DenseM_t& B2) {
    for (std::size_t c = 0; c < B1.cols(); c++) {
      for (std::size_t i = 0; i < B1.rows(); i++) B1(i, c) = b(i, c);
      for (std::size_t i = 0; i < B2.rows(); i++) B2(i, c) = b(B1.rows() + i, c);
    }
  }
};

}  // namespace BLR
}  // namespace strumpack
