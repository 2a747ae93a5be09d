// strumpack_b200 -- C++ host-side mirror of the reference's structured-matrix
// interface, header-only on top of the C ABI (include/sb200_structured.h).
//
// Same class/function names, argument meaning and error behaviour as
//   strumpack::structured::StructuredMatrix<T>   reference src/structured/StructuredMatrix.hpp:209-418
//   strumpack::structured::construct_from_dense  reference src/structured/StructuredMatrix.cpp:53-127
//   strumpack::structured::StructuredOptions<T>  reference src/structured/StructuredOptions.hpp:106-162
//   strumpack::HSS::HSSMatrix<T>                 reference src/HSS/HSSMatrix.hpp:95-511 (hot-path subset)
//   strumpack::DenseMatrix<T> / DenseMatrixWrapper<T>  reference src/dense/DenseMatrix.hpp:139-146
// so that code written against the reference (e.g. examples/dense/testStructured.cpp,
// test/test_HSS_seq.cpp:235-250) compiles against this header with only the
// include changed.  Only T = double is implemented in round 1.
#pragma once
#include <cstddef>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../sb200_structured.h"

namespace strumpack {

enum class Trans : char { N = 'N', C = 'C', T = 'T' };

// column-major owning matrix (reference DenseMatrix: data_, rows_, cols_, ld_)
template <typename scalar_t> class DenseMatrix {
 public:
  DenseMatrix() = default;
  DenseMatrix(std::size_t m, std::size_t n) : rows_(m), cols_(n), ld_(m ? m : 1), own_(m * n) {
    data_ = own_.data();
  }
  virtual ~DenseMatrix() = default;
  std::size_t rows() const { return rows_; }
  std::size_t cols() const { return cols_; }
  std::size_t ld() const { return ld_; }
  scalar_t* data() { return data_; }
  const scalar_t* data() const { return data_; }
  scalar_t& operator()(std::size_t i, std::size_t j) { return data_[i + j * ld_]; }
  const scalar_t& operator()(std::size_t i, std::size_t j) const { return data_[i + j * ld_]; }

 protected:
  scalar_t* data_ = nullptr;
  std::size_t rows_ = 0, cols_ = 0, ld_ = 1;
  std::vector<scalar_t> own_;
};

// non-owning view (reference DenseMatrixWrapper)
template <typename scalar_t> class DenseMatrixWrapper : public DenseMatrix<scalar_t> {
 public:
  DenseMatrixWrapper(std::size_t m, std::size_t n, scalar_t* D, std::size_t ld) {
    this->data_ = D; this->rows_ = m; this->cols_ = n; this->ld_ = ld;
  }
};

namespace structured {

enum class Type : int { HSS = SP_TYPE_HSS, BLR = SP_TYPE_BLR };

// reference StructuredOptions.hpp:106-162 (same defaults)
template <typename scalar_t> class StructuredOptions {
 public:
  StructuredOptions() { SP_d_struct_default_options(&o_); }
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

 private:
  CSPOptions o_;
};

// reference StructuredMatrix.hpp:209-418: unsupported operations throw
template <typename scalar_t> class StructuredMatrix {
  static_assert(sizeof(scalar_t) == sizeof(double), "round 1 implements double only");

 public:
  explicit StructuredMatrix(CSPStructMat h) : h_(h) {}
  StructuredMatrix(const StructuredMatrix&) = delete;
  StructuredMatrix& operator=(const StructuredMatrix&) = delete;
  virtual ~StructuredMatrix() { SP_d_struct_destroy(&h_); }

  std::size_t rows() const { return SP_d_struct_rows(h_); }
  std::size_t cols() const { return SP_d_struct_cols(h_); }
  std::size_t memory() const { return SP_d_struct_memory(h_); }
  std::size_t nonzeros() const { return SP_d_struct_nonzeros(h_); }
  std::size_t rank() const { return SP_d_struct_rank(h_); }

  // y = op(A) x                                  (StructuredMatrix.hpp:280-300)
  void mult(Trans op, const DenseMatrix<scalar_t>& x, DenseMatrix<scalar_t>& y) const {
    mult(op, int(x.cols()), x.data(), int(x.ld()), y.data(), int(y.ld()));
  }
  void mult(Trans op, int m, const scalar_t* x, int ldx, scalar_t* y, int ldy) const {
    check(SP_d_struct_mult(h_, char(op), m, x, ldx, y, ldy), "mult");
  }
  void factor() { check(SP_d_struct_factor(h_), "factor"); }
  // b <- A^{-1} b                                (StructuredMatrix.hpp:340-360)
  void solve(DenseMatrix<scalar_t>& b) const { solve(int(b.cols()), b.data(), int(b.ld())); }
  void solve(int nrhs, scalar_t* b, int ldb) const {
    check(SP_d_struct_solve(h_, nrhs, b, ldb), "solve");
  }
  void shift(scalar_t s) { check(SP_d_struct_shift(h_, s), "shift"); }
  CSPStructMat handle() const { return h_; }

 protected:
  static void check(int rc, const char* what) {
    if (rc) throw std::logic_error(std::string("strumpack_b200: ") + what + " failed");
  }
  CSPStructMat h_ = nullptr;
};

// reference StructuredMatrix.cpp:53-127
template <typename scalar_t>
std::unique_ptr<StructuredMatrix<scalar_t>> construct_from_dense(
    const DenseMatrix<scalar_t>& A, const StructuredOptions<scalar_t>& opts) {
  CSPStructMat h = nullptr;
  if (SP_d_struct_from_dense(&h, int(A.rows()), int(A.cols()), A.data(), int(A.ld()), opts.c()))
    throw std::invalid_argument("construct_from_dense failed");
  return std::unique_ptr<StructuredMatrix<scalar_t>>(new StructuredMatrix<scalar_t>(h));
}
template <typename scalar_t>
std::unique_ptr<StructuredMatrix<scalar_t>> construct_from_dense(
    int rows, int cols, const scalar_t* A, int ldA, const StructuredOptions<scalar_t>& opts) {
  CSPStructMat h = nullptr;
  if (SP_d_struct_from_dense(&h, rows, cols, A, ldA, opts.c()))
    throw std::invalid_argument("construct_from_dense failed");
  return std::unique_ptr<StructuredMatrix<scalar_t>>(new StructuredMatrix<scalar_t>(h));
}

}  // namespace structured

namespace HSS {

// hot-path subset of reference HSSMatrix<T> (src/HSS/HSSMatrix.hpp:95-511)
template <typename scalar_t> class HSSMatrix : public structured::StructuredMatrix<scalar_t> {
  using base = structured::StructuredMatrix<scalar_t>;

 public:
  explicit HSSMatrix(CSPStructMat h) : base(h) {}
  // HSSMatrix::read(fname)                        (HSSMatrix.cpp:488-510)
  static HSSMatrix read(const std::string& fname) {
    CSPStructMat h = nullptr;
    if (SB200_d_hss_read(&h, fname.c_str())) throw std::runtime_error("HSSMatrix::read failed");
    return HSSMatrix(h);
  }
  HSSMatrix(HSSMatrix&& o) noexcept : base(o.h_) { o.h_ = nullptr; }
  void write(const std::string& fname) const { base::check(SB200_d_hss_write(this->h_, fname.c_str()), "write"); }
  DenseMatrix<scalar_t> apply(const DenseMatrix<scalar_t>& b) const {   // apply.hpp:39-45
    DenseMatrix<scalar_t> c(this->rows(), b.cols());
    this->mult(Trans::N, b, c);
    return c;
  }
  DenseMatrix<scalar_t> applyC(const DenseMatrix<scalar_t>& b) const {  // apply.hpp:47-53
    DenseMatrix<scalar_t> c(this->cols(), b.cols());
    this->mult(Trans::C, b, c);
    return c;
  }
  std::size_t levels() const { return SB200_d_struct_levels(this->h_); }
  std::size_t factor_nonzeros() const { return SB200_d_struct_factor_nonzeros(this->h_); }
  void print_info() const { SB200_d_struct_print_info(this->h_); }
  DenseMatrix<scalar_t> dense() const {
    DenseMatrix<scalar_t> A(this->rows(), this->cols());
    base::check(SB200_d_struct_dense(this->h_, A.data(), int(A.ld())), "dense");
    return A;
  }
};

}  // namespace HSS
}  // namespace strumpack
