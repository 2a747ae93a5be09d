// strumpack_b200 -- host-side HSS tree container (see hss_tree.hpp).
#include "hss_tree.hpp"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <stdexcept>

namespace sb200 {

std::vector<int32_t> ipiv_to_gather(const int32_t* ipiv, int n) {
  // laswp(fwd): for i = 0..n-1 swap rows i and ipiv[i]-1
  // (reference DenseMatrix::laswp, src/dense/DenseMatrix.cpp:288-297).
  std::vector<int32_t> g(n);
  for (int i = 0; i < n; i++) g[i] = i;
  for (int i = 0; i < n; i++) {
    int p = ipiv[i] - 1;
    if (p < 0 || p >= n) throw std::invalid_argument("ipiv out of range");
    if (p != i) std::swap(g[i], g[p]);
  }
  return g;
}

std::vector<int32_t> gather_to_ipiv(const int32_t* g, int n) {
  // Find sequential swaps reproducing the gather: after step i position i
  // must hold source g[i].
  std::vector<int32_t> cur(n), where(n), ipiv(n);
  for (int i = 0; i < n; i++) cur[i] = where[i] = i;
  for (int i = 0; i < n; i++) {
    int src = g[i];
    int p = where[src];  // current position of the wanted row
    ipiv[i] = p + 1;
    if (p != i) {
      int other = cur[i];
      std::swap(cur[i], cur[p]);
      where[src] = i;
      where[other] = p;
    }
  }
  return ipiv;
}

int HSSHost::max_rank() const {
  int r = 0;
  for (auto& n : nodes) r = std::max(r, std::max(n.u_rank, n.v_rank));
  return r;
}

static long long blk(const HSSNode& n, int which) {
  switch (which) {
    case 0: return n.off_D >= 0 ? 1LL * n.rows * n.cols : 0;
    case 1: return n.off_Eu >= 0 ? 1LL * (n.u_rows - n.u_rank) * n.u_rank : 0;
    case 2: return n.off_Ev >= 0 ? 1LL * (n.v_rows - n.v_rank) * n.v_rank : 0;
  }
  return 0;
}

long long HSSHost::nonzeros() const {
  long long nnz = 0;
  for (auto& n : nodes) {
    nnz += blk(n, 0) + blk(n, 1) + blk(n, 2);
    // the reference counts the pivot vectors too (HSSBasisID::nonzeros)
    nnz += (n.off_Pu >= 0 ? n.u_rows : 0) + (n.off_Pv >= 0 ? n.v_rows : 0);
    if (!n.leaf()) {
      nnz += 1LL * nodes[n.ch0].u_rank * nodes[n.ch1].v_rank;
      nnz += 1LL * nodes[n.ch1].u_rank * nodes[n.ch0].v_rank;
    }
  }
  return nnz;
}

long long HSSHost::memory_bytes() const {
  return (long long)(vals.size() * sizeof(double) +
                     perms.size() * sizeof(int32_t));
}

void HSSHost::finalize() {
  const int N = int(nodes.size());
  if (!N) throw std::invalid_argument("empty HSS tree");
  // offsets + depth, top-down (pre-order => parent index < child index)
  nodes[0].row_off = nodes[0].col_off = 0;
  nodes[0].depth = 0;
  for (int i = 0; i < N; i++) {
    auto& n = nodes[i];
    if (n.leaf()) continue;
    if (n.ch0 <= i || n.ch1 <= i || n.ch0 >= N || n.ch1 >= N)
      throw std::invalid_argument("HSS nodes must be in pre-order");
    auto& a = nodes[n.ch0];
    auto& b = nodes[n.ch1];
    a.parent = b.parent = i;
    a.row_off = n.row_off; a.col_off = n.col_off;
    b.row_off = n.row_off + a.rows; b.col_off = n.col_off + a.cols;
    a.depth = b.depth = n.depth + 1;
    if (a.rows + b.rows != n.rows || a.cols + b.cols != n.cols)
      throw std::invalid_argument("child sizes do not add up");
  }
  // heights, bottom-up
  int maxh = 0;
  for (int i = N - 1; i >= 0; i--) {
    auto& n = nodes[i];
    n.height = n.leaf() ? 0
      : 1 + std::max(nodes[n.ch0].height, nodes[n.ch1].height);
    maxh = std::max(maxh, n.height);
  }
  hptr.assign(maxh + 2, 0);
  for (auto& n : nodes) hptr[n.height + 1]++;
  for (int h = 0; h <= maxh; h++) hptr[h + 1] += hptr[h];
  by_height.resize(N);
  std::vector<int> pos(hptr.begin(), hptr.end() - 1);
  for (int i = 0; i < N; i++) by_height[pos[nodes[i].height]++] = i;
  // consistency of generator shapes
  for (int i = 0; i < N; i++) {
    auto& n = nodes[i];
    const bool root = (i == 0);
    if (n.leaf()) {
      if (n.off_D < 0 && n.rows * n.cols)
        throw std::invalid_argument("leaf without D");
      if (!root && (n.u_rows != n.rows || n.v_rows != n.cols))
        throw std::invalid_argument("leaf basis size mismatch");
    } else {
      auto& a = nodes[n.ch0];
      auto& b = nodes[n.ch1];
      if (!root && (n.u_rows != a.u_rank + b.u_rank ||
                    n.v_rows != a.v_rank + b.v_rank))
        throw std::invalid_argument("inner basis size mismatch");
      if ((n.off_B01 < 0 && a.u_rank * b.v_rank) ||
          (n.off_B10 < 0 && b.u_rank * a.v_rank))
        throw std::invalid_argument("inner node without B01/B10");
    }
    if (!root) {
      if (n.u_rank > n.u_rows || n.v_rank > n.v_rows)
        throw std::invalid_argument("rank larger than basis rows");
      if ((n.off_Pu < 0 && n.u_rows) || (n.off_Pv < 0 && n.v_rows))
        throw std::invalid_argument("basis without permutation");
    }
  }
}

// ---------------------------------------------------------------- flop model
static inline long long gemm_fl(long long m, long long n, long long k) {
  return 2 * m * n * k;
}

long long HSSHost::apply_flops() const {
  // SURVEY 8d F_apply(1): every generator entry is used once (2 flops)
  long long f = 0;
  for (int i = 0; i < int(nodes.size()); i++) {
    auto& n = nodes[i];
    f += 2 * blk(n, 0);
    if (i) f += 2 * (blk(n, 1) + blk(n, 2));
    if (!n.leaf()) {
      f += 2LL * nodes[n.ch0].u_rank * nodes[n.ch1].v_rank;
      f += 2LL * nodes[n.ch1].u_rank * nodes[n.ch0].v_rank;
    }
  }
  return f;
}

// Flop formulas exactly as the reference evaluates them (same mixed
// integer/double arithmetic), so the totals equal params::ULV_factor_flops and
// params::hss_solve_flops bit for bit:
//   gemm_flops   src/dense/BLASLAPACKWrapper.hpp:215-221
//   gelqf_flops  :621-628,  xxglq_flops :684-687,  getrf_flops :649-655
//   getrs_flops  :662-664,  trsm_flops  :450-456
//   LQ_flops     src/dense/DenseMatrix.hpp:1473-1478
static long long ref_gemm(long long m, long long n, long long k, double alpha,
                          double beta) {
  return (alpha != 0.) * m * n * (k * 2 - 1) +
         (alpha != 0. && beta != 0.) * m * n +
         (alpha != 0. && alpha != 1.) * m * n +
         (beta != 0. && beta != 1.) * m * n;
}
static long long ref_gelqf(long long m, long long n) {
  if (m > n)
    return n * (n * (.5 - (1. / 3.) * n + m) + m + 29. / 6.) +
           n * (n * (-.5 - (1. / 3.) * n + m) + m + 5. / 6.);
  else
    return m * (m * (-.5 - (1. / 3.) * m + n) + 2 * n + 29. / 6.) +
           m * (m * (.5 - (1. / 3.) * m + n) + 5. / 6.);
}
static long long ref_xxglq(long long m, long long n, long long k) {
  if (m == k) return 2 * m * m * (3 * n - m) / 3;
  else return 4 * m * n * k - 2 * (m + n) * k * k + 4 * k * k * k / 3;
}
static long long ref_getrf(long long m, long long n) {
  if (m < n) return (m / 2 * (m * (n - m / 3 - 1) + n) + 2 * m / 3) +
                    (m / 2 * (m * (n - m / 3) - n) + m / 6);
  else return n * n * (m - n / 3 - 1) / 2 + m + 2 * n / 3 +
              n * (n * (m - (1. / 3.) * n - 1) / 2 - m) + n / 6;
}

long long HSSHost::factor_flops_ref() const {
  long long f = 0;
  for (int i = 0; i < int(nodes.size()); i++) {
    auto& n = nodes[i];
    long long m = n.leaf() ? n.rows : nodes[n.ch0].u_rank + nodes[n.ch1].u_rank;
    if (!n.leaf()) {
      auto& a = nodes[n.ch0];
      auto& b = nodes[n.ch1];
      if (m)                                               // factor.hpp:68-81
        f += ref_gemm(a.u_rank, b.u_rank, b.v_rank, 1., 0.) +
             ref_gemm(b.u_rank, a.u_rank, a.v_rank, 1., 0.);
      if (i)                                               // :82-98
        f += ref_gemm(a.u_rank, n.v_rank, a.v_rank, 1., 0.) +
             ref_gemm(b.u_rank, n.v_rank, b.v_rank, 1., 0.);
    }
    if (i == 0) { f += ref_getrf(m, m); continue; }        // :104-107
    long long r = n.u_rank, k = m - r;
    if (k > 0) {
      f += ref_gemm(k, m, r, -1., 1.);                     // :116-121
      f += ref_gelqf(k, m) + ref_xxglq(m, m, std::min(k, m));  // :122-123
      f += 2 * ref_gemm(k, n.v_rank, m, 1., 0.) +          // :138-141
           ref_gemm(r, r, m, 1., 0.);
    }
  }
  return f;
}

long long HSSHost::solve_flops_ref() const {
  const long long s = 1;
  long long f = 0;
  for (int i = 0; i < int(nodes.size()); i++) {
    auto& n = nodes[i];
    long long m = n.leaf() ? n.rows : nodes[n.ch0].u_rank + nodes[n.ch1].u_rank;
    if (!n.leaf()) {
      auto& a = nodes[n.ch0];
      auto& b = nodes[n.ch1];
      f += ref_gemm(a.u_rank, s, b.v_rank, -1., 1.) +      // solve.hpp:92-100
           ref_gemm(b.u_rank, s, a.v_rank, -1., 1.);
      for (const HSSNode* c : {&a, &b}) {
        long long cm = c->u_rows, ck = cm - c->u_rank;
        if (ck > 0) {
          f += ref_gemm(cm, s, ck, 1., 0.) +               // :101-128
               ref_gemm(c->u_rank, s, cm, -1., 1.);
          f += ref_gemm(cm, s, cm, 1., 0.);                // bwd :209-224
        }
      }
    }
    if (i == 0) { f += 2 * m * m * s; continue; }          // :133-135
    long long r = n.u_rank, k = m - r;
    long long applyC = ref_gemm(n.v_rank, s, n.v_rows - n.v_rank, 1., 1.);
    if (k > 0) {
      f += ref_gemm(k, s, r, -1., 1.) + s * k * (k + 1);   // :158-165
      if (!n.leaf()) f += applyC + ref_gemm(n.v_rank, s, k, 1., 1.);  // :166-174
      else f += ref_gemm(n.v_rank, s, k, 1., 0.);          // :176-181
    } else if (!n.leaf()) f += applyC;                     // :185-187
  }
  return f;
}

long long HSSHost::qr_class_flops_ref(int h) const {
  long long f = 0;
  for (int q = hptr[h]; q < hptr[h + 1]; q++) {
    const int i = by_height[q];
    if (i == 0) continue;
    auto& n = nodes[i];
    long long m = n.leaf() ? n.rows : nodes[n.ch0].u_rank + nodes[n.ch1].u_rank;
    long long r = n.u_rank, k = m - r;
    if (k <= 0) continue;
    f += ref_gelqf(k, m) + ref_xxglq(m, m, std::min(k, m));
    f += 2 * ref_gemm(k, n.v_rank, m, 1., 0.) + ref_gemm(r, r, m, 1., 0.);
  }
  return f;
}

long long HSSHost::qr_class_flops_exec(int h) const {
  long double f = 0;
  for (int q = hptr[h]; q < hptr[h + 1]; q++) {
    const int i = by_height[q];
    if (i == 0) continue;
    auto& n = nodes[i];
    long long m = n.leaf() ? n.rows : nodes[n.ch0].u_rank + nodes[n.ch1].u_rank;
    long long k = m - n.u_rank, na = m + n.v_rank;
    for (long long j = 0; j < k; j++) f += 4.0L * (m - j) * (na - j - 1);
  }
  return (long long)f;
}

long long HSSHost::factor_flops_exec() const {
  // Householder QR of the m x k block applied to m x (k + r_v + r) columns,
  // no explicit Q: 2 * sum_j (m-j) * (cols right of j) * 2
  long double f = 0;
  for (int i = 1; i < int(nodes.size()); i++) {
    auto& n = nodes[i];
    long long m = n.leaf() ? n.rows : nodes[n.ch0].u_rank + nodes[n.ch1].u_rank;
    long long r = n.u_rank, k = m - r, na = m + n.v_rank;
    if (k <= 0) continue;
    f += 2.0L * k * m * r;  // W0 = Dp_bot - E W1
    for (long long j = 0; j < k; j++) f += 4.0L * (m - j) * (na - j - 1);
  }
  return (long long)f;
}

// ------------------------------------------------------------------ file IO
namespace {
struct Reader {
  std::ifstream f;
  std::uint64_t size = 0;   // file size: no record may claim more bytes than are left
  explicit Reader(const std::string& p) : f(p, std::ios::binary) {
    if (!f) throw std::runtime_error("cannot open " + p);
    f.seekg(0, std::ios::end);
    size = std::uint64_t(f.tellg());
    f.seekg(0, std::ios::beg);
  }
  std::uint64_t left() { return size - std::uint64_t(f.tellg()); }
  // every size field of the file is checked against what is left of the file
  // before anything is allocated or read (a truncated / corrupt dump must not
  // turn into a huge allocation or garbage generators)
  void need(std::uint64_t bytes, const char* what) {
    if (!f || bytes > left())
      throw std::runtime_error(std::string("corrupt or truncated HSS file (") + what + ")");
  }
  template <typename T> T get() {
    T v;
    f.read(reinterpret_cast<char*>(&v), sizeof(T));
    if (!f) throw std::runtime_error("truncated HSS file");
    return v;
  }
  void skip(std::size_t n) { need(n, "header"); f.seekg(std::streamoff(n), std::ios::cur); }
  // DenseMatrix record: int v[3], 40-byte object image, data
  // (reference src/dense/DenseMatrix.cpp:881-889)
  void dense(std::vector<double>& arena, int64_t& off, int& rows, int& cols) {
    skip(3 * sizeof(int));
    uint64_t img[5];
    need(sizeof(img), "DenseMatrix record");
    f.read(reinterpret_cast<char*>(img), sizeof(img));
    if (!f) throw std::runtime_error("truncated HSS file");
    if (img[2] > 0x7fffffffull || img[3] > 0x7fffffffull)
      throw std::runtime_error("corrupt HSS file (DenseMatrix dimensions)");
    rows = int(img[2]);
    cols = int(img[3]);
    std::size_t n = std::size_t(rows) * cols;
    need(std::uint64_t(n) * sizeof(double), "DenseMatrix data");
    off = n ? int64_t(arena.size()) : -1;
    if (n) {
      arena.resize(arena.size() + n);
      f.read(reinterpret_cast<char*>(arena.data() + off), n * sizeof(double));
      if (!f) throw std::runtime_error("truncated HSS file");
    }
  }
};

struct Writer {
  std::ofstream f;
  explicit Writer(const std::string& p)
      : f(p, std::ios::binary | std::ios::trunc) {
    if (!f) throw std::runtime_error("cannot open " + p);
  }
  template <typename T> void put(const T& v) {
    f.write(reinterpret_cast<const char*>(&v), sizeof(T));
  }
  void version() { int v[3] = {8, 0, 0}; f.write((const char*)v, sizeof(v)); }
  void dense(const double* d, int rows, int cols) {
    version();
    uint64_t img[5] = {0, 0, uint64_t(rows), uint64_t(cols),
                       uint64_t(std::max(rows, 1))};
    f.write(reinterpret_cast<const char*>(img), sizeof(img));
    if (rows * cols)
      f.write(reinterpret_cast<const char*>(d),
              sizeof(double) * std::size_t(rows) * cols);
  }
};
}  // namespace

HSSHost HSSHost::read_file(const std::string& path) {
  Reader r(path);
  HSSHost H;
  r.skip(3 * sizeof(int));  // version triple, HSSMatrix.cpp:476-478
  std::function<int(int)> rec = [&](int parent) -> int {
    HSSNode n;
    n.parent = parent;
    const uint64_t rows64 = r.get<uint64_t>(), cols64 = r.get<uint64_t>();
    if (rows64 > 0x7fffffffull || cols64 > 0x7fffffffull)
      throw std::runtime_error("corrupt HSS file (node dimensions)");
    n.rows = int(rows64);
    n.cols = int(cols64);
    if (parent >= 0 && (n.rows > H.nodes[parent].rows || n.cols > H.nodes[parent].cols))
      throw std::runtime_error("corrupt HSS file (child larger than its parent)");
    r.get<char>(); r.get<char>();   // U_state_, V_state_
    r.get<int>();                   // openmp_task_depth_
    r.get<char>();                  // active_
    n.u_rank = r.get<int>(); n.u_rows = r.get<int>();
    n.v_rank = r.get<int>(); n.v_rows = r.get<int>();
    int64_t off; int a, b;
    std::vector<double> scratch;
    r.dense(scratch, off, a, b);    // Asub_
    for (int w = 0; w < 2; w++) {   // U_, V_   (HSSBasisID.hpp:93-101)
      uint64_t ps = r.get<uint64_t>();
      // a basis has at most rows (leaf) / sum of the children's ranks <= rows entries
      if (ps > uint64_t(std::max(n.rows, n.cols)))
        throw std::runtime_error("corrupt HSS file (basis permutation longer than the node)");
      r.need(ps * 4, "basis permutation");
      std::vector<int32_t> ipiv(ps);
      if (ps) {
        r.f.read(reinterpret_cast<char*>(ipiv.data()), ps * 4);
        if (!r.f) throw std::runtime_error("truncated HSS file");
        for (uint64_t q = 0; q < ps; q++)
          if (ipiv[q] < 1 || uint64_t(ipiv[q]) > ps)
            throw std::runtime_error("corrupt HSS file (pivot index out of range)");
      }
      int64_t offE; int er, ec;
      r.dense(H.vals, offE, er, ec);
      int64_t offP = -1;
      if (ps) {
        auto g = ipiv_to_gather(ipiv.data(), int(ps));
        offP = int64_t(H.perms.size());
        H.perms.insert(H.perms.end(), g.begin(), g.end());
      }
      if (w == 0) { n.off_Eu = offE; n.off_Pu = offP; n.u_rows = int(ps); n.u_rank = ps ? ec : 0; }
      else        { n.off_Ev = offE; n.off_Pv = offP; n.v_rows = int(ps); n.v_rank = ps ? ec : 0; }
    }
    r.dense(H.vals, n.off_D, a, b);
    r.dense(H.vals, n.off_B01, a, b);
    r.dense(H.vals, n.off_B10, a, b);
    int nc = r.get<int>();
    if (H.nodes.size() > 0x3fffffffu) throw std::runtime_error("corrupt HSS file (too many nodes)");
    int me = int(H.nodes.size());
    H.nodes.push_back(n);
    if (nc == 2) {
      int c0 = rec(me);
      int c1 = rec(me);
      H.nodes[me].ch0 = c0;
      H.nodes[me].ch1 = c1;
    } else if (nc != 0) throw std::runtime_error("HSS node with 1 child");
    return me;
  };
  rec(-1);
  H.finalize();
  return H;
}

void HSSHost::write_file(const std::string& path) const {
  Writer w(path);
  w.version();
  std::function<void(int)> rec = [&](int i) {
    auto& n = nodes[i];
    w.put<uint64_t>(n.rows); w.put<uint64_t>(n.cols);
    w.put<char>('C'); w.put<char>('C');   // State::COMPRESSED
    w.put<int>(0); w.put<char>(1);
    w.put<int>(n.u_rank); w.put<int>(n.u_rows);
    w.put<int>(n.v_rank); w.put<int>(n.v_rows);
    w.dense(nullptr, 0, 0);               // Asub_
    for (int b = 0; b < 2; b++) {
      int rows = b ? n.v_rows : n.u_rows, rank = b ? n.v_rank : n.u_rank;
      int64_t offP = b ? n.off_Pv : n.off_Pu, offE = b ? n.off_Ev : n.off_Eu;
      if (offP < 0) rows = 0;
      w.put<uint64_t>(rows);
      if (rows) {
        auto ipiv = gather_to_ipiv(perms.data() + offP, rows);
        w.f.write(reinterpret_cast<const char*>(ipiv.data()), rows * 4);
      }
      w.dense(offE >= 0 ? vals.data() + offE : nullptr,
              rows ? rows - rank : 0, rows ? rank : 0);
    }
    w.dense(n.off_D >= 0 ? vals.data() + n.off_D : nullptr,
            n.off_D >= 0 ? n.rows : 0, n.off_D >= 0 ? n.cols : 0);
    if (!n.leaf()) {
      auto& a = nodes[n.ch0];
      auto& b = nodes[n.ch1];
      w.dense(n.off_B01 >= 0 ? vals.data() + n.off_B01 : nullptr,
              a.u_rank, b.v_rank);
      w.dense(n.off_B10 >= 0 ? vals.data() + n.off_B10 : nullptr,
              b.u_rank, a.v_rank);
      w.put<int>(2);
      rec(n.ch0);
      rec(n.ch1);
    } else {
      w.dense(nullptr, 0, 0);
      w.dense(nullptr, 0, 0);
      w.put<int>(0);
    }
  };
  rec(0);
}

HSSHost HSSHost::from_flat(int n_nodes, const int64_t* tab, const double* v,
                           int64_t n_vals, const int32_t* p, int64_t n_perms) {
  HSSHost H;
  H.nodes.resize(n_nodes);
  for (int i = 0; i < n_nodes; i++) {
    const int64_t* t = tab + 16 * std::size_t(i);
    auto& n = H.nodes[i];
    n.parent = int(t[0]); n.ch0 = int(t[1]); n.ch1 = int(t[2]);
    n.rows = int(t[3]); n.cols = int(t[4]);
    n.u_rows = int(t[5]); n.u_rank = int(t[6]);
    n.v_rows = int(t[7]); n.v_rank = int(t[8]);
    n.off_D = t[9]; n.off_Eu = t[10]; n.off_Ev = t[11];
    n.off_B01 = t[12]; n.off_B10 = t[13];
    n.off_Pu = t[14]; n.off_Pv = t[15];
    for (int64_t o : {n.off_D, n.off_Eu, n.off_Ev, n.off_B01, n.off_B10})
      if (o >= n_vals) throw std::invalid_argument("value offset out of range");
    for (int64_t o : {n.off_Pu, n.off_Pv})
      if (o >= n_perms) throw std::invalid_argument("perm offset out of range");
  }
  H.vals.assign(v, v + n_vals);
  H.perms.assign(p, p + n_perms);
  H.finalize();
  return H;
}

void HSSHost::print_info() const {
  // same line format as HSSMatrix::print_info (HSSMatrix.cpp:333-356)
  std::function<void(int)> rec = [&](int i) {
    auto& n = nodes[i];
    std::cout << "SEQ rank=0 b = [" << n.row_off << "," << n.row_off + n.rows
              << " x " << n.col_off << "," << n.col_off + n.cols
              << "]  U = " << n.u_rows << " x " << n.u_rank
              << " V = " << n.v_rows << " x " << n.v_rank
              << (n.leaf() ? " leaf" : " non-leaf") << std::endl;
    if (!n.leaf()) { rec(n.ch0); rec(n.ch1); }
  };
  rec(0);
}

}  // namespace sb200
