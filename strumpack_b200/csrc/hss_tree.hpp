// strumpack_b200 -- host-side HSS cluster tree and generator container.
//
// The reference keeps one heap object per HSS node (reference
// src/HSS/HSSMatrix.hpp:520-521: U_, V_, D_, B01_, B10_; children in
// HSSMatrixBase.hpp:330 ch_).  Here the tree is a flat table: nodes are
// numbered in pre-order (root = 0), all generator blocks live in ONE double
// arena and ONE int arena, and nodes are grouped into "height classes"
// (height 0 = leaves) so that every sweep of apply / ULV factor / ULV solve is
// a handful of batched launches over a ragged list instead of a recursion.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace sb200 {

struct HSSNode {
  int parent = -1, ch0 = -1, ch1 = -1;
  int rows = 0, cols = 0;         // size of the block this node represents
  int row_off = 0, col_off = 0;   // offset of the block in the full matrix
  int u_rows = 0, u_rank = 0;     // U = P_u [I; E_u] is u_rows x u_rank
  int v_rows = 0, v_rank = 0;
  int height = 0, depth = 0;
  // offsets into HSSHost::vals (column-major blocks), -1 = absent
  int64_t off_D = -1, off_Eu = -1, off_Ev = -1, off_B01 = -1, off_B10 = -1;
  // offsets into HSSHost::perms: 0-based gather index, (P^T b)[i] = b[p[i]]
  int64_t off_Pu = -1, off_Pv = -1;
  bool leaf() const { return ch0 < 0; }
};

struct HSSHost {
  std::vector<HSSNode> nodes;     // pre-order, root = 0
  std::vector<double> vals;       // all generator blocks
  std::vector<int32_t> perms;     // all gather permutations
  // nodes sorted by height: class h = by_height[hptr[h] .. hptr[h+1])
  std::vector<int> by_height, hptr;

  int rows() const { return nodes.empty() ? 0 : nodes[0].rows; }
  int cols() const { return nodes.empty() ? 0 : nodes[0].cols; }
  int levels() const { return int(hptr.size()) - 1; }
  int max_rank() const;
  // reference accounting (HSSMatrix.cpp:316-323 minus sizeof(*this))
  long long nonzeros() const;
  long long memory_bytes() const;

  // Fill row_off/col_off/height/depth and the height classes; validate sizes.
  // Throws std::invalid_argument on inconsistent generators.
  void finalize();

  // Flop counts in the reference's accounting.
  long long apply_flops() const;     // 1 rhs, HSSMatrix.apply.hpp tallies
  long long factor_flops_ref() const;  // params::ULV_factor_flops formula
  long long solve_flops_ref() const;   // params::hss_solve_flops, 1 rhs
  long long factor_flops_exec() const; // what the engine really executes
  // The batched Householder-QR launch over height class h (the dominant
  // kernel for h = 0): flops in the reference's accounting (LQ_flops + the
  // three Q-GEMMs, factor.hpp:122-141) and as executed by the engine.
  long long qr_class_flops_ref(int h) const;
  long long qr_class_flops_exec(int h) const;

  // Reference dump format, HSSMatrix<double>::write/read
  // (reference src/HSS/HSSMatrix.cpp:438-510).
  static HSSHost read_file(const std::string& path);
  void write_file(const std::string& path) const;

  // From flat arrays (include/sb200_structured.h SB200_d_hss_from_generators)
  static HSSHost from_flat(int n_nodes, const int64_t* tab, const double* vals,
                           int64_t n_vals, const int32_t* perms,
                           int64_t n_perms);

  void print_info() const;   // HSSMatrix.cpp:333-356
};

// LAPACK ipiv (1-based sequential swaps) <-> gather index
std::vector<int32_t> ipiv_to_gather(const int32_t* ipiv, int n);
std::vector<int32_t> gather_to_ipiv(const int32_t* g, int n);

}  // namespace sb200
