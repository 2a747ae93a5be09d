// strumpack_b200 -- HSS construction on the GPU.
//
// Produces the generators consumed by HSSEngine.  Replaces (reference, CPU):
//   HSSMatrix(const DenseM_t&, opts) / compress   src/HSS/HSSMatrix.cpp:49-54,148-161
//   HSSMatrix(kernel::Kernel&, opts)              src/HSS/HSSMatrix.cpp:88-106
//   compress_kernel / ANN sampling                src/HSS/HSSMatrix.compress_kernel.hpp:39-293
// The algorithm here is NOT the reference's randomized sampling: see
// hss_compress.cu (sampled-column interpolative decomposition, batched per
// height class on the device).
#pragma once
#include "hss_tree.hpp"
#include "sb200_common.cuh"

namespace sb200 {

struct CompressOptions {
  double rel_tol = 1e-2, abs_tol = 1e-8;  // HSSOptions defaults (HSSOptions.hpp:465-490)
  int leaf_size = 512;
  int max_rank = 50000;
  int verbose = 0;
  // sampled-ID parameters: points per near-field / coarse-scale stratum
  // (the role HSSOptions::d0/ann_number play in the reference)
  int sample_near = 96, sample_far = 128;
  // scale every sampled column by sqrt(number of complement points it stands
  // for) before the interpolative decomposition.  -1 = automatic: on for
  // algebraically decaying kernels (1/(1+|i-j|), sampled dense input), off for
  // the exponentially decaying Gauss / Laplace kernels whose far field is
  // numerically zero (there the weights only move the relative stopping
  // threshold).  Measured: profiles/r1b_compress_accuracy.txt.
  // env SB200_COMPRESS_WEIGHTED=0/1 overrides.
  int weighted_samples = -1;
  // dense / block-callback input: use the whole complement of every node instead
  // of a sample (exact interpolative decomposition, O(n^2) entries per level).
  // -1 = automatic (n <= 8192)
  int full_complement = -1;
};

// host callback that fills the sub-block B (nI x nJ, column-major, ld ldB) =
// A(I, J) for 0-based index lists: the reference's elem_t
// (HSSMatrix.hpp:68-70, compress(Amult, Aelem, opts) HSSMatrix.cpp:173-186)
using BlockElemFn = void (*)(int nI, const int* I, int nJ, const int* J, double* B, int ldB, void* user);

// a cluster tree given by the caller (structured::ClusterTree), pre-order
struct GivenTree {
  int nnodes = 0;
  const int* sizes = nullptr;    // rows of every node
  const int* nchild = nullptr;   // 0 or 2
};

// A: host column-major rows x cols
// keep_dA: if given, receives the packed device copy of A (ld = rows) that the
// construction made, so that the caller can verify the result without a second
// host-to-device copy
HSSHost compress_dense(int rows, int cols, const double* A, int ldA,
                       const CompressOptions& o, DevBuf<double>* keep_dA = nullptr,
                       const GivenTree* tree = nullptr);
// element callback evaluated on the host
HSSHost compress_elements(int rows, int cols, double (*A)(int, int),
                          const CompressOptions& o);
// entries from a block callback: only the sampled blocks are evaluated
// (O(n * samples) entries), so n is not limited by a dense n^2 buffer
// tree: partition to use instead of recursive bisection down to leaf_size;
// coords (d x n, column-major): sample selection by geometric distance
// (HSSMatrix::compress_with_coordinates) instead of index distance
HSSHost compress_element_blocks(int n, BlockElemFn elem, void* user, const CompressOptions& o,
                                const GivenTree* tree = nullptr, int d = 0, const double* coords = nullptr);
// kernel matrix on n points (d x n, column-major); pts is reordered in place,
// perm[new] = old (may be null). kernel_type: SB200_KERNEL_TYPE.  clustering: the
// reference's HSS::ClusteringAlgorithm (0 natural order, 1 recursive 2-means,
// 2 kd-tree; PCA / COBBLE are refused).
HSSHost compress_kernel(int n, int d, double* pts, int kernel_type, double h,
                        double lambda, const CompressOptions& o, int* perm, int clustering = 2);

}  // namespace sb200
