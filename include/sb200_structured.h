/*
 * strumpack_b200 -- C ABI of the B200-native HSS/BLR engine.
 *
 * Drop-in boundary: the SP_d_struct_* entry points below have
 * exactly the names, argument meaning, ownership and error behaviour of the
 * reference's C interface in
 *     reference src/structured/StructuredMatrix.h:46-85 (types),
 *     :118-607 (functions), glue src/structured/StructuredMatrixC.cpp:39-119
 * so a program written against StructuredMatrix.h links against
 * libstrumpack_b200.so unchanged.  All pointers passed to SP_* functions are
 * HOST pointers (reference doc/doxygen/pages/GPU_support.txt:24-26), matrices
 * are column-major with a leading dimension, `solve` overwrites B in place,
 * every function returns 0 on success and 1 after printing
 * "Operation failed: <what>" (StructuredMatrixC.cpp:107-119).
 *
 * Limits of this engine (the reference has none of them; its default max_rank is
 * 5000): a reduced HSS block (a leaf, or the sum of two children's ranks) of at
 * most 1600 rows, BLR tiles of at most 1024 rows; larger ones are refused with an
 * error.  Real arithmetic only (SP_d_*, and SP_s_* as a float boundary over fp64).
 *
 * The SB200_* entry points are engine extensions (device-resident operands,
 * generator import, statistics); they never change the meaning of SP_*.
 *
 * There is no CPU fallback behind this interface: every function that does
 * arithmetic launches sm_100a kernels and fails (return 1) without a GPU.
 */
#ifndef SB200_STRUCTURED_C_H
#define SB200_STRUCTURED_C_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* reference StructuredMatrix.h:46-55 */
typedef enum {
  SP_TYPE_HSS = 0,
  SP_TYPE_BLR,
  SP_TYPE_HODLR,
  SP_TYPE_HODBF,
  SP_TYPE_BUTTERFLY,
  SP_TYPE_LR,
  SP_TYPE_LOSSY,
  SP_TYPE_LOSSLESS
} SP_STRUCTURED_TYPE;

/* reference StructuredMatrix.h:68-75 */
typedef struct CSPOptions {
  SP_STRUCTURED_TYPE type;
  double rel_tol;
  double abs_tol;
  int leaf_size;
  int max_rank;
  int verbose;
} CSPOptions;

/* reference StructuredMatrix.h:85 */
typedef void* CSPStructMat;

/* ---- reference interface, double precision ------------------------------ */
/* StructuredMatrix.h:118  (defaults of StructuredOptions.hpp:106-162:
 * type BLR, rel_tol 1e-4, abs_tol 1e-10, leaf 128, max_rank 5000) */
void SP_d_struct_default_options(CSPOptions* opts);
/* StructuredMatrix.h:131 */
void SP_d_struct_destroy(CSPStructMat* S);
/* StructuredMatrix.h:143,153,164,175,186 */
int SP_d_struct_rows(const CSPStructMat S);
int SP_d_struct_cols(const CSPStructMat S);
long long int SP_d_struct_memory(const CSPStructMat S);
long long int SP_d_struct_nonzeros(const CSPStructMat S);
int SP_d_struct_rank(const CSPStructMat S);
/* StructuredMatrix.h:210  -- compress a host column-major dense matrix.
 * Supported types: SP_TYPE_HSS, SP_TYPE_BLR (others: return 1, like the
 * reference built without the optional back ends). */
int SP_d_struct_from_dense(CSPStructMat* S, int rows, int cols,
                           const double* A, int ldA, const CSPOptions* opts);
/* StructuredMatrix.h:244  -- compress from an element callback. The callback
 * runs on the host; the engine samples it into device buffers. */
int SP_d_struct_from_elements(CSPStructMat* S, int rows, int cols,
                              double A(int i, int j), const CSPOptions* opts);
/* StructuredMatrix.h:292  C = op(S) * B, trans in {N,n,T,t,C,c}; m = columns
 * of B and C. */
int SP_d_struct_mult(const CSPStructMat S, char trans, int m,
                     const double* B, int ldB, double* C, int ldC);
/* StructuredMatrix.h:340 */
int SP_d_struct_factor(CSPStructMat S);
/* StructuredMatrix.h:360  B <- S^{-1} B */
int SP_d_struct_solve(const CSPStructMat S, int nrhs, double* B, int ldB);
/* StructuredMatrix.h:395  S <- S + s*I (invalidates the factorization) */
int SP_d_struct_shift(CSPStructMat S, double s);

/* ---- reference interface, single precision (StructuredMatrix.h:103,130,158-259,
 * 295,354,387,464,510,569): float operands at the boundary, fp64 arithmetic
 * inside (operands are widened on the way in, results rounded on the way out;
 * memory/nonzeros report the fp64 object).  The complex variants SP_c_* / SP_z_*
 * are not provided. */
void SP_s_struct_default_options(CSPOptions* opts);
void SP_s_struct_destroy(CSPStructMat* S);
int SP_s_struct_rows(const CSPStructMat S);
int SP_s_struct_cols(const CSPStructMat S);
long long int SP_s_struct_memory(const CSPStructMat S);
long long int SP_s_struct_nonzeros(const CSPStructMat S);
int SP_s_struct_rank(const CSPStructMat S);
int SP_s_struct_from_dense(CSPStructMat* S, int rows, int cols, const float* A, int ldA,
                           const CSPOptions* opts);
int SP_s_struct_from_elements(CSPStructMat* S, int rows, int cols, float A(int i, int j),
                              const CSPOptions* opts);
int SP_s_struct_mult(const CSPStructMat S, char trans, int m, const float* B, int ldB,
                     float* C, int ldC);
int SP_s_struct_factor(CSPStructMat S);
int SP_s_struct_solve(const CSPStructMat S, int nrhs, float* B, int ldB);
int SP_s_struct_shift(CSPStructMat S, float s);

/* ---- engine extensions --------------------------------------------------- */

/* Kernel-matrix types for SB200_d_hss_from_kernel
 * (reference src/kernel/Kernel.hpp:333-357 GaussKernel, :370-396 Laplace). */
typedef enum {
  SB200_KERNEL_GAUSS = 0,    /* exp(-|x-y|_2^2 / (2 h^2)) + lambda [i==j] */
  SB200_KERNEL_LAPLACE = 1,  /* exp(-|x-y|_1 / h)         + lambda [i==j] */
  SB200_KERNEL_TOEPLITZ_INVDIST = 2 /* 1/(1+|i-j|), d=1, pts ignored:
                                       test/test_HSS_seq.cpp:69-79 */
} SB200_KERNEL_TYPE;

/* Mirrors HSSMatrix(kernel::Kernel&, opts), reference HSSMatrix.cpp:88-106:
 * clusters the n points (d x n, column-major, one point per column), reorders
 * them IN PLACE into the cluster ordering, writes the 0-based permutation to
 * perm (may be NULL; perm[new] = old) and compresses on the GPU.  All later
 * mult/solve calls act in the permuted ordering, as in the reference. */
int SB200_d_hss_from_kernel(CSPStructMat* S, int n, int d, double* pts,
                            int kernel_type, double h, double lambda,
                            const CSPOptions* opts, int* perm);
/* The same with the clustering algorithm of the reference's HSSOptions
 * (HSS::ClusteringAlgorithm, HSSOptions.hpp): 0 NATURAL (the given order,
 * recursive bisection), 1 TWO_MEANS (the reference's default: recursive 2-means,
 * ragged leaves), 2 KD_TREE (median bisection; what SB200_d_hss_from_kernel
 * uses).  3 PCA and 4 COBBLE are not implemented and refused. */
int SB200_d_hss_from_kernel_ex(CSPStructMat* S, int n, int d, double* pts,
                               int kernel_type, double h, double lambda,
                               const CSPOptions* opts, int* perm, int clustering);


/* HSSMatrix::compress(Amult, Aelem, opts) (reference src/HSS/HSSMatrix.cpp:173-186;
 * what FrontHSS calls, src/sparse/fronts/FrontHSS.cpp:385): construction from a
 * callback that fills whole sub-blocks, the reference's elem_t
 * (std::function<void(I, J, B)>, HSSMatrix.hpp:68-70): elem must write
 * B[a + b*ldB] = A(I[a], J[b]) for the 0-based index lists I (nI) and J (nJ).
 * The engine's compressor is a sampled interpolative decomposition: it asks
 * only for the O(n * samples) entries it needs (no Amult, no n^2 buffer; for
 * n <= 8192 the whole complement of every node is used and the ID is exact to
 * the tolerance).  The callback runs on the host, the factorizations of the
 * sampled blocks on the device. */
typedef void (*SB200ElemBlockFn)(int nI, const int* I, int nJ, const int* J, double* B, int ldB,
                                 void* user);
int SB200_d_hss_from_element_blocks(CSPStructMat* S, int n, SB200ElemBlockFn elem, void* user,
                                    const CSPOptions* opts);
/* The same with (a) a cluster tree given by the caller instead of recursive
 * bisection down to leaf_size -- pre-order arrays of tree_nodes entries: rows
 * and number of children (0 or 2) of every node; what HSSMatrix(const
 * structured::ClusterTree&, opts) fixes in the reference (HSSMatrix.cpp:71-82);
 * tree_nodes = 0: none -- and (b) coordinates (d x n, column-major) of the
 * unknowns: sampled columns are then chosen by geometric distance
 * (HSSMatrix::compress_with_coordinates, HSSMatrix.hpp:302); d = 0: none. */
int SB200_d_hss_from_element_blocks_ex(CSPStructMat* S, int n, SB200ElemBlockFn elem, void* user,
                                       const CSPOptions* opts, int tree_nodes, const int* tree_sizes,
                                       const int* tree_nchild, int d, const double* coords);
/* HSS from a dense matrix on a given cluster tree (see above). */
int SB200_d_hss_from_dense_tree(CSPStructMat* S, int n, const double* A, int ldA, const CSPOptions* opts,
                                int tree_nodes, const int* tree_sizes, const int* tree_nchild);
/* The HSS tree: nodes x 10 int64 in pre-order (parent, child 0, child 1, rows,
 * cols, row offset, column offset, U rank, V rank, height).  Returns the number
 * of nodes (-1 on error); out may be NULL to query it. */
int SB200_d_hss_node_table(const CSPStructMat S, long long int* out);


/* BLRMatrix<double>::compress_and_factor(A, weak admissibility, opts)
 * (reference src/BLR/BLRMatrix.cpp:113-241, RL variant; tiles from
 * ClusterTree(n).refine(leaf_size) as in test/test_BLR_seq.cpp:136-145;
 * pivot_threshold as BLROptions::pivot_threshold, DenseTile.cpp:111-117).
 * The result supports SP_d_struct_solve (BLRMatrix::solve, BLRMatrix.hpp:118).
 * A BLR matrix built by SP_d_struct_from_dense (compress only,
 * StructuredMatrix.cpp:78-98) supports SP_d_struct_mult instead. */
int SB200_d_blr_compress_and_factor(CSPStructMat* S, int n, const double* A,
                                    int ldA, const CSPOptions* opts,
                                    double pivot_threshold);
/* Same with the dense matrix already resident on the device (dA: device
 * pointer); the engine keeps its own working copy. */
int SB200_d_blr_compress_and_factor_device(CSPStructMat* S, int n, const double* dA,
                                           int ldA, const CSPOptions* opts,
                                           double pivot_threshold);
/* BLRMatrix<double>::construct_and_partial_factor(A11, A12, A21, A22, B11, B12,
 * B21, tiles1, tiles2, admissible, opts) (reference src/BLR/BLRMatrix.cpp:739-1037,
 * RL variant, weak admissibility; caller src/sparse/fronts/FrontBLR.cpp:429-433):
 * the front [A11 A12; A21 A22] (n1 + n2 rows, tiles ClusterTree(n1).refine(leaf)
 * and ClusterTree(n2).refine(leaf)) is eliminated over the tiles of A11 only.
 * S then holds F11 = LU(A11) in BLR form and the compressed F12 = L^{-1} P A12,
 * F21 = A21 U^{-1}; A22 is overwritten IN PLACE with the dense Schur complement
 * A22 - A21 A11^{-1} A12 (what extend_add passes to the parent front).  Host
 * pointers; the _device form takes device pointers. */
int SB200_d_blr_partial_factor(CSPStructMat* S, int n1, int n2, const double* A11, int ld11,
                               const double* A12, int ld12, const double* A21, int ld21,
                               double* A22, int ld22, const CSPOptions* opts,
                               double pivot_threshold);
int SB200_d_blr_partial_factor_device(CSPStructMat* S, int n1, int n2, const double* dA11, int ld11,
                                      const double* dA12, int ld12, const double* dA21, int ld21,
                                      double* dA22, int ld22, const CSPOptions* opts,
                                      double pivot_threshold);
/* The same two factorizations with the BLROptions members that CSPOptions does
 * not carry (reference src/BLR/BLROptions.hpp:81-142):
 *   factor_algorithm  BLRFactorAlgorithm by its enum value (BLROptions.hpp:65;
 *                     BLRMatrix.cpp:170-235, 846-1013): 0 COLWISE, 1 RL, 2 LL,
 *                     3 COMB, 4 STAR.  RL and LL are schedules of the same tile
 *                     kernels (LL: block row / column i receives the updates of
 *                     all earlier steps right before it is factored; identical
 *                     factors); COLWISE, COMB and STAR run as RL.
 *   admissible        the reference's adm_t (BLRMatrix.hpp:78, strong
 *                     admissibility: FrontBLR.cpp:262-281): nb x nb column-major
 *                     over the tiles of the eliminated block, nonzero = the tile
 *                     may be compressed, 0 = it stays a DenseTile
 *                     (BLRMatrix.cpp:146-147).  NULL = weak admissibility. */
typedef struct SB200BLRParams {
  double pivot_threshold;
  int factor_algorithm;
  const int* admissible;
  int n_admissible;
  /* tile partition given by the caller (the reference's tiles1 / tiles2: leaf
   * sizes of the separator / update cluster trees); NULL / 0: recursive
   * bisection down to leaf_size.  tiles1 must add up to n (or n1), tiles2 to n2. */
  const int* tiles1;
  int n_tiles1;
  const int* tiles2;
  int n_tiles2;
} SB200BLRParams;
int SB200_d_blr_compress_and_factor_ex(CSPStructMat* S, int n, const double* A, int ldA,
                                       const CSPOptions* opts, const SB200BLRParams* params);
int SB200_d_blr_partial_factor_ex(CSPStructMat* S, int n1, int n2, const double* A11, int ld11,
                                  const double* A12, int ld12, const double* A21, int ld21,
                                  double* A22, int ld22, const CSPOptions* opts,
                                  const SB200BLRParams* params);
/* The extract_t forms (reference BLRMatrix::compress / compress_and_factor(const
 * extract_t& Aelem, admissible, opts), BLRMatrix.hpp:104-112, and
 * construct_and_partial_factor(n1, n2, A11, A12, A21, A22 extractors, ...),
 * BLRMatrix.hpp:223-232): the matrix (resp. the whole front, indices
 * 0 .. n1+n2-1) is defined by the block callback, which is called once per
 * tile pair.  factor = 0: compress only (supports mult), 1: compress_and_factor.
 * The partial form writes the Schur complement to A22 (n2 x n2, may be NULL). */
int SB200_d_blr_from_element_blocks(CSPStructMat* S, int n, SB200ElemBlockFn elem, void* user,
                                    const CSPOptions* opts, const SB200BLRParams* params, int factor);
int SB200_d_blr_partial_factor_element_blocks(CSPStructMat* S, int n1, int n2, SB200ElemBlockFn elem,
                                              void* user, double* A22, int ld22,
                                              const CSPOptions* opts, const SB200BLRParams* params);
/* n1 of a partially factored front (rows of S otherwise). */
int SB200_d_blr_sep_rows(const CSPStructMat S);
/* The two halves of the front solve (FrontBLR::fwd_solve_node / bwd_solve_node,
 * FrontBLR.cpp:525-568) on B = [b_sep; b_upd] ((n1 + n2) x nrhs, host, in place):
 * forward:  b_sep <- L11^{-1} P b_sep,  b_upd <- b_upd - F21 b_sep
 *           (laswp + BLRMatrix::trsmLNU_gemm, BLRMatrix.cpp:1552-1608)
 * backward: b_sep <- U11^{-1} (b_sep - F12 b_upd)
 *           (BLRMatrix::gemm_trsmUNN, BLRMatrix.cpp:1610-1665)
 * (b_upd is not touched by backward; the caller solves the Schur system in between). */
int SB200_d_blr_partial_forward_solve(const CSPStructMat S, int nrhs, double* B, int ldB);
int SB200_d_blr_partial_backward_solve(const CSPStructMat S, int nrhs, double* Y, int ldY);
int SB200_d_blr_tiles(const CSPStructMat S);
/* Number of off-diagonal tiles that did not compress to rank <= min(m,n)/2 and
 * are kept dense (DenseTile in the reference, src/BLR/BLRMatrix.cpp:563-570). */
int SB200_d_blr_dense_tiles(const CSPStructMat S);

/* Reads a reference HSS dump (HSSMatrix<double>::write, reference
 * HSSMatrix.cpp:438-486) and uploads its generators. */
int SB200_d_hss_read(CSPStructMat* S, const char* path);
/* Writes the generators in that same format (HSSMatrix.cpp:474-480). */
int SB200_d_hss_write(const CSPStructMat S, const char* path);

/* Import generators from flat host arrays (pre-order node numbering, root=0).
 * node_tab: n_nodes x SB200_NODE_FIELDS int64, see sb200 DESIGN.md section 3:
 *   [0] parent [1] child0 [2] child1 [3] rows [4] cols
 *   [5] U_rows [6] U_rank [7] V_rows [8] V_rank
 *   [9] off_D [10] off_Eu [11] off_Ev [12] off_B01 [13] off_B10  (into vals,
 *        column-major blocks, -1 when absent)
 *   [14] off_Pu [15] off_Pv   (into perms; 0-based GATHER index of length
 *        U_rows / V_rows such that (P^T b)[i] = b[perm[i]])
 */
#define SB200_NODE_FIELDS 16
int SB200_d_hss_from_generators(CSPStructMat* S, int n_nodes,
                                const int64_t* node_tab, const double* vals,
                                int64_t n_vals, const int32_t* perms,
                                int64_t n_perms);

/* Device-resident operands: dB/dC are DEVICE pointers, work is queued on
 * `stream` (a cudaStream_t, 0 = legacy default stream) and NOT synchronised. */
int SB200_d_struct_mult_device(const CSPStructMat S, char trans, int m,
                               const double* dB, int ldB, double* dC, int ldC,
                               void* stream);
int SB200_d_struct_factor_device(CSPStructMat S, void* stream);
int SB200_d_struct_solve_device(const CSPStructMat S, int nrhs, double* dB,
                                int ldB, void* stream);

/* apply_HSS(op, A, B, beta, C): C = op(S) B + beta C (reference
 * src/HSS/HSSMatrix.cpp:419-435, free function HSSMatrix.hpp:705-713).  Host
 * pointers; the _device form takes device pointers and a stream. */
int SB200_d_hss_apply(const CSPStructMat S, char trans, int m, const double* B,
                      int ldB, double beta, double* C, int ldC);
int SB200_d_hss_apply_device(const CSPStructMat S, char trans, int m,
                             const double* dB, int ldB, double beta, double* dC,
                             int ldC, void* stream);
/* HSSMatrix::extract(I, J) / extract_add(I, J, B) / get(i, j) (reference
 * src/HSS/HSSMatrix.extract.hpp:8-188, test/test_HSS_seq.cpp:204-233): the
 * nI x nJ sub-block H(I, J) for arbitrary 0-based index lists, written to
 * (add = 0) or added to (add = 1) the host matrix B (column-major, ld ldB). */
int SB200_d_hss_extract(const CSPStructMat S, int nI, const int* I, int nJ,
                        const int* J, double* B, int ldB, int add);
/* HSSMatrix::forward_solve / backward_solve (reference HSSMatrix.hpp:360-376,
 * HSSMatrix.solve.hpp:52-66): solve = forward (bottom-up elimination of the
 * right-hand side and the root solve) followed by backward (top-down
 * reconstruction).  The state passed between the two (WorkSolve in the
 * reference) is kept inside S.  forward reads B, backward writes X. */
int SB200_d_hss_forward_solve(const CSPStructMat S, int nrhs, const double* B, int ldB);
int SB200_d_hss_backward_solve(const CSPStructMat S, int nrhs, double* X, int ldX);
int SB200_d_hss_forward_solve_device(const CSPStructMat S, int nrhs, double* dB,
                                     int ldB, void* stream);
int SB200_d_hss_backward_solve_device(const CSPStructMat S, int nrhs, double* dX,
                                      int ldX, void* stream);

/* ---- Schur complement of the (0,0) block: what the reference's HSS fronts
 * use (src/sparse/fronts/FrontHSS.cpp:385-410 factor, :150-222 sampling,
 * :452-462 / :487-495 solve).  With H = [H00 H01; H10 H11] split at the root's
 * children (rows0 + rows1 = rows):
 *   S = H11 - H10 H00^{-1} H01 = H11 - Theta Vhat^H Phi^H.
 * partial_factor   HSSMatrix::partial_factor          (HSSMatrix.factor.hpp:44-50)
 * schur_sizes      out[0..6] = rows1, cols1, r_v(child 0), m0, r_v(child 1),
 *                  r_u(child 1), rows0; m0 = size of child 0's reduced block D0
 * schur_update     HSSMatrix::Schur_update             (HSSMatrix.Schur.hpp:40-59)
 *                  Theta rows1 x r_v0, DUB01 m0 x r_v1, Phi cols1 x m0 (host,
 *                  any may be NULL; device copies are kept inside S)
 * vhat             child(0)->ULV().Vhat(), m0 x r_v0   (HSSExtra.hpp:197-212)
 * schur_product_direct    Sr = S R, Sc = S^H R, R rows1 x c
 *                                                      (HSSMatrix.Schur.hpp:73-137)
 *                  Theta / DUB01 / Phi = NULL: the ones of the last schur_update
 * schur_product_indirect  Sr = Sr1 - H10 R0 - H10 H00^{-1} H01 R1 (and the
 *                  transposed analogue for Sc) from Sr1 = (H [R0; R1])_1,
 *                  Sc1 = (H^H [R0; R1])_1               (HSSMatrix.Schur.hpp:139-215)
 * partial_forward_solve   child(0)->forward_solve(w, b0, partial = true): keeps
 *                  x = D0^{-1} f inside S and returns reduced_rhs (r_v0 x nrhs)
 *                                                      (HSSMatrix.solve.hpp:133-152)
 * partial_x        get (set = 0) / overwrite (set = 1) that x (m0 x nrhs), as
 *                  FrontHSS::bwd_solve_node updates it with Phi^H y_upd
 * partial_backward_solve  child(0)->backward_solve(w, x0)
 * The basis of the reduced space (hence Vhat, DUB01, Phi, x individually)
 * depends on the orthogonal factors chosen by the ULV elimination; Theta,
 * Vhat^H DUB01, Vhat^H Phi^H, reduced_rhs, Sr, Sc and x0 do not. */
int SB200_d_hss_partial_factor(CSPStructMat S);
int SB200_d_hss_schur_sizes(const CSPStructMat S, int* out);
int SB200_d_hss_schur_update(const CSPStructMat S, double* Theta, int ldT, double* DUB01,
                             int ldD, double* Phi, int ldP);
int SB200_d_hss_vhat(const CSPStructMat S, double* Vhat, int ldV);
int SB200_d_hss_schur_product_direct(const CSPStructMat S, const double* Theta, int ldT,
                                     const double* DUB01, int ldD, const double* Phi, int ldP,
                                     int c, const double* R, int ldR, double* Sr, int ldSr,
                                     double* Sc, int ldSc);
int SB200_d_hss_schur_product_indirect(const CSPStructMat S, const double* DUB01, int ldD, int c,
                                       const double* R0, int ldR0, const double* R1, int ldR1,
                                       const double* Sr1, int ldSr1, const double* Sc1, int ldSc1,
                                       double* Sr, int ldSr, double* Sc, int ldSc);
/* device-resident forms (device pointers, queued on `stream`, not synchronised;
 * leading dimensions >= 1 also for empty matrices) */
int SB200_d_hss_partial_factor_device(CSPStructMat S, void* stream);
int SB200_d_hss_schur_update_device(const CSPStructMat S, double* dTheta, int ldT, double* dDUB01,
                                    int ldD, double* dPhi, int ldP, void* stream);
int SB200_d_hss_schur_product_direct_device(const CSPStructMat S, const double* dTheta, int ldT,
                                            const double* dDUB01, int ldD, const double* dPhi,
                                            int ldP, int c, const double* dR, int ldR, double* dSr,
                                            int ldSr, double* dSc, int ldSc, void* stream);
int SB200_d_hss_partial_forward_solve(const CSPStructMat S, int nrhs, const double* B0, int ldB,
                                      double* reduced_rhs, int ldR);
int SB200_d_hss_partial_x(const CSPStructMat S, int nrhs, double* X, int ldX, int set);
int SB200_d_hss_partial_backward_solve(const CSPStructMat S, int nrhs, double* X0, int ldX);

/* ---- subtree sharding over the GPUs of one node -------------
 * Every rank holds a handle on the same matrix and calls
 * SB200_d_hss_set_partition(S, nparts, part): it then owns the subtree of the
 * part-th node at depth log2(nparts) (row range from SB200_d_hss_owned_range)
 * while the nparts-1 nodes above the cut are replicated.  Each operation needs
 * ONE small exchange, done by the caller between _begin and _end as an
 * all-gather (NCCL) of `SB200_d_hss_dist_sizes(S, nrhs, out)[op]` doubles per
 * rank: out[0] apply, out[1] factor, out[2] solve.  dSend/dRecv, dB/dC are
 * device pointers; x/b/y are full-length vectors of which only the owned rows
 * are read/written.  This replaces the reference's MPI/BLACS subtree mapping
 * (src/HSS/HSSMatrixMPI.cpp:317-345, pgemr2d moves HSSMatrixMPI.factor.hpp:63-66). */
/* The same sharded operations with the exchange INSIDE the engine: local sweep,
 * one ncclAllGather on `stream`, replicated top and the sweep back down are queued
 * back to back by one call (and replayed as one CUDA graph from the second call on).
 * NCCL is bound at run time (dlopen; the copy a host framework has loaded is
 * shared).  Rank 0 obtains a 128-byte unique id (SB200_nccl_unique_id) and hands
 * it to the other ranks by any means; every rank then calls SB200_d_hss_dist_init
 * (partition + communicator).  Operands are DEVICE pointers, full length on every
 * rank; only the owned rows are read / written. */
int SB200_nccl_unique_id(char* out128);
int SB200_d_hss_dist_init(CSPStructMat S, int nparts, int part, const char* unique_id128);
int SB200_d_hss_dist_mult(const CSPStructMat S, char trans, int m, const double* dB, int ldB, double* dC,
                          int ldC, void* stream);
int SB200_d_hss_dist_factor(CSPStructMat S, void* stream);
int SB200_d_hss_dist_solve(const CSPStructMat S, int nrhs, double* dB, int ldB, void* stream);

int SB200_d_hss_set_partition(CSPStructMat S, int nparts, int part);
int SB200_d_hss_owned_range(const CSPStructMat S, int* lo, int* hi);
int SB200_d_hss_dist_sizes(const CSPStructMat S, int nrhs, long long int* out);
int SB200_d_hss_dist_mult_begin(const CSPStructMat S, char trans, int m,
                                const double* dB, int ldB, double* dSend,
                                void* stream);
int SB200_d_hss_dist_mult_end(const CSPStructMat S, char trans, int m,
                              const double* dB, int ldB, double* dC, int ldC,
                              const double* dRecv, void* stream);
int SB200_d_hss_dist_factor_begin(CSPStructMat S, double* dSend, void* stream);
int SB200_d_hss_dist_factor_end(CSPStructMat S, const double* dRecv, void* stream);
int SB200_d_hss_dist_solve_begin(const CSPStructMat S, int nrhs, double* dB,
                                 int ldB, double* dSend, void* stream);
int SB200_d_hss_dist_solve_end(const CSPStructMat S, int nrhs, double* dB,
                               int ldB, const double* dRecv, void* stream);

/* Statistics (HSSMatrix::levels, factor_nonzeros; reference
 * HSSMatrix.cpp:326-332, HSSMatrixBase.cpp:74). */
int SB200_d_struct_levels(const CSPStructMat S);
long long int SB200_d_struct_factor_nonzeros(const CSPStructMat S);
/* Download the ULV factors (reference accessor HSSMatrix::ULV(),
 * HSSMatrix.hpp:497; layout DESIGN.md 3): the arena of factor blocks and the
 * block-reflector T arena; sizes[0..1] receive the two lengths (their sum is
 * SB200_d_struct_factor_nonzeros), either pointer may be NULL to query sizes
 * only. */
int SB200_d_struct_ulv_data(const CSPStructMat S, double* factors, double* tfactors,
                            long long int* sizes);
/* Algorithmic flop counts in the reference's own accounting:
 * which = 0: apply with 1 rhs (2*nnz terms of HSSMatrix.apply.hpp),
 *         1: ULV factor  (params::ULV_factor_flops formula, factor.hpp:78-141)
 *         2: ULV solve, 1 rhs (params::hss_solve_flops, solve.hpp:94-223)
 *         3: flops the engine actually executes in factor (no explicit Q)
 *         4: the leaf-class Householder-QR launch (dominant kernel), reference
 *            accounting: LQ_flops + the three Q GEMMs (factor.hpp:122-141)
 *         5: the same launch, flops actually executed */
long long int SB200_d_struct_flops(const CSPStructMat S, int which);
/* Live kernel timing: when enabled, factor records CUDA events around the
 * leaf-class QR launch on the launching stream; kernel_ms(S, 0) returns the
 * duration (ms) of that launch in the last factor call. */
int SB200_d_struct_set_profile(CSPStructMat S, int on);
double SB200_d_struct_kernel_ms(const CSPStructMat S, int which);
/* Number of kernels launched by this object since creation. */
long long int SB200_d_struct_launches(const CSPStructMat S);
/* H.print_info() equivalent to stdout (HSSMatrix.cpp:333-356). */
int SB200_d_struct_print_info(const CSPStructMat S);
/* Dense reconstruction into a host buffer (HSSMatrix::dense, BLRMatrix::dense of
 * a compressed, unfactored BLR matrix). */
int SB200_d_struct_dense(const CSPStructMat S, double* A, int ldA);

/* Host-only helpers (no GPU needed): parse a reference HSS dump and report
 * out[0..9] = rows, cols, n_nodes, levels, max rank, nonzeros (reference
 * accounting), apply flops, ULV factor flops, ULV solve flops (1 rhs; both in
 * the reference's accounting), flops the engine executes in factor. */
int SB200_d_hss_file_info(const char* path, long long int* out);
/* Read a reference HSS dump and write it back (round trip of the format). */
int SB200_d_hss_file_copy(const char* in_path, const char* out_path);

/* Library identification: returns "strumpack_b200 <ver> sm_100a". */
/* Device-resident extend-add (front assembly): for each of nf parent fronts,
 * F(I[y], I[x]) += CB(y, x) for the contribution blocks of its left and right
 * child.  The parent is stored as its four column-major blocks F11 (d1 x d1, ld
 * d1), F12 (d1 x d2, ld d1), F21 (d2 x d1, ld d2), F22 (d2 x d2, ld d2); CB1 /
 * CB2 are dCB x dCB (ld dCB) or NULL, I1 / I2 the child's update indices in the
 * parent's numbering (0 .. d1+d2-1, no repeats).  Every pointer, including
 * d_fronts itself, is a DEVICE pointer; the work is queued on `stream`.
 * Replaces extend_add_kernel / AssembleData of src/sparse/fronts/FrontCUDA.cu:60-148
 * and the extend-add loops of FrontBLR.cpp:338-403. */
typedef struct {
  double *F11, *F12, *F21, *F22;
  int d1, d2;
  const double* CB1;
  const int* I1;
  int dCB1;
  const double* CB2;
  const int* I2;
  int dCB2;
} SB200FrontAssemble;
int SB200_d_front_extend_add_device(int nf, const SB200FrontAssemble* d_fronts, int max_dCB, void* stream);

/* The roofline denominator of the factor kernels, measured on the current device
 * when called (about 10 ms): TFLOP/s of a register-resident stream of
 * mma.sync.m8n8k4.f64 (fp64 has no tcgen05 kind; this is the fp64 tensor pipe). */
double SB200_fp64_dmma_peak_tflops(void);

/* Test / microbenchmark hook for the ULV leaf QR kernels, outside any tree:
 * `count` copies of the column-major m x naug block A (m <= 256) get the
 * blocked Householder QR of their first k columns, reflectors applied to all
 * naug columns.  variant 0: right-looking ulv_qr_kernel, 1: left-looking
 * warp-specialised ulv_qr3_kernel.  out (m x naug) / T (16 x k, one 16 x 16
 * block-reflector factor per 16-column panel) receive the last copy's result,
 * ms the best kernel time of `reps` launches (CUDA events). */
int SB200_debug_qr_batch(int m, int k, int naug, int count, const double* A, double* out, double* T,
                         int variant, int reps, float* ms);

const char* SB200_version(void);

#ifdef __cplusplus
}
#endif
#endif
