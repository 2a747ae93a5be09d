/* Oracle build configuration (TEST INFRASTRUCTURE ONLY).
 * Hand-written equivalent of what CMake would generate from the reference's
 * src/StrumpackConfig.h.in:33-81 for a sequential OpenMP CPU build:
 * no MPI, no CUDA, no METIS, flop counters on. */
#ifndef STRUMPACK_CONFIG_H
#define STRUMPACK_CONFIG_H
#include <stdbool.h>
#define STRUMPACK_USE_OPENMP
#define strumpack_blas_int int
#define STRUMPACK_USE_GETOPT
#define STRUMPACK_COUNT_FLOPS
#define STRUMPACK_USE_OPENMP_TASKLOOP
#define STRUMPACK_USE_OPENMP_TASK_DEPEND
#define STRUMPACK_PBLAS_BLOCKSIZE 32
#define STRUMPACK_VERSION_MAJOR 8
#define STRUMPACK_VERSION_MINOR 0
#define STRUMPACK_VERSION_PATCH 0
inline void get_version(int* major, int* minor, int* patch) {
  *major = STRUMPACK_VERSION_MAJOR;
  *minor = STRUMPACK_VERSION_MINOR;
  *patch = STRUMPACK_VERSION_PATCH;
}
inline bool have_parmetis() { return false; }
inline bool have_scotch() { return false; }
inline bool have_pt_scotch() { return false; }
inline bool have_papi() { return false; }
inline bool have_combblas() { return false; }
inline bool have_butterflypack() { return false; }
inline bool have_zfp() { return false; }
inline bool have_slate() { return false; }
inline bool have_getopt() { return true; }
inline bool have_magma() { return false; }
inline bool have_kblas() { return false; }
inline bool have_matlab() { return false; }
#endif
