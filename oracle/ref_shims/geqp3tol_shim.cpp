/*
 * ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product.
 *
 * The reference ships 8 Fortran-77 routines that cannot be compiled in this
 * image (no gfortran):
 *   {s,d,c,z}geqp3tol_   reference src/dense/lapack/{s,d,c,z}geqp3tol.f
 *   my{s,d,c,z}lapmr_    reference src/dense/lapack/{s,d,c,z}lapmr.f
 * This file restates them in C++ on top of the LAPACK building blocks the
 * Fortran itself calls (xLAQPS / xLAQP2 / xGEQRF / xORMQR|xUNMQR / xLAPMR),
 * so that the reference's C++ sources link unchanged.
 *
 * xGEQP3TOL follows dgeqp3tol.f line by line:
 *   - caller supplies the column norms in WORK(1:N) (RWORK for complex),
 *     dgeqp3tol.f:178-181 (the DNRM2 call is commented out there)
 *   - blocked sweep with xLAQPS, NB/NX from ILAENV (32 / 128 for xGEQRF in
 *     reference LAPACK), then xLAQP2 on the tail        (:183-232)
 *   - after every block the new diagonal entries are tested and the routine
 *     returns as soon as |A(c,c)|/|A(1,1)| <= RTOL or |A(c,c)| <= ATOL,
 *     RANK = number of accepted columns               (:203-209, :225-231)
 * MYxLAPMR is textually LAPACK's xLAPMR -> forwarded to it.
 */
#include <algorithm>
#include <cmath>
#include <complex>
#include <vector>

extern "C" {
  int ilaenv_(int* ispec, const char* name, const char* opts, int* n1, int* n2,
              int* n3, int* n4, int lname, int lopts);
#define DECL_REAL(p, T)                                                        \
  void p##laqps_(int*, int*, int*, int*, int*, T*, int*, int*, T*, T*, T*, T*, \
                 T*, int*);                                                    \
  void p##laqp2_(int*, int*, int*, T*, int*, int*, T*, T*, T*, T*);            \
  void p##geqrf_(int*, int*, T*, int*, T*, T*, int*, int*);                    \
  void p##ormqr_(const char*, const char*, int*, int*, int*, T*, int*, T*, T*, \
                 int*, T*, int*, int*, int, int);                              \
  void p##lapmr_(int*, int*, int*, T*, int*, int*);                            \
  void p##swap_(int*, T*, int*, T*, int*);
  DECL_REAL(s, float)
  DECL_REAL(d, double)
#define DECL_CPLX(p, T, R)                                                     \
  void p##laqps_(int*, int*, int*, int*, int*, T*, int*, int*, T*, R*, R*, T*, \
                 T*, int*);                                                    \
  void p##laqp2_(int*, int*, int*, T*, int*, int*, T*, R*, R*, T*);            \
  void p##geqrf_(int*, int*, T*, int*, T*, T*, int*, int*);                    \
  void p##unmqr_(const char*, const char*, int*, int*, int*, T*, int*, T*, T*, \
                 int*, T*, int*, int*, int, int);                              \
  void p##lapmr_(int*, int*, int*, T*, int*, int*);                            \
  void p##swap_(int*, T*, int*, T*, int*);
  DECL_CPLX(c, std::complex<float>, float)
  DECL_CPLX(z, std::complex<double>, double)
}

namespace {

  template<typename T> struct Lapack;
#define LAPACK_REAL(p, T)                                                      \
  template<> struct Lapack<T> {                                                \
    using real = T;                                                            \
    static constexpr const char* qrf = #p "GEQRF";                            \
    static void laqps(int* m, int* n, int* off, int* nb, int* kb, T* a,       \
                      int* lda, int* jpvt, T* tau, real* vn1, real* vn2,      \
                      T* auxv, T* f, int* ldf) {                               \
      p##laqps_(m, n, off, nb, kb, a, lda, jpvt, tau, vn1, vn2, auxv, f, ldf);\
    }                                                                          \
    static void laqp2(int* m, int* n, int* off, T* a, int* lda, int* jpvt,    \
                      T* tau, real* vn1, real* vn2, T* work) {                 \
      p##laqp2_(m, n, off, a, lda, jpvt, tau, vn1, vn2, work);                 \
    }                                                                          \
    static void geqrf(int* m, int* n, T* a, int* lda, T* tau, T* w, int* lw,  \
                      int* info) { p##geqrf_(m, n, a, lda, tau, w, lw, info); }\
    static void mqr(int* m, int* n, int* k, T* a, int* lda, T* tau, T* c,     \
                    int* ldc, T* w, int* lw, int* info) {                      \
      p##ormqr_("L", "T", m, n, k, a, lda, tau, c, ldc, w, lw, info, 1, 1);    \
    }                                                                          \
    static void swap(int* n, T* x, int* ix, T* y, int* iy) {                   \
      p##swap_(n, x, ix, y, iy);                                               \
    }                                                                          \
  };
  LAPACK_REAL(s, float)
  LAPACK_REAL(d, double)
#define LAPACK_CPLX(p, T, R)                                                   \
  template<> struct Lapack<T> {                                                \
    using real = R;                                                            \
    static constexpr const char* qrf = #p "GEQRF";                            \
    static void laqps(int* m, int* n, int* off, int* nb, int* kb, T* a,       \
                      int* lda, int* jpvt, T* tau, real* vn1, real* vn2,      \
                      T* auxv, T* f, int* ldf) {                               \
      p##laqps_(m, n, off, nb, kb, a, lda, jpvt, tau, vn1, vn2, auxv, f, ldf);\
    }                                                                          \
    static void laqp2(int* m, int* n, int* off, T* a, int* lda, int* jpvt,    \
                      T* tau, real* vn1, real* vn2, T* work) {                 \
      p##laqp2_(m, n, off, a, lda, jpvt, tau, vn1, vn2, work);                 \
    }                                                                          \
    static void geqrf(int* m, int* n, T* a, int* lda, T* tau, T* w, int* lw,  \
                      int* info) { p##geqrf_(m, n, a, lda, tau, w, lw, info); }\
    static void mqr(int* m, int* n, int* k, T* a, int* lda, T* tau, T* c,     \
                    int* ldc, T* w, int* lw, int* info) {                      \
      p##unmqr_("L", "C", m, n, k, a, lda, tau, c, ldc, w, lw, info, 1, 1);    \
    }                                                                          \
    static void swap(int* n, T* x, int* ix, T* y, int* iy) {                   \
      p##swap_(n, x, ix, y, iy);                                               \
    }                                                                          \
  };
  LAPACK_CPLX(c, std::complex<float>, float)
  LAPACK_CPLX(z, std::complex<double>, double)

  /* One body for the four precisions.  `norms` is WORK for real types and
   * RWORK for complex ones (length 2N: vn1 | vn2); `work` is the scratch the
   * Fortran passes to xLAQPS/xLAQP2 (for real types that is WORK(2N+1:),
   * for complex types WORK(1:)). */
  template<typename T> void geqp3tol_body
  (int M, int N, T* A, int LDA, int* JPVT, T* TAU, T* work, int lwork_scratch,
   typename Lapack<T>::real* norms, int* INFO, int* RANK,
   typename Lapack<T>::real RTOL, typename Lapack<T>::real ATOL) {
    using L = Lapack<T>;
    *INFO = 0;
    *RANK = 0;
    const int MINMN = std::min(M, N);
    if (MINMN == 0) return;
    auto a = [&](int i, int j) -> T& {  // 1-based
      return A[(i-1) + std::size_t(j-1)*LDA];
    };
    // move initial (fixed) columns up front            dgeqp3tol.f:141-157
    int NFXD = 1, one = 1;
    for (int J=1; J<=N; J++) {
      if (JPVT[J-1] != 0) {
        if (J != NFXD) {
          L::swap(&M, &a(1,J), &one, &a(1,NFXD), &one);
          JPVT[J-1] = JPVT[NFXD-1];
          JPVT[NFXD-1] = J;
        } else JPVT[J-1] = J;
        NFXD++;
      } else JPVT[J-1] = J;
    }
    NFXD--;
    // factorize fixed columns                           dgeqp3tol.f:158-170
    if (NFXD > 0) {
      int NA = std::min(M, NFXD), info;
      L::geqrf(&M, &NA, A, &LDA, TAU, work, &lwork_scratch, &info);
      if (NA < N) {
        int nc = N - NA;
        L::mqr(&M, &nc, &NA, A, &LDA, TAU, &a(1,NA+1), &LDA, work,
               &lwork_scratch, &info);
      }
    }
    if (NFXD >= MINMN) return;
    int SM = M - NFXD, SN = N - NFXD, SMINMN = MINMN - NFXD;
    int ispec1 = 1, ispec3 = 3, m1 = -1;
    int NB = ilaenv_(&ispec1, L::qrf, " ", &SM, &SN, &m1, &m1, 6, 1);
    int NBMIN = 2, NX = 0;
    if (NB > 1 && NB < SMINMN)
      NX = std::max(0, ilaenv_(&ispec3, L::qrf, " ", &SM, &SN, &m1, &m1, 6, 1));
    // norms supplied by the caller; copy vn1 -> vn2     dgeqp3tol.f:178-181
    auto* vn1 = norms;
    auto* vn2 = norms + N;
    for (int J=NFXD+1; J<=N; J++) vn2[J-1] = vn1[J-1];
    auto small = [&](int C) {
      auto d = std::abs(a(C,C));
      return d / std::abs(a(1,1)) <= RTOL || d <= ATOL;
    };
    int J = NFXD + 1;
    if (NB >= NBMIN && NB < SMINMN && NX < SMINMN) {
      const int TOPBMN = MINMN - NX;
      while (J <= TOPBMN) {
        int JB = std::min(NB, TOPBMN-J+1), FJB = 0;
        int nc = N-J+1, off = J-1, ldf = N-J+1;
        L::laqps(&M, &nc, &off, &JB, &FJB, &a(1,J), &LDA, &JPVT[J-1],
                 &TAU[J-1], &vn1[J-1], &vn2[J-1], work, work+JB, &ldf);
        for (int C=J; C<=J+FJB-1; C++) {
          if (small(C)) return;
          (*RANK)++;
        }
        J += FJB;
      }
    }
    if (J <= MINMN) {
      int nc = N-J+1, off = J-1;
      L::laqp2(&M, &nc, &off, &a(1,J), &LDA, &JPVT[J-1], &TAU[J-1],
               &vn1[J-1], &vn2[J-1], work);
      for (int C=J; C<=MINMN; C++) {
        if (small(C)) return;
        (*RANK)++;
      }
    }
  }

  template<typename T> int lwkopt(int M, int N) {
    if (std::min(M, N) == 0) return 1;
    int ispec1 = 1, m1 = -1;
    int NB = ilaenv_(&ispec1, Lapack<T>::qrf, " ", &M, &N, &m1, &m1, 6, 1);
    return 2*N + (N+1)*NB;
  }
}

extern "C" {

  void sgeqp3tol_(int* M, int* N, float* A, int* LDA, int* JPVT, float* TAU,
                  float* WORK, int* LWORK, int* INFO, int* RANK,
                  float* RTOL, float* ATOL) {
    *RANK = 0; *INFO = 0;
    if (*LWORK == -1) { WORK[0] = float(lwkopt<float>(*M, *N)); return; }
    geqp3tol_body<float>(*M, *N, A, *LDA, JPVT, TAU, WORK+2*(*N),
                         *LWORK-2*(*N), WORK, INFO, RANK, *RTOL, *ATOL);
  }
  void dgeqp3tol_(int* M, int* N, double* A, int* LDA, int* JPVT, double* TAU,
                  double* WORK, int* LWORK, int* INFO, int* RANK,
                  double* RTOL, double* ATOL) {
    *RANK = 0; *INFO = 0;
    if (*LWORK == -1) { WORK[0] = double(lwkopt<double>(*M, *N)); return; }
    geqp3tol_body<double>(*M, *N, A, *LDA, JPVT, TAU, WORK+2*(*N),
                          *LWORK-2*(*N), WORK, INFO, RANK, *RTOL, *ATOL);
  }
  void cgeqp3tol_(int* M, int* N, std::complex<float>* A, int* LDA, int* JPVT,
                  std::complex<float>* TAU, std::complex<float>* WORK,
                  int* LWORK, float* RWORK, int* INFO, int* RANK,
                  float* RTOL, float* ATOL) {
    *RANK = 0; *INFO = 0;
    if (*LWORK == -1) { WORK[0] = float(lwkopt<std::complex<float>>(*M, *N)); return; }
    geqp3tol_body<std::complex<float>>(*M, *N, A, *LDA, JPVT, TAU, WORK,
                                       *LWORK, RWORK, INFO, RANK, *RTOL, *ATOL);
  }
  void zgeqp3tol_(int* M, int* N, std::complex<double>* A, int* LDA, int* JPVT,
                  std::complex<double>* TAU, std::complex<double>* WORK,
                  int* LWORK, double* RWORK, int* INFO, int* RANK,
                  double* RTOL, double* ATOL) {
    *RANK = 0; *INFO = 0;
    if (*LWORK == -1) { WORK[0] = double(lwkopt<std::complex<double>>(*M, *N)); return; }
    geqp3tol_body<std::complex<double>>(*M, *N, A, *LDA, JPVT, TAU, WORK,
                                        *LWORK, RWORK, INFO, RANK, *RTOL, *ATOL);
  }

  /* MYxLAPMR == LAPACK xLAPMR (reference src/dense/lapack/dlapmr.f). */
  void myslapmr_(int* fwd, int* m, int* n, float* x, int* ldx, int* k) {
    slapmr_(fwd, m, n, x, ldx, k);
  }
  void mydlapmr_(int* fwd, int* m, int* n, double* x, int* ldx, int* k) {
    dlapmr_(fwd, m, n, x, ldx, k);
  }
  void myclapmr_(int* fwd, int* m, int* n, std::complex<float>* x, int* ldx, int* k) {
    clapmr_(fwd, m, n, x, ldx, k);
  }
  void myzlapmr_(int* fwd, int* m, int* n, std::complex<double>* x, int* ldx, int* k) {
    zlapmr_(fwd, m, n, x, ldx, k);
  }
}
