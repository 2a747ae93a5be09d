// strumpack_b200 -- ulv_qr3: left-looking blocked Householder QR of one ULV
// factor block per CTA, warp-specialised, fed by the TMA bulk-copy engine.
//
// Replaces (reference, CPU): W0.LQ(...) + the three GEMMs with Q of
// HSSMatrix<T>::factor_recursive (src/HSS/HSSMatrix.factor.hpp:122-141,
// DenseMatrix::LQ src/dense/DenseMatrix.cpp:693-719) for every node whose
// reduced block has m <= 256 rows (all leaves of a leaf-256 tree, all inner
// nodes).  Output format = ulv_qr_kernel<16,...>: R + Householder vectors in the
// factor block, one 16 x 16 T per 16-column panel, Q^T applied to [Vh | W1^T].
//
// Why another kernel: the right-looking kernel streams the trailing matrix from
// L2 once per panel (2.7-4 flop per L2 byte, 5x the algorithmic DRAM traffic,
// DMMA pipe 34 % busy, round-1 ncu).  Here every column of the block is read
// ONCE and written ONCE; what is re-read are the finished reflector blocks V_j,
// and they arrive through an mbarrier ring filled by cp.async.bulk (UBLKCP),
// never through a load instruction that a warp waits on:
//
//   warps 0-3  CHAIN group   factor panel p (16 columns, register-resident
//                            8-column sub-panels: the latency-bound Householder
//                            column chain), store R / V_p / T_p
//   warps 4-7  UPDATE group  hold panel p+1 in REGISTERS (as C^T accumulator
//                            tiles), apply V_0..V_{p-1} from the ring while the
//                            chain of panel p runs (lookahead), then V_p from the
//                            chain group's shared-memory panel, hand the panel over
//   (pump)     PRODUCER      no warp of its own (a ninth warp would cap the CTA
//                            at 96 registers per thread): lane 0 of the first
//                            update warp issues the bulk copies of (T_j, V_j) in
//                            64-row chunks into a 6-stage ring whenever the
//                            consumers' "empty" mbarriers and the "block j
//                            stored" flag allow, at every point where that warp
//                            would otherwise spin on a "full" barrier
//
// All fp64 products run on the tensor pipe (mma.sync m8n8k4).  The update keeps
// C^T in accumulator layout: W^T = C^T V, W2^T = W^T T and C^T -= W2^T V^T chain
// through the accumulators with K-slot permutations, so C never goes through
// shared memory and every V fragment is one conflict-free LDS.128 feeding four
// DMMAs.  2 CTAs (two chains + two update groups) per SM.
#pragma once

#include <cstdint>

#include "hss_engine.hpp"
#include "sb200_common.cuh"

namespace sb200 {
namespace qr3 {

constexpr int NB = 16;            // panel width
constexpr int CH = 64;            // rows per ring chunk (4 row blocks of 16)
#ifdef QR3_NST
constexpr int NST = QR3_NST;      // experiment: deeper ring (one CTA per SM)
#else
constexpr int NST = 6;            // ring stages
#endif
constexpr int LDX = 258;          // ld of the panel buffer (== 2 mod 16), m <= 256
constexpr int LDT = 20;           // ld of the chain group's T block in shared memory
constexpr int LDTR = 16;          // ld of a T block in the T ring (= its layout in global memory: one bulk copy)
constexpr int NTS = 2;            // T ring stages
// one ring stage = 4 row blocks x [16 columns][16 rows], each 128-byte line (one
// column of one row block) XOR-swizzled in 16-byte units by (column & 7): the
// layout cp.async.bulk.tensor writes with CU_TENSOR_MAP_SWIZZLE_128B, and the one
// that makes every fragment load below a conflict-free LDS.128
constexpr int STG = 4 * NB * 16;  // doubles per ring stage (8 KB)
constexpr int TSTG = LDT * NB;    // doubles of the chain group's T
constexpr int TRSTG = LDTR * NB;  // doubles per T ring stage
constexpr int NTHREADS = 256;     // 4 chain + 4 update warps
constexpr int MAXM = 256;

// shared memory map (doubles), from a 1024-byte aligned base (swizzle atoms)
constexpr int OFF_RING = 0;                      // NST stages
constexpr int OFF_X = OFF_RING + NST * STG;      // LDX x 16 panel buffer
constexpr int OFF_TRING = OFF_X + LDX * NB;      // NTS stages
constexpr int OFF_TS = OFF_TRING + NTS * TRSTG;  // T of the panel being factored
constexpr int OFF_WX = OFF_TS + TSTG;            // update group exchange: 2 parities x 4 warps x 256
constexpr int OFF_CGS = OFF_WX + 2048;           // chain group scratch (672 doubles)
constexpr int CGS_PAIR = 0, CGS_DIAG = 64, CGS_YS = 80, CGS_XCH = 144;   // XCH: 512 (also Ss)
constexpr int OFF_BAR = OFF_CGS + 672;           // mbarriers (uint64) + control block
constexpr int NBAR = 2 * NST + 2 * NTS;
#ifdef QR3_TIMING
constexpr int OFF_STAT = OFF_BAR + NBAR + 10;       // issue clocks [NST] + counters [8] (long long)
constexpr int SMEM_DOUBLES = OFF_STAT + 32 + 8;
#else
constexpr int SMEM_DOUBLES = OFF_BAR + NBAR + 10;   // + Ctl
#endif
constexpr size_t SMEM_BYTES = sizeof(double) * (size_t)SMEM_DOUBLES + 1024;   // + alignment slack
#ifndef QR3_NST
static_assert(SMEM_BYTES <= 115200, "two CTAs per SM");
#endif
static_assert(LDX * NB >= 4096, "chain group exchange aliases the panel buffer in the aug phase");
static_assert((OFF_X % 2) == 0 && (OFF_TRING % 2) == 0 && (OFF_TS % 2) == 0 && (OFF_WX % 2) == 0 && (OFF_CGS % 2) == 0 &&
              (OFF_BAR % 2) == 0, "16-byte aligned regions");

// ---- PTX wrappers ---------------------------------------------------------------
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(s32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b, int count) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;\n" ::"r"(s32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok)
      : "r"(s32(b)), "r"(parity)
      : "memory");
  return ok != 0;
}
// non-blocking probe (try_wait may put the thread to sleep for a system-defined
// time before it answers "not yet": fine for a consumer, fatal for the pump's
// "is this stage free?" polls)
__device__ __forceinline__ bool mbar_test(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok)
      : "r"(s32(b)), "r"(parity)
      : "memory");
  return ok != 0;
}
#ifdef QR3_WATCHDOG
#define QR3_SPIN_GUARD(what)                                                                         \
  if (++spins_ > (1ll << 22)) {                                                                      \
    printf("[qr3 watchdog] block %d thread %d stuck in %s (line %d)\n", blockIdx.x, threadIdx.x, what, __LINE__); \
    __trap();                                                                                        \
  }
#define QR3_SPIN_DECL long long spins_ = 0;
#else
#define QR3_SPIN_GUARD(what)
#define QR3_SPIN_DECL
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  QR3_SPIN_DECL
  while (!mbar_try(b, parity)) { QR3_SPIN_GUARD("mbar_wait") }
}
// TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(double* dst, const double* src, uint32_t bytes, uint64_t* b) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(s32(dst)),
      "l"(src), "r"(bytes), "r"(s32(b))
      : "memory");
}
// TMA tensor copy (2-D tile of the factor block) global -> shared (SASS: UTMALDG)
__device__ __forceinline__ void tma_tile_2d(double* dst, const void* tmap, int c0, int c1, uint64_t* b) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::
          "r"(s32(dst)),
      "l"(tmap), "r"(c0), "r"(c1), "r"(s32(b))
      : "memory");
}
// the arrival fires when all cp.async of this thread issued so far have landed
__device__ __forceinline__ void cpasync_arrive(uint64_t* b) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(s32(b)) : "memory");
}
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async;\n" ::: "memory"); }
__device__ __forceinline__ void nbar_sync(int id, int n) { asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void nbar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;\n" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.cta.shared.s32 %0, [%1];\n" : "=r"(v) : "r"(s32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.cta.shared.s32 [%0], %1;\n" ::"r"(s32(p)), "r"(v) : "memory");
}

#ifdef QR3_TIMING
#define QT_DECL long long qt_[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; long long qt_last_ = clock64();
#define QT(i) { const long long now_ = clock64(); qt_[i] += now_ - qt_last_; qt_last_ = now_; }
// after a bar.sync: the barrier blocks at the first dependent shared-memory access, not at issue
#define QTB(i) { unsigned d_; asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(d_) : "r"(s32(sm + OFF_BAR)) : "memory"); \
                 long long now_ = clock64(); now_ += (d_ == 0xffffffffu); qt_[i] += now_ - qt_last_; qt_last_ = now_; }
#define QT_PRINT(role)                                                                                     \
  if (lane == 0 && blockIdx.x == 0)                                                                        \
    printf("[qr3 timing] %s warp %d: %lld %lld %lld %lld %lld %lld %lld %lld | wait %lld ph1 %lld xT %lld ph2 %lld\n", role, wq, qt_[0], qt_[1], qt_[2], \
           qt_[3], qt_[4], qt_[5], qt_[6], qt_[7], qt_[8], qt_[9], qt_[10], qt_[11]);
#else
#define QT_DECL
#define QT(i)
#define QTB(i)
#define QT_PRINT(role)
#endif

// named barriers
constexpr int BAR_CG = 1, BAR_UG = 2, BAR_H1 = 3, BAR_H2 = 4, BAR_H3 = 5;

struct Geo {
  int m, k, naug, P, Q, nq;
  bool vec;   // 16-byte aligned columns: bulk copies + 128-bit global accesses
};

__device__ __forceinline__ void panel_cols(const Geo& G, int q, int& c0, int& jb) {
  if (q < G.P) { c0 = q * NB; jb = min(NB, G.k - c0); }
  else { c0 = G.k + (q - G.P) * NB; jb = min(NB, G.naug - c0); }
}

// position in a ring: stage + parity of its current use
struct RingPos {
  int s, ph;
  __device__ __forceinline__ void adv(int n, int nst) {
    s += n;
    while (s >= nst) { s -= nst; ph ^= 1; }
  }
};

// V[i][a] of block j as the update must see it, from whatever the ring / the
// panel buffer holds at that position: explicit unit lower trapezoid, zero
// beyond row m and beyond the block's jb columns
__device__ __forceinline__ double vfix(double v, int row, int a, int j, int m, int jb) {
  const int rl = row - NB * j;
  return (row < m && a < jb) ? (rl > a ? v : (rl == a ? 1. : 0.)) : 0.;
}

// ---------------------------------------------------------------------------------
// Register-resident panel of an update group (4 warps, wq = warp in group):
//   ct[s][b][u][e] = C[row 16*(4b+wq) + 4t + 2e + u][column 8s + g]
// i.e. for every own 16-row block two DMMA accumulator tiles (u = row parity) of
// C^T: D[g][2t+e] = C^T[col g][row 2(2t+e)+u].
// ---------------------------------------------------------------------------------
struct Panel {
  double ct[2][4][2][2];
};

__device__ __forceinline__ void panel_load(Panel& Pn, const Geo& G, const double* A, int q, int wq, int lane) {
  int c0, jb;
  panel_cols(G, q, c0, jb);
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int s = 0; s < 2; s++) {
    const int col = 8 * s + g;
    const double* src = A + (size_t)(c0 + col) * G.m;
    const bool cin = col < jb;
#pragma unroll
    for (int b = 0; b < 4; b++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int row = 16 * (4 * b + wq) + 4 * t + 2 * e;
        double x = 0., y = 0.;
        if (cin && row < G.m) {
          if (G.vec) { const double2 v = *reinterpret_cast<const double2*>(src + row); x = v.x; y = v.y; }
          else { x = src[row]; if (row + 1 < G.m) y = src[row + 1]; }
        }
        Pn.ct[s][b][0][e] = x;
        Pn.ct[s][b][1][e] = y;
      }
  }
}

__device__ __forceinline__ void panel_store_global(const Panel& Pn, const Geo& G, double* A, int q, int wq, int lane) {
  int c0, jb;
  panel_cols(G, q, c0, jb);
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int s = 0; s < 2; s++) {
    const int col = 8 * s + g;
    double* dst = A + (size_t)(c0 + col) * G.m;
    if (col >= jb) continue;
#pragma unroll
    for (int b = 0; b < 4; b++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int row = 16 * (4 * b + wq) + 4 * t + 2 * e;
        if (row < G.m) {
          if (G.vec) *reinterpret_cast<double2*>(dst + row) = make_double2(Pn.ct[s][b][0][e], Pn.ct[s][b][1][e]);
          else { dst[row] = Pn.ct[s][b][0][e]; if (row + 1 < G.m) dst[row + 1] = Pn.ct[s][b][1][e]; }
        }
      }
  }
}

// all 256 rows and 16 columns (zeros beyond m / jb) into the panel buffer
__device__ __forceinline__ void panel_store_x(const Panel& Pn, double* X, int wq, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int s = 0; s < 2; s++)
#pragma unroll
    for (int b = 0; b < 4; b++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int row = 16 * (4 * b + wq) + 4 * t + 2 * e;
        *reinterpret_cast<double2*>(X + (8 * s + g) * LDX + row) = make_double2(Pn.ct[s][b][0][e], Pn.ct[s][b][1][e]);
      }
}

// Fragment addressing.  Ring stage (swz): row block rb at base = stage + rb*256,
// element (row r, column a) at a*16 + ((r/2 ^ (a&7)) * 2) + (r&1).  Panel buffer
// X (!swz): base = X + 16 ib, element at a*LDX + r.
// phase 1 on one own row block: wt += C^T V.  The eight DMMAs of one k-step pair
// cycle over the four accumulators (the asm statements keep their order:
// back-to-back DMMAs on one accumulator would serialise on the 32-clock
// tensor-pipe latency).
__device__ __forceinline__ void blk_phase1(const Panel& Pn, int b, const double* base, bool swz, double (&wt)[2][2][2],
                                           bool fix, int ib, int j, int m, int jb, int g, int t) {
#pragma unroll
  for (int e = 0; e < 2; e++) {
    // rows 4t+2e, 4t+2e+1 (16-byte unit 2t+e) of columns g and 8+g
    const int o0 = swz ? g * 16 + (((2 * t + e) ^ g) << 1) : g * LDX + 4 * t + 2 * e;
    const int o1 = swz ? o0 + 8 * 16 : o0 + 8 * LDX;
    double2 v0 = *reinterpret_cast<const double2*>(base + o0);
    double2 v1 = *reinterpret_cast<const double2*>(base + o1);
    if (fix) {
      const int row = 16 * ib + 4 * t + 2 * e;
      v0.x = vfix(v0.x, row, g, j, m, jb);
      v0.y = vfix(v0.y, row + 1, g, j, m, jb);
      v1.x = vfix(v1.x, row, 8 + g, j, m, jb);
      v1.y = vfix(v1.y, row + 1, 8 + g, j, m, jb);
    }
    dmma(wt[0][0][0], wt[0][0][1], Pn.ct[0][b][0][e], v0.x);
    dmma(wt[1][0][0], wt[1][0][1], Pn.ct[1][b][0][e], v0.x);
    dmma(wt[0][1][0], wt[0][1][1], Pn.ct[0][b][0][e], v1.x);
    dmma(wt[1][1][0], wt[1][1][1], Pn.ct[1][b][0][e], v1.x);
    dmma(wt[0][0][0], wt[0][0][1], Pn.ct[0][b][1][e], v0.y);
    dmma(wt[1][0][0], wt[1][0][1], Pn.ct[1][b][1][e], v0.y);
    dmma(wt[0][1][0], wt[0][1][1], Pn.ct[0][b][1][e], v1.y);
    dmma(wt[1][1][0], wt[1][1][1], Pn.ct[1][b][1][e], v1.y);
  }
}

// phase 2 on one own row block: C^T += w2n V^T   (w2n = -(W^T T))
__device__ __forceinline__ void blk_phase2(Panel& Pn, int b, const double* base, bool swz, const double (&w2n)[2][2][2],
                                           bool fix, int ib, int j, int m, int jb, int g, int t) {
#pragma unroll
  for (int atp = 0; atp < 2; atp++)
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int a = 8 * atp + 2 * t + e;   // rows 2g, 2g+1 (unit g) of column a
      const int o = swz ? a * 16 + ((g ^ (2 * t + e)) << 1) : a * LDX + 2 * g;
      double2 v = *reinterpret_cast<const double2*>(base + o);
      if (fix) {
        const int row = 16 * ib + 2 * g;
        v.x = vfix(v.x, row, a, j, m, jb);
        v.y = vfix(v.y, row + 1, a, j, m, jb);
      }
#pragma unroll
      for (int s = 0; s < 2; s++) {
        dmma(Pn.ct[s][b][0][0], Pn.ct[s][b][0][1], w2n[s][atp][e], v.x);
        dmma(Pn.ct[s][b][1][0], Pn.ct[s][b][1][1], w2n[s][atp][e], v.y);
      }
    }
}

// combine the 4 warps' partial W^T through shared memory, then W2^T = W^T T
// (T upper triangular), negated.  xch: this group's 2 x 4 x 256 exchange buffer.
__device__ __forceinline__ void exchange_and_T(double (&wt)[2][2][2], double (&w2n)[2][2][2], double* xch, int& par,
                                               int barid, const double* Tp, int ldt, int jb, int wq, int lane) {
  const int g = lane >> 2, t = lane & 3;
  double* mine = xch + par * 1024 + wq * 256;
#pragma unroll
  for (int s = 0; s < 2; s++)
#pragma unroll
    for (int at = 0; at < 2; at++)
#pragma unroll
      for (int e = 0; e < 2; e++) mine[((s * 2 + at) * 2 + e) * 32 + lane] = wt[s][at][e];
  nbar_sync(barid, 128);
  const double* all = xch + par * 1024;
#pragma unroll
  for (int s = 0; s < 2; s++)
#pragma unroll
    for (int at = 0; at < 2; at++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int idx = ((s * 2 + at) * 2 + e) * 32 + lane;
        wt[s][at][e] = (all[idx] + all[256 + idx]) + (all[512 + idx] + all[768 + idx]);
      }
  par ^= 1;
#pragma unroll
  for (int s = 0; s < 2; s++)
#pragma unroll
    for (int q = 0; q < 2; q++) w2n[s][q][0] = w2n[s][q][1] = 0.;
#pragma unroll
  for (int atp = 0; atp < 2; atp++)
#pragma unroll
    for (int at = 0; at < 2; at++)
      if (at <= atp) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int ta = 8 * at + 2 * t + e, tc = 8 * atp + g;
          const double bb = (ta < jb && tc < jb) ? Tp[ta + tc * ldt] : 0.;
#pragma unroll
          for (int s = 0; s < 2; s++) dmma(w2n[s][atp][0], w2n[s][atp][1], wt[s][at][e], bb);
        }
      }
#pragma unroll
  for (int s = 0; s < 2; s++)
#pragma unroll
    for (int q = 0; q < 2; q++) { w2n[s][q][0] = -w2n[s][q][0]; w2n[s][q][1] = -w2n[s][q][1]; }
}

// chunks of block j: c in [j >> 2, (m - 1) >> 6]
__device__ __forceinline__ int first_chunk(int j) { return (NB * j) / CH; }
__device__ __forceinline__ int last_chunk(int m) { return (m - 1) / CH; }

// control block in shared memory (behind the mbarriers)
struct Ctl {
  int nstored;   // panels whose R / V / T the chain group has written to global memory
  // pump: next ring / T stage to fill, position in the schedule
  int pos_s, pos_ph, tpos_s, tpos_ph;
  int p;         // chain phase: blocks 0..p-1 are streamed for panel p+1; aug phase: first panel of the round
  int j, c;      // next block, next chunk of it (c < 0: T_j not yet issued)
  int aug, done;
  int m, k, P, nq, vec;
};
__device__ __forceinline__ uint64_t* bar_full(double* sm) { return reinterpret_cast<uint64_t*>(sm + OFF_BAR); }
__device__ __forceinline__ uint64_t* bar_empty(double* sm) { return bar_full(sm) + NST; }
__device__ __forceinline__ uint64_t* bar_tfull(double* sm) { return bar_full(sm) + 2 * NST; }
__device__ __forceinline__ uint64_t* bar_tempty(double* sm) { return bar_full(sm) + 2 * NST + NTS; }
__device__ __forceinline__ Ctl* ctl_of(double* sm) { return reinterpret_cast<Ctl*>(sm + OFF_BAR + NBAR); }

// ---------------------------------------------------------------------------------
// PRODUCER ("pump"): called by the first warp of the update group (all 32 lanes)
// wherever it would otherwise spin, and after every release.  Issues as many of
// the next (T_j, V_j chunk) loads of the fixed schedule as the ring has free
// stages for and whose source block the chain group has stored; never blocks.
// Its state lives in shared memory (few live registers); inlined at four sites
// only: an ABI call here would spill the caller's whole register panel.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void pump_try_issue(double* sm, const double* A, const double* Tg, const void* tmap,
                                               int lane) {
  Ctl* C = ctl_of(sm);
  if (C->done) return;
#ifdef QR3_TIMING
  long long* stat = reinterpret_cast<long long*>(sm + OFF_STAT);
  const long long pump_t0 = clock64();
#endif
  int pos_s = C->pos_s, pos_ph = C->pos_ph, tpos_s = C->tpos_s, tpos_ph = C->tpos_ph;
  int p = C->p, j = C->j, c = C->c, aug = C->aug, done = 0;
  const int m = C->m, k = C->k, P = C->P, nq = C->nq, vec = C->vec;
  const int c_hi = last_chunk(m);
  bool moved = false;
  while (!done) {
    const int jb = min(NB, k - NB * j);
    int ok = 0;
    if (lane == 0) {
      ok = ld_acquire(&C->nstored) > j;
      if (ok) ok = c < 0 ? mbar_test(bar_tempty(sm) + tpos_s, tpos_ph ^ 1) : mbar_test(bar_empty(sm) + pos_s, pos_ph ^ 1);
    }
    ok = __shfl_sync(0xffffffffu, ok, 0);
    if (!ok) break;
    moved = true;
    if (c < 0) {   // T_j: 16 x 16 doubles, contiguous in global memory: one bulk copy
      uint64_t* fb = bar_tfull(sm) + tpos_s;
      double* dst = sm + OFF_TRING + tpos_s * TRSTG;
      const double* src = Tg + (size_t)(NB * j) * NB;
      if (vec) {
        if (lane == 0) {
          mbar_arrive_tx(fb, (uint32_t)(NB * NB * 8));
          bulk_g2s(dst, src, NB * NB * 8, fb);
        }
      } else {
        for (int idx = lane; idx < jb * NB; idx += 32) cp_async8(dst + idx, src + idx, true);
        cpasync_arrive(fb);
      }
      if (++tpos_s == NTS) { tpos_s = 0; tpos_ph ^= 1; }
      c = first_chunk(j);
      continue;
    }
    {
      uint64_t* fb = bar_full(sm) + pos_s;
      double* stage = sm + OFF_RING + pos_s * STG;
#ifdef QR3_TIMING
      if (lane == 0) { stat[pos_s] = clock64(); stat[32 + 2]++; }
#endif
      if (vec) {
        // lane rb issues the 16 x 16 tile of row block 4c + rb (rows above the block's
        // diagonal tile and beyond m are never read: not loaded)
        const int ib = 4 * c + lane;
        const bool need = lane < 4 && ib >= j && 16 * ib < m;
        const unsigned mask = __ballot_sync(0xffffffffu, need);
        if (lane == 0) mbar_arrive_tx(fb, (uint32_t)(__popc(mask) * NB * 16 * 8));
        __syncwarp();
        if (need) tma_tile_2d(stage + lane * 256, tmap, 16 * ib, NB * j, fb);
      } else {
        const int r_lo = max(NB * j, CH * c), r_hi = min(m, CH * c + CH);
        const int nr = r_hi - r_lo;
        const double* src = A + (size_t)(NB * j) * m;
        for (int idx = lane; idx < jb * nr; idx += 32) {
          const int a = idx / nr, r_abs = r_lo + (idx - a * nr);
          const int rb = (r_abs >> 4) & 3, r = r_abs & 15;
          cp_async8(stage + rb * 256 + a * 16 + ((((r >> 1) ^ (a & 7))) << 1) + (r & 1), src + (size_t)a * m + r_abs, true);
        }
        cpasync_arrive(fb);
      }
      if (++pos_s == NST) { pos_s = 0; pos_ph ^= 1; }
      if (c < c_hi) { c++; continue; }
    }
    // next block of the schedule
    c = -1;
    j++;
    if (!aug) {
      if (j < p) continue;
      j = 0; p++;
      if (p < P && p + 1 < nq) continue;
      aug = 1; p = P + 1;          // first aug round: panels P+1 (chain group), P+2
      if (p >= nq) done = 1;
    } else {
      if (j < P) continue;
      j = 0; p += 2;
      if (p >= nq) done = 1;
    }
  }
#ifdef QR3_TIMING
  if (lane == 0) { stat[32 + 0] += clock64() - pump_t0; stat[32 + 1]++; }
#endif
  if (moved) {
    __syncwarp();
    if (lane == 0) {
      C->pos_s = pos_s; C->pos_ph = pos_ph; C->tpos_s = tpos_s; C->tpos_ph = tpos_ph;
      C->p = p; C->j = j; C->c = c; C->aug = aug; C->done = done;
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------
// CONSUMER loop of a group of four warps.  group 1 (warps 4-7, "update group"):
// chain-phase steps p = 0..P-1 on panel p+1 (reflector blocks 0..p-1 from the
// ring, block p from the chain group's panel buffer), then its share of the
// remaining aug panels.  group 0 (the chain group, once its chain is done): its
// share of the remaining aug panels.  ONE copy of the apply code for all of it.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void consumer(const Geo& G, double* sm, double* A, const double* Tg, const void* tmap,
                                         int group, int wq, int lane, RingPos pos, RingPos tpos) {
  const int g = lane >> 2, t = lane & 3;
  const bool ug = group == 1, is_pump = ug && wq == 0;
  double* X = sm + OFF_X;
  double* xch = ug ? sm + OFF_WX : sm + OFF_X;   // the chain group only gets here when X is idle
  const int barid = ug ? BAR_UG : BAR_CG;
  const int c_hi = last_chunk(G.m);
  Panel Pn;
  int par = 0;
  // the pumping warp keeps the ring fed while it waits; leaving the loop is lane
  // 0's decision (the pump is a warp-collective), every lane then observes the
  // completed phase itself
  auto wait_full = [&](uint64_t* b, int ph) {
    if (is_pump) {
      QR3_SPIN_DECL
      for (;;) {
        int ok = lane == 0 ? (int)mbar_test(b, ph) : 0;
        ok = __shfl_sync(0xffffffffu, ok, 0);
        if (ok) break;
        pump_try_issue(sm, A, Tg, tmap, lane);
        QR3_SPIN_GUARD("wait_full(pump)")
      }
    }
    mbar_wait(b, ph);
  };
  QT_DECL
  for (int step = ug ? 0 : G.P;; step++) {
    const bool chain = step < G.P;
    int q, nring, rel;
    if (chain) { q = step + 1; nring = step; rel = 2; }
    else {
      const int qa = G.P + 1 + 2 * (step - G.P);
      if (qa >= G.nq) break;
      q = qa + (ug ? 1 : 0); nring = G.P; rel = 1;
    }
    const bool has = q < G.nq;
    if (is_pump) pump_try_issue(sm, A, Tg, tmap, lane);
    if (has) panel_load(Pn, G, A, q, wq, lane);
    QT(0)
    const int nblk = (chain && !has) ? 0 : nring;   // a chain step without a panel has no ring traffic at all
    for (int jj = 0; jj <= nblk; jj++) {
      const bool from_x = jj == nblk;
      if (from_x) {
        if (!chain) break;
        QT(1)
        nbar_sync(BAR_H1, 256);   // panel `step` factored: explicit V in X, T in Ts
        QTB(2)
        if (!has) break;
      }
      const int j = from_x ? step : jj;
      const int jb = min(NB, G.k - NB * j);
      const int c_lo = first_chunk(j);
      const RingPos p0 = pos;
      if (!from_x) {   // T_j, then every chunk of V_j (one wait loop: one inlined copy of the pump)
        RingPos pc = pos;
        for (int c = c_lo - 1; c <= c_hi; c++) {
          uint64_t* b = bar_tfull(sm) + tpos.s;
          int ph = tpos.ph;
#ifdef QR3_TIMING
          const int st_ = pc.s;
          const bool chunk_ = c >= c_lo;
          const bool ready_ = chunk_ ? mbar_test(bar_full(sm) + pc.s, pc.ph) : true;
#endif
          if (c >= c_lo) { b = bar_full(sm) + pc.s; ph = pc.ph; pc.adv(1, NST); }
          wait_full(b, ph);
#ifdef QR3_TIMING
          if (is_pump && lane == 0 && chunk_) {
            long long* stat = reinterpret_cast<long long*>(sm + OFF_STAT);
            stat[32 + 3]++;
            if (!ready_) { stat[32 + 4]++; stat[32 + 5] += clock64() - stat[st_]; }
          }
#endif
        }
      }
      QT(8)
      if (has) {
        double wt[2][2][2], w2n[2][2][2];
#pragma unroll
        for (int s = 0; s < 2; s++)
#pragma unroll
          for (int a = 0; a < 2; a++) wt[s][a][0] = wt[s][a][1] = 0.;
#pragma unroll
        for (int c = 0; c < 4; c++) {
          const int ib = 4 * c + wq;
          if (ib >= j && 16 * ib < G.m) {
            int sc = p0.s + (c - c_lo);
            if (sc >= NST) sc -= NST;
            const double* vp = from_x ? X + 16 * ib : sm + OFF_RING + sc * STG + 256 * wq;
            const bool fix = ib == j || 16 * ib + 16 > G.m || jb < NB;
            blk_phase1(Pn, c, vp, !from_x, wt, fix, ib, j, G.m, jb, g, t);
          }
        }
        QT(9)
        exchange_and_T(wt, w2n, xch, par, barid, from_x ? sm + OFF_TS : sm + OFF_TRING + tpos.s * TRSTG,
                       from_x ? LDT : LDTR, jb, wq, lane);
        QTB(10)
        if (!from_x) {
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_tempty(sm) + tpos.s, rel);
        }
#pragma unroll
        for (int c = 0; c < 4; c++) {
          const int ib = 4 * c + wq;
          int sc = p0.s + (c - c_lo);
          if (sc >= NST) sc -= NST;
          if (ib >= j && 16 * ib < G.m) {
            const double* vp = from_x ? X + 16 * ib : sm + OFF_RING + sc * STG + 256 * wq;
            const bool fix = ib == j || 16 * ib + 16 > G.m || jb < NB;
            blk_phase2(Pn, c, vp, !from_x, w2n, fix, ib, j, G.m, jb, g, t);
          }
          if (!from_x && c >= c_lo && c <= c_hi) {   // last use of this chunk
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty(sm) + sc, rel);
            // the stage is free once the other three warps have released it too: probing
            // after every second chunk is as good as after each
            if ((c & 1) && is_pump) pump_try_issue(sm, A, Tg, tmap, lane);
          }
        }
      } else {
        // no panel in this aug round: release what the other group consumes
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(bar_tempty(sm) + tpos.s, rel);
          RingPos pc = pos;
          for (int c = c_lo; c <= c_hi; c++) { mbar_arrive(bar_empty(sm) + pc.s, rel); pc.adv(1, NST); }
        }
        if (is_pump) pump_try_issue(sm, A, Tg, tmap, lane);
      }
      if (!from_x) { tpos.adv(1, NTS); pos.adv(c_hi - c_lo + 1, NST); }
      QT(11)
    }
    QT(3)
    if (chain) {
      nbar_sync(BAR_H2, 256);   // the chain group has stored panel `step`: X is free
      QTB(4)
      if (has) {
        if (q < G.P) panel_store_x(Pn, X, wq, lane);
        else panel_store_global(Pn, G, A, q, wq, lane);
      }
      nbar_arrive(BAR_H3, 256);   // panel step+1 is in X / this group no longer reads X
    } else if (has) {
      panel_store_global(Pn, G, A, q, wq, lane);
    }
    QT(5)
  }
  QT_PRINT(ug ? "UG load/ring/waitH1/applyX/waitH2/store" : "CG(aug) load/ring/-/-/-/store")
#ifdef QR3_TIMING
  if (is_pump && lane == 0 && blockIdx.x == 0) {
    long long* stat = reinterpret_cast<long long*>(sm + OFF_STAT);
    printf("[qr3 pump] clk in pump %lld calls %lld chunks issued %lld | chunk waits: %lld of %lld not ready, sum(issue->seen) %lld\n",
           stat[32], stat[33], stat[34], stat[36], stat[35], stat[37]);
  }
#endif
}

// ---------------------------------------------------------------------------------
// CHAIN group (warps 0-3): the panel lives in X (absolute rows, ld LDX)
// ---------------------------------------------------------------------------------
// cols 8..15 of the panel <- (I - V0 T0^T V0^T) cols 8..15, V0 = cols 0..7 (explicit)
__device__ __forceinline__ void cg_inpanel(double* X, const double* Ts, int p, int m, double* xch, int& par, int wq,
                                           int lane) {
  const int g = lane >> 2, t = lane & 3;
  double w0 = 0., w1 = 0.;
  for (int ib = wq; 16 * ib < m; ib += 4) {
    if (ib < p) continue;
    const double* cp = X + (8 + g) * LDX + 16 * ib + t;
    const double* vp = X + g * LDX + 16 * ib + t;
#pragma unroll
    for (int ks = 0; ks < 4; ks++) dmma(w0, w1, cp[4 * ks], vp[4 * ks]);   // W^T[n=g][a=2t+e]
  }
  double* mine = xch + par * 256 + wq * 64;
  mine[lane] = w0;
  mine[32 + lane] = w1;
  nbar_sync(BAR_CG, 128);
  const double* all = xch + par * 256;
  w0 = (all[lane] + all[64 + lane]) + (all[128 + lane] + all[192 + lane]);
  w1 = (all[32 + lane] + all[96 + lane]) + (all[160 + lane] + all[224 + lane]);
  par ^= 1;
  // W2^T[n][a'] = sum_a W^T[n][a] T[a][a']: k-step e has slot t <-> a = 2t+e
  double z0 = 0., z1 = 0.;
  dmma(z0, z1, w0, Ts[(2 * t) + g * LDT]);
  dmma(z0, z1, w1, Ts[(2 * t + 1) + g * LDT]);
  z0 = -z0; z1 = -z1;   // z_e = -W2^T[n=g][a'=2t+e]  == B[slot t <-> a'=2t+e][n=g]
  for (int ib = wq; 16 * ib < m; ib += 4) {
    if (ib < p) continue;
#pragma unroll
    for (int rt = 0; rt < 2; rt++) {
      const int i0 = 16 * ib + 8 * rt;
      double* c0p = X + (8 + 2 * t) * LDX + i0 + g;
      double c0 = c0p[0], c1 = c0p[LDX];
      dmma(c0, c1, X[(2 * t) * LDX + i0 + g], z0);
      dmma(c0, c1, X[(2 * t + 1) * LDX + i0 + g], z1);
      c0p[0] = c0;
      c0p[LDX] = c1;
    }
  }
}

// register-resident 8-column sub-panel chain (see ulv_qr_kernel<NB, true>): rows
// c0 + [0, 256) of columns cs..cs+7; 4 warps, lane owns rows r0 and r0 + 32.
__device__ __forceinline__ void cg_subpanel(double* Pn, double* Ts, int cs, int sbw, int mp, double* pair, double* diag,
                                            int wq, int lane) {
  double a[8][2];
  const int r0 = wq * 64 + lane;
#pragma unroll
  for (int q = 0; q < 8; q++)
#pragma unroll
    for (int rr = 0; rr < 2; rr++) {
      const int i = r0 + 32 * rr;
      a[q][rr] = i < mp ? Pn[i + (cs + q) * LDX] : 0.;
    }
  double tr[8];
#pragma unroll
  for (int q = 0; q < 8; q++) tr[q] = 0.;
#pragma unroll
  for (int cq = 0; cq < 8; cq++) {
    if (cq < sbw) {
      const int c = cs + cq;   // diagonal row c < 16: warp 0, rr = 0, lane c
      double p[8];
#pragma unroll
      for (int q = 0; q < 8; q++) p[q] = 0.;
#pragma unroll
      for (int rr = 0; rr < 2; rr++) {
        const int i = r0 + 32 * rr;
        const double xv = i > c ? a[cq][rr] : 0.;
#pragma unroll
        for (int q = 0; q < 8; q++) p[q] += xv * a[q][rr];
      }
      // transpose-reduce: lane L ends up with the warp total of value (L >> 2) & 7
      double rsum;
      {
        const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
        double r1[4], r2[2];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const double send = b4 ? p[q] : p[q + 4];
          const double keep = b4 ? p[q + 4] : p[q];
          r1[q] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
#pragma unroll
        for (int q = 0; q < 2; q++) {
          const double send = b3 ? r1[q] : r1[q + 2];
          const double keep = b3 ? r1[q + 2] : r1[q];
          r2[q] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
        {
          const double send = b2 ? r2[0] : r2[1];
          const double keep = b2 ? r2[1] : r2[0];
          rsum = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
        rsum += __shfl_xor_sync(0xffffffffu, rsum, 2);
        rsum += __shfl_xor_sync(0xffffffffu, rsum, 1);
      }
      double* pw = pair + (cq & 1) * 32 + wq * 8;
      if ((lane & 3) == 0) pw[lane >> 2] = rsum;
      if (wq == 0 && lane == c) {
#pragma unroll
        for (int q = 0; q < 8; q++) diag[(cq & 1) * 8 + q] = a[q][0];
      }
      nbar_sync(BAR_CG, 128);
      const double* pp_ = pair + (cq & 1) * 32;
      double sm_[8], dg[8];
      {
        const double2* p2 = reinterpret_cast<const double2*>(pp_);
        const double2* d2 = reinterpret_cast<const double2*>(diag + (cq & 1) * 8);
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const double2 a0 = p2[q], a1 = p2[4 + q], a2 = p2[8 + q], a3 = p2[12 + q], dd = d2[q];
          sm_[2 * q] = (a0.x + a1.x) + (a2.x + a3.x);
          sm_[2 * q + 1] = (a0.y + a1.y) + (a2.y + a3.y);
          dg[2 * q] = dd.x; dg[2 * q + 1] = dd.y;
        }
      }
      const double pn = sm_[cq], alpha = dg[cq];
      double tc = 0., scal = 0., beta = alpha;
      if (pn > 0.) {   // dlarfg
        beta = -copysign(sqrt(alpha * alpha + pn), alpha);
        const double d = alpha - beta;
        scal = 1. / d;
        tc = -d / beta;
      }
      const bool isdiag = (wq == 0 && lane == c);
#pragma unroll
      for (int q = 0; q < 8; q++) {
        if (q > cq) {   // apply H_c to the columns to the right
          const double w = tc * (dg[q] + scal * sm_[q]);
          const double ws = w * scal;
#pragma unroll
          for (int rr = 0; rr < 2; rr++)
            if (r0 + 32 * rr > c) a[q][rr] -= ws * a[cq][rr];
          if (isdiag) a[q][0] -= w;
        }
      }
#pragma unroll
      for (int rr = 0; rr < 2; rr++)
        if (r0 + 32 * rr > c) a[cq][rr] *= scal;     // v_c
      if (isdiag) a[cq][0] = beta;                    // R(c,c)
      {   // column cq of the 8x8 T (dlarft): lane a (< 8, warp 0) keeps row a
        double val = (lane == cq) ? tc : 0.;
        double acc = 0.;
#pragma unroll
        for (int b = 0; b < 8; b++)
          if (b < cq) acc += (b >= lane ? tr[b] : 0.) * (dg[b] + scal * sm_[b]);
        if (lane < cq) val = -tc * acc;
        tr[cq] = val;
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 8; q++)
#pragma unroll
    for (int rr = 0; rr < 2; rr++) {
      const int i = r0 + 32 * rr;
      if (i < mp) Pn[i + (cs + q) * LDX] = a[q][rr];
    }
  if (wq == 0 && lane < 8) {
#pragma unroll
    for (int q = 0; q < 8; q++) Ts[(cs + lane) + (cs + q) * LDT] = (q >= lane) ? tr[q] : 0.;
  }
}

__device__ __forceinline__ void chain_group(const Geo& G, double* sm, double* A, double* Tg, const void* tmap, int wq,
                                            int lane) {
  double* X = sm + OFF_X;
  double* Ts = sm + OFF_TS;
  double* cgs = sm + OFF_CGS;
  int* nstored = &ctl_of(sm)->nstored;
  const int tid = wq * 32 + lane;
  const int m = G.m;
  double* pair = cgs + CGS_PAIR;
  double* diag = cgs + CGS_DIAG;
  double* Ys = cgs + CGS_YS;
  double* xch = cgs + CGS_XCH;   // in-panel exchange (2 x 256) / Ss (4 x 64)
  int par = 0;
  QT_DECL
  for (int p = 0; p < G.P; p++) {
    const int c0 = p * NB, jb = min(NB, G.k - c0);
    QT(7)
    if (p == 0) {
      // first panel straight from global (no earlier reflectors), zero padded
      for (int c = wq; c < NB; c += 4) {
        const double* src = A + (size_t)c * m;
        double* dst = X + c * LDX;
        const bool cin = c < jb;
        for (int i = lane; i < LDX; i += 32) dst[i] = (cin && i < m) ? src[i] : 0.;
      }
    } else {
      nbar_sync(BAR_H3, 256);   // the update group has written panel p into X
    }
    QTB(0)
    for (int idx = tid; idx < TSTG; idx += 128) Ts[idx] = 0.;
    nbar_sync(BAR_CG, 128);
    double* Pn = X + c0;
    const int mp = m - c0;
    const int nsub = (jb + 7) >> 3;
    for (int sp = 0; sp < nsub; sp++) {
      const int cs = sp * 8, sbw = min(8, jb - cs);
      cg_subpanel(Pn, Ts, cs, sbw, mp, pair, diag, wq, lane);
      nbar_sync(BAR_CG, 128);
      QTB(1)
      // R entries of the diagonal block to global; V explicit (unit diagonal)
      for (int cw = wq; cw < sbw; cw += 4) {
        const int c = cs + cw;
        double* dst = A + c0 + (size_t)(c0 + c) * m;
        if (lane <= c) {
          dst[lane] = Pn[lane + c * LDX];
          Pn[lane + c * LDX] = (lane == c) ? 1. : 0.;
        }
      }
      nbar_sync(BAR_CG, 128);
      if (sp == 0 && jb > 8) {
        QT(2)
        cg_inpanel(X, Ts, p, m, xch, par, wq, lane);
        nbar_sync(BAR_CG, 128);
        QT(3)
      }
    }
    QT(2)
    // T01 = -T00 (V0^T V1) T11 (block dlarft)
    if (nsub > 1) {
      const int cw = jb - 8;
      double* Ss = xch;
      {
        const int g = lane >> 2, t = lane & 3;
        double s0 = 0., s1 = 0.;
        const double* va = Pn + t + g * LDX;
        const double* vb = Pn + t + (8 + g) * LDX;
        const int mp8 = (mp + 7) & ~7;
        for (int i = 8 + 4 * wq; i < mp8; i += 16) dmma(s0, s1, va[i], vb[i]);
        Ss[wq * 64 + g + (2 * t) * 8] = s0;
        Ss[wq * 64 + g + (2 * t + 1) * 8] = s1;
      }
      nbar_sync(BAR_CG, 128);
      if (tid < 64) {
        const int a = tid & 7, cp = tid >> 3;
        double acc = 0.;
        for (int b = a; b < 8; b++) {
          const int si = b + cp * 8;
          acc += Ts[a + b * LDT] * ((Ss[si] + Ss[64 + si]) + (Ss[128 + si] + Ss[192 + si]));
        }
        Ys[a + cp * 8] = acc;
      }
      nbar_sync(BAR_CG, 128);
      if (tid < 64) {
        const int a = tid & 7, cp = tid >> 3;
        if (cp < cw) {
          double acc = 0.;
          for (int d = 0; d <= cp; d++) acc += Ys[a + d * 8] * Ts[(8 + d) + (8 + cp) * LDT];
          Ts[a + (8 + cp) * LDT] = -acc;
        }
      }
      nbar_sync(BAR_CG, 128);
    }
    QT(4)
    nbar_arrive(BAR_H1, 256);   // update group may apply V_p from X / Ts
    // ---- store: T_p, rows above the panel (final R), V strictly lower
    for (int idx = tid; idx < NB * NB; idx += 128) {
      const int a = idx & 15, c = idx >> 4;
      if (c < jb) Tg[a + (size_t)(c0 + c) * NB] = Ts[a + c * LDT];
    }
    for (int c = wq; c < jb; c += 4) {
      double* dst = A + (size_t)(c0 + c) * m;
      const double* src = X + c * LDX;
      for (int i = lane; i < c0; i += 32) dst[i] = src[i];
      for (int i = c0 + c + 1 + lane; i < m; i += 32) dst[i] = src[i];
    }
    __threadfence();
    fence_async();
    nbar_sync(BAR_CG, 128);
    if (tid == 0) st_release(nstored, p + 1);   // the producer may stream block p
    nbar_arrive(BAR_H2, 256);   // X may be overwritten
    QT(5)
  }
  nbar_sync(BAR_H3, 256);   // the update group has applied V_{P-1}: X is idle from here on
  QT(6)
  QT_PRINT("CG waitH3/chain/conv/inpanel/merge/store/-/-")
  // remaining aug panels: join the ring where the chain phase left it (skip what
  // the update group consumed alone: same schedule as the pump's)
  if (G.P + 1 < G.nq) {
    RingPos pos{0, 0}, tpos{0, 0};
    for (int p = 1; p < G.P && p + 1 < G.nq; p++)
      for (int j = 0; j < p; j++) {
        tpos.adv(1, NTS);
        pos.adv(last_chunk(G.m) - first_chunk(j) + 1, NST);
      }
    consumer(G, sm, A, Tg, tmap, 0, wq, lane, pos, tpos);
  }
}

__global__ void __launch_bounds__(NTHREADS, 2)
ulv_qr3_kernel(const DNode* __restrict__ nodes, const int* __restrict__ list, double* fact, double* tfac,
               const unsigned char* __restrict__ tmaps) {
  extern __shared__ __align__(16) double sm_raw[];
  // 1024-byte aligned base: the swizzle pattern of the ring stages is a function of the address
  double* sm = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(sm_raw) + 1023) & ~uintptr_t(1023));
  const int node = list[blockIdx.x];
  const DNode nd = nodes[node];
  if (nd.parent < 0 || nd.k == 0) return;
  Geo G;
  G.m = nd.m; G.k = nd.k; G.naug = nd.naug;
  G.P = (G.k + NB - 1) / NB;
  G.Q = (G.naug - G.k + NB - 1) / NB;
  G.nq = G.P + G.Q;
  // TMA feed needs 16-byte aligned columns (even m) and the node's tensor map
  G.vec = tmaps != nullptr && !(G.m & 1) && !(nd.F & 1);
  const void* tmap = tmaps ? tmaps + (size_t)node * 128 : nullptr;
#ifdef QR3_EXP_NOVEC
  G.vec = false;
#endif
  double* A = fact + nd.F;
  double* Tg = tfac + nd.T;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    const int fc = G.vec ? 1 : 32;   // bulk: one expect_tx arrival; else one cp.async arrival per lane of the pumping warp
    for (int s = 0; s < NST; s++) { mbar_init(bar_full(sm) + s, fc); mbar_init(bar_empty(sm) + s, 8); }
    for (int s = 0; s < NTS; s++) { mbar_init(bar_tfull(sm) + s, fc); mbar_init(bar_tempty(sm) + s, 8); }
    Ctl* C = ctl_of(sm);
    C->nstored = 0;
    C->pos_s = C->pos_ph = C->tpos_s = C->tpos_ph = 0;
    C->p = 1; C->j = 0; C->c = -1; C->aug = 0; C->done = 0;
    if (!(1 < G.P && 2 < G.nq)) {   // no chain-phase ring work
      C->aug = 1; C->p = G.P + 1;
      if (C->p >= G.nq) C->done = 1;
    }
    C->m = G.m; C->k = G.k; C->P = G.P; C->nq = G.nq; C->vec = G.vec;
#ifdef QR3_TIMING
    for (int i = 0; i < 40; i++) reinterpret_cast<long long*>(sm + OFF_STAT)[i] = 0;
#endif
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  if (warp < 4) chain_group(G, sm, A, Tg, tmap, warp, lane);
  else consumer(G, sm, A, Tg, tmap, 1, warp - 4, lane, RingPos{0, 0}, RingPos{0, 0});
}

}  // namespace qr3
}  // namespace sb200
