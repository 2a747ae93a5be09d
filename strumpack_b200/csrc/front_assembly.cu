// strumpack_b200 -- device-resident extend-add of contribution blocks into a
// parent front.  Replaces (reference) extend_add_kernel,
// src/sparse/fronts/FrontCUDA.cu:111-148, and the host loops of
// FrontBLR::extend_add / F22blr extend-add, src/sparse/fronts/FrontBLR.cpp:338-403:
//     F(I[y], I[x]) += CB(y, x)     for every entry of a child's contribution block,
// the parent front stored as its four blocks F11 (d1 x d1), F12 (d1 x d2),
// F21 (d2 x d1), F22 (d2 x d2), I the child's update indices in the parent's
// numbering (0 .. d1+d2-1).  HBM-bound: every CB entry is read once (coalesced
// along the rows), every target entry read-modified-written once.  One launch per
// side (left / right child): the two children may hit the same entries, the
// entries of one child never collide (I is injective).
#include "../../include/sb200_structured.h"
#include "sb200_common.cuh"

namespace sb200 {
namespace {

__global__ void __launch_bounds__(256)
extend_add_kernel(int nf, const SB200FrontAssemble* __restrict__ fronts, int right) {
  const int f = blockIdx.z;
  if (f >= nf) return;
  const SB200FrontAssemble F = fronts[f];
  const double* CB = right ? F.CB2 : F.CB1;
  const int* I = right ? F.I2 : F.I1;
  const int n = right ? F.dCB2 : F.dCB1;
  if (!CB || n <= 0) return;
  const int y = blockIdx.x * 32 + (threadIdx.x & 31);
  if (y >= n) return;
  const int Iy = I[y];
  const int d1 = F.d1, d2 = F.d2;
  // row Iy of the parent lives in (F11 | F12) or (F21 | F22)
  double* left = Iy < d1 ? F.F11 + Iy : F.F21 + (Iy - d1);
  double* rght = Iy < d1 ? F.F12 + Iy : F.F22 + (Iy - d1);
  const int ld = Iy < d1 ? d1 : d2;
  for (int x = blockIdx.y * 8 + (threadIdx.x >> 5); x < n; x += gridDim.y * 8) {
    const int Ix = I[x];
    const double v = CB[y + (size_t)x * n];
    if (Ix < d1) left[(size_t)Ix * ld] += v;
    else rght[(size_t)(Ix - d1) * ld] += v;
  }
}

}  // namespace
}  // namespace sb200

extern "C" int SB200_d_front_extend_add_device(int nf, const SB200FrontAssemble* d_fronts, int max_dCB,
                                               void* stream) {
  if (nf <= 0 || max_dCB <= 0) return 0;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    std::fprintf(stderr, "Operation failed: no CUDA device: strumpack_b200 has no CPU fallback (sm_100a only)\n");
    return 1;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const dim3 grid((max_dCB + 31) / 32, std::min((max_dCB + 7) / 8, 64), nf);
  for (int right = 0; right < 2; right++)
    sb200::extend_add_kernel<<<grid, 256, 0, st>>>(nf, d_fronts, right);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    std::fprintf(stderr, "Operation failed: extend_add: %s\n", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}
