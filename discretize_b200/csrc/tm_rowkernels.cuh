// Generic kernels over a row generator `gen` (see tm_rows.cuh): one thread per output
// row, 3-D thread blocks (32 x 4 x 4) so that y/z neighbours of a row hit L1.
//   k_apply : y[row] = sum_j val_j * x[col_j]
//   k_count : row_nnz[row] = number of stored entries
//   k_fill  : sorted column indices and values at indptr[row]
#pragma once
#include "tm_common.cuh"

namespace dzb {

constexpr int BX = 32, BY = 4, BZ = 4;

template <int LOC> static dim3 grid_for(const TM &m) {
  return dim3((unsigned)((m.sx<LOC>() + BX - 1) / BX), (unsigned)((m.sy<LOC>() + BY - 1) / BY),
              (unsigned)((m.sz<LOC>() + BZ - 1) / BZ));
}

// Apply kernels.  A thread owns the point (i, j) of ZT planes k = k0 + t * CBZ (t < ZT): the ZT rows are
// independent, the loop is fully unrolled and all their loads are in flight together -- at ZT = 1 the kernels
// were latency-bound (ncu round 1: DRAM traffic == algorithmic bytes, yet only 59 % of peak, long_scoreboard).
template <int CBX, int CBY, int CBZ, int CZT> struct ApplyCfg {
  static constexpr int X = CBX, Y = CBY, Z = CBZ, ZT = CZT, THREADS = CBX * CBY * CBZ;
};

template <class R, int COMP>
__device__ __forceinline__ void apply_one(const R &gen, const TM &m, i64 i, i64 j, i64 k, const double *__restrict__ x,
                                          double *__restrict__ y) {
  constexpr int LOC = R::template loc<COMP>();
  if (i >= m.sx<LOC>() || j >= m.sy<LOC>() || k >= m.sz<LOC>()) return;
  double acc = 0.0;
  gen.template row<COMP>(m, i, j, k, [&](i64 c, double v) { acc += v * __ldg(x + c); });
  y[R::template row_off<COMP>(m) + m.lidx<LOC>(i, j, k)] = acc;
}

template <class R, int COMP, class CFG>
__global__ void __launch_bounds__(CFG::THREADS) k_apply(R gen, TM m, const double *__restrict__ x, double *__restrict__ y) {
  const i64 i = (i64)blockIdx.x * CFG::X + threadIdx.x;
  const i64 j = (i64)blockIdx.y * CFG::Y + threadIdx.y;
  const i64 k0 = (i64)blockIdx.z * (CFG::Z * CFG::ZT) + threadIdx.z;
#pragma unroll
  for (int t = 0; t < CFG::ZT; ++t) apply_one<R, COMP>(gen, m, i, j, k0 + t * CFG::Z, x, y);
}

// All three row blocks of a vector-valued operator in ONE launch: thread (i,j,k) emits the row of every
// component that exists at (i,j,k).  The components read overlapping input entries (e.g. the three curl
// components share every edge value twice), so launching them together lets L1/L2 serve the re-reads and
// the input crosses HBM once instead of once per component.
template <class R, class CFG>
__global__ void __launch_bounds__(CFG::THREADS) k_apply3(R gen, TM m, const double *__restrict__ x, double *__restrict__ y) {
  const i64 i = (i64)blockIdx.x * CFG::X + threadIdx.x;
  const i64 j = (i64)blockIdx.y * CFG::Y + threadIdx.y;
  const i64 k0 = (i64)blockIdx.z * (CFG::Z * CFG::ZT) + threadIdx.z;
#pragma unroll
  for (int t = 0; t < CFG::ZT; ++t) {
    const i64 k = k0 + t * CFG::Z;
    apply_one<R, 0>(gen, m, i, j, k, x, y);
    apply_one<R, (R::NCOMP > 1 ? 1 : 0)>(gen, m, i, j, k, x, y);
    apply_one<R, (R::NCOMP > 2 ? 2 : 0)>(gen, m, i, j, k, x, y);
  }
}

template <class R, int COMP>
__global__ void __launch_bounds__(BX *BY *BZ) k_count(R gen, TM m, i64 *__restrict__ row_nnz) {
  constexpr int LOC = R::template loc<COMP>();
  const i64 i = (i64)blockIdx.x * BX + threadIdx.x;
  const i64 j = (i64)blockIdx.y * BY + threadIdx.y;
  const i64 k = (i64)blockIdx.z * BZ + threadIdx.z;
  if (i >= m.sx<LOC>() || j >= m.sy<LOC>() || k >= m.sz<LOC>()) return;
  int n = 0;
  gen.template row<COMP>(m, i, j, k, [&](i64, double) { ++n; });
  row_nnz[R::template row_off<COMP>(m) + m.lidx<LOC>(i, j, k)] = n;
}

template <class R, int COMP>
__global__ void __launch_bounds__(BX *BY *BZ)
    k_fill(R gen, TM m, const i64 *__restrict__ indptr, i64 *__restrict__ indices, double *__restrict__ data) {
  constexpr int LOC = R::template loc<COMP>();
  const i64 i = (i64)blockIdx.x * BX + threadIdx.x;
  const i64 j = (i64)blockIdx.y * BY + threadIdx.y;
  const i64 k = (i64)blockIdx.z * BZ + threadIdx.z;
  if (i >= m.sx<LOC>() || j >= m.sy<LOC>() || k >= m.sz<LOC>()) return;
  i64 cols[R::MAXNNZ];
  double vals[R::MAXNNZ];
  int n = 0;
  gen.template row<COMP>(m, i, j, k, [&](i64 c, double v) {
    // insertion into the sorted prefix (rows have <= 12 entries)
    int p = n++;
    while (p > 0 && cols[p - 1] > c) {
      cols[p] = cols[p - 1];
      vals[p] = vals[p - 1];
      --p;
    }
    cols[p] = c;
    vals[p] = v;
  });
  const i64 base = indptr[R::template row_off<COMP>(m) + m.lidx<LOC>(i, j, k)];
  for (int q = 0; q < n; ++q) {
    indices[base + q] = cols[q];
    data[base + q] = vals[q];
  }
}

enum Mode { MODE_APPLY, MODE_COUNT, MODE_FILL };

struct Args {
  const double *x;
  double *y;
  i64 *row_nnz;
  const i64 *indptr;
  i64 *indices;
  double *data;
};

template <int LOC, class CFG> static dim3 apply_grid(const TM &m) {
  return dim3((unsigned)((m.sx<LOC>() + CFG::X - 1) / CFG::X), (unsigned)((m.sy<LOC>() + CFG::Y - 1) / CFG::Y),
              (unsigned)((m.sz<LOC>() + CFG::Z * CFG::ZT - 1) / (CFG::Z * CFG::ZT)));
}

template <class R, int COMP, class CFG> static void launch_apply_comp(const R &gen, const TM &m, const Args &a, cudaStream_t s) {
  constexpr int LOC = R::template loc<COMP>();
  if (m.count<LOC>() == 0) return;
  k_apply<R, COMP, CFG><<<apply_grid<LOC, CFG>(m), dim3(CFG::X, CFG::Y, CFG::Z), 0, s>>>(gen, m, a.x, a.y);
}

template <class R, class CFG> static void launch_apply_cfg(const R &gen, const TM &m, const Args &a, cudaStream_t s) {
  if (R::NCOMP == 3) {  // node-extent grid covers every component's extent
    k_apply3<R, CFG><<<apply_grid<LOC_N, CFG>(m), dim3(CFG::X, CFG::Y, CFG::Z), 0, s>>>(gen, m, a.x, a.y);
    return;
  }
  launch_apply_comp<R, 0, CFG>(gen, m, a, s);
  if (R::NCOMP > 1) launch_apply_comp<R, (R::NCOMP > 1 ? 1 : 0), CFG>(gen, m, a, s);
}

// launch configuration of the apply kernels (dzb_tune_elementary; default chosen by tools/sweep_elementary.py)
extern int g_apply_cfg;

template <class R> static void launch_apply(const R &gen, const TM &m, const Args &a, cudaStream_t s) {
  switch (g_apply_cfg) {
#ifdef DZB_APPLY_SWEEP
    case 0: launch_apply_cfg<R, ApplyCfg<32, 4, 4, 1>>(gen, m, a, s); break;
    case 1: launch_apply_cfg<R, ApplyCfg<32, 4, 4, 2>>(gen, m, a, s); break;
    case 3: launch_apply_cfg<R, ApplyCfg<64, 4, 2, 4>>(gen, m, a, s); break;
    case 4: launch_apply_cfg<R, ApplyCfg<128, 2, 2, 4>>(gen, m, a, s); break;
    case 5: launch_apply_cfg<R, ApplyCfg<32, 8, 2, 4>>(gen, m, a, s); break;
    case 6: launch_apply_cfg<R, ApplyCfg<64, 2, 2, 4>>(gen, m, a, s); break;
    case 7: launch_apply_cfg<R, ApplyCfg<32, 4, 2, 4>>(gen, m, a, s); break;
    case 8: launch_apply_cfg<R, ApplyCfg<32, 4, 4, 8>>(gen, m, a, s); break;
    case 9: launch_apply_cfg<R, ApplyCfg<32, 2, 4, 4>>(gen, m, a, s); break;
#endif
    default: launch_apply_cfg<R, ApplyCfg<32, 4, 4, 4>>(gen, m, a, s); break;
  }
}

template <class R, int COMP> static void launch_comp(const R &gen, int mode, const TM &m, const Args &a, cudaStream_t s) {
  constexpr int LOC = R::template loc<COMP>();
  if (m.count<LOC>() == 0) return;
  dim3 g = grid_for<LOC>(m), b(BX, BY, BZ);
  if (mode == MODE_COUNT) k_count<R, COMP><<<g, b, 0, s>>>(gen, m, a.row_nnz);
  else k_fill<R, COMP><<<g, b, 0, s>>>(gen, m, a.indptr, a.indices, a.data);
}

template <class R> static void launch_all(const R &gen, int mode, const TM &m, const Args &a, cudaStream_t s) {
  if (mode == MODE_APPLY) {
    launch_apply(gen, m, a, s);
    return;
  }
  launch_comp<R, 0>(gen, mode, m, a, s);
  if (R::NCOMP > 1) launch_comp<R, (R::NCOMP > 1 ? 1 : 0)>(gen, mode, m, a, s);
  if (R::NCOMP > 2) launch_comp<R, (R::NCOMP > 2 ? 2 : 0)>(gen, mode, m, a, s);
}

}  // namespace dzb
