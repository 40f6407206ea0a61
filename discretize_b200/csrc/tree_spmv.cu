// K12: TreeMesh operators as CSR SpMV with the values generated on the fly.
//
// The reference assembles each OcTree operator as "per-entity loop -> COO -> CSR, then multiply
// by deflation matrices that eliminate hanging faces / edges / nodes" (tree_ext.pyx:3114-4762)
// and applies it with scipy csr_matvec (12 bytes per stored entry).  Here the deflated
// connectivity is uploaded once as int32 column indices plus ONE byte per entry; the byte
// indexes a small table of (multiplier, geometry selector) pairs and the value is
//     multiplier * geometry[selector][row]
// e.g. +-1/h_axis(cell) for face_divergence, +-w / face side length for edge_curl (w = product of
// the 1/2 hanging-edge weights), +-w / L for nodal_gradient, w for the averages: 5 bytes per entry.
//
// LPR lanes cooperate on a row (rows have 2..24 entries, so a full warp per row would idle);
// consecutive rows sit in consecutive lane groups, so index / code loads stay coalesced.
#include "tm_common.cuh"

namespace dzb {

struct CodeTable {
  double mult[256];
  unsigned char sel[256];
  int n;
};

template <int LPR, bool BYCOL, typename IP>
__global__ void __launch_bounds__(256)
    k_tree_spmv(i64 nrows, const IP *__restrict__ indptr, const int *__restrict__ indices,
                const unsigned char *__restrict__ codes, const double *__restrict__ tab_mult,
                const unsigned char *__restrict__ tab_sel, int ntab, const double *__restrict__ g0,
                const double *__restrict__ g1, const double *__restrict__ g2, const double *__restrict__ x,
                double *__restrict__ y) {
  __shared__ double s_mult[256];
  __shared__ unsigned char s_sel[256];
  for (int t = threadIdx.x; t < 256; t += blockDim.x) {
    s_mult[t] = t < ntab ? tab_mult[t] : 0.0;
    s_sel[t] = t < ntab ? tab_sel[t] : 0;
  }
  __syncthreads();
  const i64 gid = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  const i64 row = gid / LPR;
  const int l = (int)(gid % LPR);
  double acc = 0.0;
  if (row < nrows) {
    const i64 p0 = (i64)indptr[row], p1 = (i64)indptr[row + 1];
    double gr[3] = {1.0, 1.0, 1.0};
    if (!BYCOL) {
      if (g0) gr[0] = __ldg(g0 + row);
      if (g1) gr[1] = __ldg(g1 + row);
      if (g2) gr[2] = __ldg(g2 + row);
    }
    for (i64 p = p0 + l; p < p1; p += LPR) {
      const int c = __ldg(codes + p);
      const int col = __ldg(indices + p);
      const int s = s_sel[c];
      double g = 1.0;
      if (s) {
        if (BYCOL) g = __ldg((s == 1 ? g0 : s == 2 ? g1 : g2) + col);
        else g = gr[s - 1];
      }
      acc += (s_mult[c] * g) * __ldg(x + col);
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (row < nrows && l == 0) y[row] = acc;
}

// generic CSR with stored float64 values (scipy matrices mixed into operator algebra, full-tensor
// OcTree mass matrices): same lane-group mapping, values read instead of generated
template <int LPR, typename IP>
__global__ void __launch_bounds__(256)
    k_csr_spmv(i64 nrows, const IP *__restrict__ indptr, const int *__restrict__ indices, const double *__restrict__ vals,
               const double *__restrict__ x, double *__restrict__ y) {
  const i64 gid = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  const i64 row = gid / LPR;
  const int l = (int)(gid % LPR);
  double acc = 0.0;
  if (row < nrows) {
    const i64 p0 = (i64)indptr[row], p1 = (i64)indptr[row + 1];
    for (i64 p = p0 + l; p < p1; p += LPR) acc += __ldg(vals + p) * __ldg(x + __ldg(indices + p));
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (row < nrows && l == 0) y[row] = acc;
}

template <int LPR, typename IP>
static void launch_csr(i64 nrows, const void *indptr, const int *indices, const double *vals, const double *x, double *y,
                       cudaStream_t s) {
  const unsigned blocks = (unsigned)((nrows * LPR + 255) / 256);
  k_csr_spmv<LPR, IP><<<blocks, 256, 0, s>>>(nrows, (const IP *)indptr, indices, vals, x, y);
}

// y[i] = d[i] * x[i]  or  x[i] / d[i]
__global__ void k_diag_scale(i64 n, const double *__restrict__ d, const double *__restrict__ x, double *__restrict__ y,
                             int invert) {
  const i64 stride = (i64)gridDim.x * blockDim.x;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    y[i] = invert ? __ldg(x + i) / __ldg(d + i) : __ldg(d + i) * __ldg(x + i);
}

template <int LPR, bool BYCOL, typename IP>
static void launch_spmv(i64 nrows, const void *indptr, const int *indices, const unsigned char *codes,
                        const double *tm, const unsigned char *ts, int ntab, const double *g0, const double *g1,
                        const double *g2, const double *x, double *y, cudaStream_t s) {
  const i64 threads = nrows * LPR;
  const unsigned blocks = (unsigned)((threads + 255) / 256);
  k_tree_spmv<LPR, BYCOL, IP><<<blocks, 256, 0, s>>>(nrows, (const IP *)indptr, indices, codes, tm, ts, ntab, g0, g1, g2,
                                                     x, y);
}

template <bool BYCOL, typename IP>
static void dispatch_lpr(int lpr, i64 nrows, const void *indptr, const int *indices, const unsigned char *codes,
                         const double *tm, const unsigned char *ts, int ntab, const double *g0, const double *g1,
                         const double *g2, const double *x, double *y, cudaStream_t s) {
  switch (lpr) {
    case 1: launch_spmv<1, BYCOL, IP>(nrows, indptr, indices, codes, tm, ts, ntab, g0, g1, g2, x, y, s); break;
    case 2: launch_spmv<2, BYCOL, IP>(nrows, indptr, indices, codes, tm, ts, ntab, g0, g1, g2, x, y, s); break;
    case 4: launch_spmv<4, BYCOL, IP>(nrows, indptr, indices, codes, tm, ts, ntab, g0, g1, g2, x, y, s); break;
    case 8: launch_spmv<8, BYCOL, IP>(nrows, indptr, indices, codes, tm, ts, ntab, g0, g1, g2, x, y, s); break;
    case 16: launch_spmv<16, BYCOL, IP>(nrows, indptr, indices, codes, tm, ts, ntab, g0, g1, g2, x, y, s); break;
    default: launch_spmv<32, BYCOL, IP>(nrows, indptr, indices, codes, tm, ts, ntab, g0, g1, g2, x, y, s); break;
  }
}

}  // namespace dzb

using namespace dzb;

extern "C" int dzb_tree_spmv(int64_t nrows, const void *indptr, int indptr_is64, const int32_t *indices,
                             const uint8_t *codes, const double *tab_mult, const uint8_t *tab_sel, int ntab,
                             const double *g0, const double *g1, const double *g2, int geom_by_col, const double *x,
                             double *y, int lanes_per_row, void *stream) {
  if (nrows <= 0) return 0;
  if (ntab < 1 || ntab > 256) {
    set_error("dzb_tree_spmv: the code table must have 1..256 entries");
    return -1;
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (geom_by_col) {
    if (indptr_is64) dispatch_lpr<true, i64>(lanes_per_row, nrows, indptr, indices, codes, tab_mult, tab_sel, ntab, g0, g1, g2, x, y, s);
    else dispatch_lpr<true, int>(lanes_per_row, nrows, indptr, indices, codes, tab_mult, tab_sel, ntab, g0, g1, g2, x, y, s);
  } else {
    if (indptr_is64) dispatch_lpr<false, i64>(lanes_per_row, nrows, indptr, indices, codes, tab_mult, tab_sel, ntab, g0, g1, g2, x, y, s);
    else dispatch_lpr<false, int>(lanes_per_row, nrows, indptr, indices, codes, tab_mult, tab_sel, ntab, g0, g1, g2, x, y, s);
  }
  return check_launch("dzb_tree_spmv");
}

extern "C" int dzb_csr_spmv(int64_t nrows, const void *indptr, int indptr_is64, const int32_t *indices,
                            const double *values, const double *x, double *y, int lanes_per_row, void *stream) {
  if (nrows <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
#define DZB_CSR_CASE(L)                                                                   \
  case L:                                                                                 \
    if (indptr_is64) launch_csr<L, i64>(nrows, indptr, indices, values, x, y, s);         \
    else launch_csr<L, int>(nrows, indptr, indices, values, x, y, s);                     \
    break;
  switch (lanes_per_row) {
    DZB_CSR_CASE(1) DZB_CSR_CASE(2) DZB_CSR_CASE(4) DZB_CSR_CASE(8) DZB_CSR_CASE(16)
    default:
      if (indptr_is64) launch_csr<32, i64>(nrows, indptr, indices, values, x, y, s);
      else launch_csr<32, int>(nrows, indptr, indices, values, x, y, s);
  }
#undef DZB_CSR_CASE
  return check_launch("dzb_csr_spmv");
}

extern "C" int dzb_diag_scale(int64_t n, const double *d, const double *x, double *y, int invert, void *stream) {
  if (n <= 0) return 0;
  int sm = 148;
  const i64 need = (n + 255) / 256;
  const unsigned blocks = (unsigned)(need < (i64)sm * 16 ? need : (i64)sm * 16);
  k_diag_scale<<<blocks, 256, 0, (cudaStream_t)stream>>>(n, d, x, y, invert);
  return check_launch("dzb_diag_scale");
}
