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

// ---- coalesced variants ------------------------------------------------------------------------------------------
// ncu (profiles/r1b_k12_*): the thread-per-row kernel above moves only 1.2x the algorithmic bytes through DRAM but is
// latency-bound on its dependent loads (row pointer -> index -> gather), all of them strided across the warp.

__device__ __forceinline__ double coded_value(int c, int col, const double *s_mult, const unsigned char *s_sel, bool bycol,
                                              const double (&gr)[3], const double *__restrict__ g0,
                                              const double *__restrict__ g1, const double *__restrict__ g2) {
  const int s = s_sel[c];
  double g = 1.0;
  if (s) g = bycol ? __ldg((s == 1 ? g0 : s == 2 ? g1 : g2) + col) : gr[s - 1];
  return s_mult[c] * g;
}

constexpr int ELL_PAD = 255;  // code of a padding slot (rows shorter than the ELL width)

// ELL (column-major) layout: entry e of row r sits at e*stride + r, so a warp reads 32 consecutive indices / codes
// per step and needs no row pointers.  On a balanced OcTree every cell-rowed operator has rows of ONE length
// (face_divergence 6, average_face / edge / node_to_cell 6 / 12 / 8); edge_curl rows hold 4..6 entries and are
// padded with ELL_PAD slots.  KT > 0: row length known at compile time -- fully unrolled, so the KT index loads and
// then the KT gathers of a row are all in flight together (the gathers are latency-bound: +18..42 % over run-time K).
template <bool BYCOL, int KT, bool PADDED>
__global__ void __launch_bounds__(256)
    k_tree_ell(i64 nrows, int K, i64 stride, const int *__restrict__ indices, const unsigned char *__restrict__ codes,
               const double *__restrict__ tab_mult, const unsigned char *__restrict__ tab_sel, int ntab,
               const double *__restrict__ g0, const double *__restrict__ g1, const double *__restrict__ g2,
               const int *__restrict__ perm, const double *__restrict__ x, double *__restrict__ y) {
  // perm != NULL: the table rows (and the row geometry) are stored in an order that follows the COLUMN numbering, so that
  // the gathers of a warp share sectors; table row r is row perm[r] of the operator
  __shared__ double s_mult[256];
  __shared__ unsigned char s_sel[256];
  for (int t = threadIdx.x; t < 256; t += blockDim.x) {
    s_mult[t] = t < ntab ? tab_mult[t] : 0.0;
    s_sel[t] = t < ntab ? tab_sel[t] : 0;
  }
  __syncthreads();
  const i64 row = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nrows) return;
  double gr[3] = {1.0, 1.0, 1.0};
  if (!BYCOL) {
    if (g0) gr[0] = __ldg(g0 + row);
    if (g1) gr[1] = __ldg(g1 + row);
    if (g2) gr[2] = __ldg(g2 + row);
  }
  double acc = 0.0;
  const int *pi = indices + row;
  const unsigned char *pc = codes + row;
  if (KT > 0) {
    int col[KT > 0 ? KT : 1], c[KT > 0 ? KT : 1];
    double xv[KT > 0 ? KT : 1];
#pragma unroll
    for (int e = 0; e < KT; ++e) {
      col[e] = __ldg(pi + e * stride);
      c[e] = __ldg(pc + e * stride);
    }
#pragma unroll
    for (int e = 0; e < KT; ++e) xv[e] = (!PADDED || c[e] != ELL_PAD) ? __ldg(x + col[e]) : 0.0;
#pragma unroll
    for (int e = 0; e < KT; ++e)
      if (!PADDED || c[e] != ELL_PAD) acc += coded_value(c[e], col[e], s_mult, s_sel, BYCOL, gr, g0, g1, g2) * xv[e];
  } else {
#pragma unroll 4
    for (int e = 0; e < K; ++e) {
      const int col = __ldg(pi);
      const int c = __ldg(pc);
      if (!PADDED || c != ELL_PAD) acc += coded_value(c, col, s_mult, s_sel, BYCOL, gr, g0, g1, g2) * __ldg(x + col);
      pi += stride;
      pc += stride;
    }
  }
  y[perm ? (i64)__ldg(perm + row) : row] = acc;
}

// Sliced ELL (SELL-32) for rows of DIFFERENT lengths (the transposed operators, nodal_gradient: 1..24 entries): rows are
// cut into slices of 32 (one warp); slice s stores its entries column-major with the slice's own width
// K_s = (slice_ptr[s+1] - slice_ptr[s]) / 32, entry e of the slice's row l at slice_ptr[s] + 32 e + l, shorter rows padded
// with ELL_PAD codes.  Index and code loads are coalesced like ELL's and there is no row-pointer load in front of them (the
// CSR kernel's dependent chain is indptr -> index -> gather); on an OcTree most slices are uniform, so the padding is a
// few per cent.  Entries are processed four at a time: four index loads, then four gathers in flight.
template <bool BYCOL, typename SP>
__global__ void __launch_bounds__(256)
    k_tree_sell(i64 nrows, i64 nslices, const SP *__restrict__ slice_ptr, const int *__restrict__ indices,
                const unsigned char *__restrict__ codes, const double *__restrict__ tab_mult,
                const unsigned char *__restrict__ tab_sel, int ntab, const double *__restrict__ g0,
                const double *__restrict__ g1, const double *__restrict__ g2, const int *__restrict__ perm,
                const double *__restrict__ x, double *__restrict__ y) {
  __shared__ double s_mult[256];
  __shared__ unsigned char s_sel[256];
  for (int t = threadIdx.x; t < 256; t += blockDim.x) {
    s_mult[t] = t < ntab ? tab_mult[t] : 0.0;
    s_sel[t] = t < ntab ? tab_sel[t] : 0;
  }
  __syncthreads();
  const i64 row = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  const i64 slice = row >> 5;
  if (slice >= nslices) return;
  const i64 base = (i64)__ldg(slice_ptr + slice);
  const int K = (int)(((i64)__ldg(slice_ptr + slice + 1) - base) >> 5);
  const bool live = row < nrows;
  double gr[3] = {1.0, 1.0, 1.0};
  if (!BYCOL && live) {
    if (g0) gr[0] = __ldg(g0 + row);
    if (g1) gr[1] = __ldg(g1 + row);
    if (g2) gr[2] = __ldg(g2 + row);
  }
  const int *pi = indices + base + (threadIdx.x & 31);
  const unsigned char *pc = codes + base + (threadIdx.x & 31);
  double acc = 0.0;
  int e = 0;
  for (; e + 4 <= K; e += 4) {
    int col[4], c[4];
    double xv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      col[u] = __ldg(pi + (e + u) * 32);
      c[u] = __ldg(pc + (e + u) * 32);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) xv[u] = c[u] != ELL_PAD ? __ldg(x + col[u]) : 0.0;
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (c[u] != ELL_PAD) acc += coded_value(c[u], col[u], s_mult, s_sel, BYCOL, gr, g0, g1, g2) * xv[u];
  }
  if (e < K) {   // 1..3 entries left: loads first, predicated
    int col[3], c[3];
    double xv[3];
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const bool on = e + u < K;
      col[u] = on ? __ldg(pi + (e + u) * 32) : 0;
      c[u] = on ? (int)__ldg(pc + (e + u) * 32) : ELL_PAD;
    }
#pragma unroll
    for (int u = 0; u < 3; ++u) xv[u] = c[u] != ELL_PAD ? __ldg(x + col[u]) : 0.0;
#pragma unroll
    for (int u = 0; u < 3; ++u)
      if (c[u] != ELL_PAD) acc += coded_value(c[u], col[u], s_mult, s_sel, BYCOL, gr, g0, g1, g2) * xv[u];
  }
  if (live) y[perm ? (i64)__ldg(perm + row) : row] = acc;
}

// Overflow rows of the hybrid layout (ELL of width K0 + the entries beyond K0 of the few longer rows): one thread per
// overflow row, y[row] += sum of its extra entries.  Launched after the ELL kernel on the same stream.
template <bool BYCOL, typename IP>
__global__ void __launch_bounds__(256)
    k_tree_rows_add(i64 nrem, const int *__restrict__ rows, const IP *__restrict__ indptr, const int *__restrict__ indices,
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
  const i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrem) return;
  const i64 row = __ldg(rows + r);
  double gr[3] = {1.0, 1.0, 1.0};
  if (!BYCOL) {
    if (g0) gr[0] = __ldg(g0 + row);
    if (g1) gr[1] = __ldg(g1 + row);
    if (g2) gr[2] = __ldg(g2 + row);
  }
  double acc = 0.0;
  for (i64 p = (i64)indptr[r], p1 = (i64)indptr[r + 1]; p < p1; ++p) {
    const int c = __ldg(codes + p), col = __ldg(indices + p);
    acc += coded_value(c, col, s_mult, s_sel, BYCOL, gr, g0, g1, g2) * __ldg(x + col);
  }
  y[row] += acc;
}

// CSR with warp-staged entries: a warp owns 32 consecutive rows; the contiguous span of their entries is copied
// to shared memory with coalesced loads, then every lane walks its own row out of shared memory.  (Measured SLOWER
// than the plain kernel on OcTree operators -- the gathers, not the index loads, set the pace -- kept for comparison.)
constexpr int STREAM_CAP = 512;   // entries staged per warp (spans longer than this read global memory directly)

template <bool BYCOL, typename IP>
__global__ void __launch_bounds__(256)
    k_tree_stream(i64 nrows, const IP *__restrict__ indptr, const int *__restrict__ indices,
                  const unsigned char *__restrict__ codes, const double *__restrict__ tab_mult,
                  const unsigned char *__restrict__ tab_sel, int ntab, const double *__restrict__ g0,
                  const double *__restrict__ g1, const double *__restrict__ g2, const double *__restrict__ x,
                  double *__restrict__ y) {
  __shared__ double s_mult[256];
  __shared__ unsigned char s_sel[256];
  __shared__ int s_idx[8][STREAM_CAP];
  __shared__ unsigned char s_code[8][STREAM_CAP];
  for (int t = threadIdx.x; t < 256; t += blockDim.x) {
    s_mult[t] = t < ntab ? tab_mult[t] : 0.0;
    s_sel[t] = t < ntab ? tab_sel[t] : 0;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  int *my_idx = s_idx[wib];
  unsigned char *my_code = s_code[wib];
  for (i64 row0 = ((i64)blockIdx.x * 8 + wib) * 32; row0 < nrows; row0 += (i64)gridDim.x * 256) {
    const i64 row = row0 + lane;
    const bool live = row < nrows;
    const i64 p0 = (i64)indptr[live ? row : nrows], p1 = live ? (i64)indptr[row + 1] : p0;
    const i64 start = __shfl_sync(0xffffffffu, p0, 0), end = __shfl_sync(0xffffffffu, p1, 31);
    const i64 span = end - start;
    const bool staged = span <= STREAM_CAP;
    if (staged) {
      for (i64 t = lane; t < span; t += 32) {
        my_idx[t] = __ldg(indices + start + t);
        my_code[t] = __ldg(codes + start + t);
      }
    }
    __syncwarp();
    if (live) {
      double gr[3] = {1.0, 1.0, 1.0};
      if (!BYCOL) {
        if (g0) gr[0] = __ldg(g0 + row);
        if (g1) gr[1] = __ldg(g1 + row);
        if (g2) gr[2] = __ldg(g2 + row);
      }
      double acc = 0.0;
      for (i64 p = p0; p < p1; ++p) {
        const int col = staged ? my_idx[p - start] : __ldg(indices + p);
        const int c = staged ? my_code[p - start] : __ldg(codes + p);
        acc += coded_value(c, col, s_mult, s_sel, BYCOL, gr, g0, g1, g2) * __ldg(x + col);
      }
      y[row] = acc;
    }
    __syncwarp();
  }
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
  if (lanes_per_row == 0) {   // warp-staged CSR
    int sm = 148;
    const i64 need = (nrows + 255) / 256;
    const unsigned blocks = (unsigned)(need < (i64)sm * 32 ? need : (i64)sm * 32);
    if (geom_by_col) {
      if (indptr_is64) k_tree_stream<true, i64><<<blocks, 256, 0, s>>>(nrows, (const i64 *)indptr, indices, codes, tab_mult, tab_sel, ntab, g0, g1, g2, x, y);
      else k_tree_stream<true, int><<<blocks, 256, 0, s>>>(nrows, (const int *)indptr, indices, codes, tab_mult, tab_sel, ntab, g0, g1, g2, x, y);
    } else {
      if (indptr_is64) k_tree_stream<false, i64><<<blocks, 256, 0, s>>>(nrows, (const i64 *)indptr, indices, codes, tab_mult, tab_sel, ntab, g0, g1, g2, x, y);
      else k_tree_stream<false, int><<<blocks, 256, 0, s>>>(nrows, (const int *)indptr, indices, codes, tab_mult, tab_sel, ntab, g0, g1, g2, x, y);
    }
    return check_launch("dzb_tree_spmv (stream)");
  }
  if (geom_by_col) {
    if (indptr_is64) dispatch_lpr<true, i64>(lanes_per_row, nrows, indptr, indices, codes, tab_mult, tab_sel, ntab, g0, g1, g2, x, y, s);
    else dispatch_lpr<true, int>(lanes_per_row, nrows, indptr, indices, codes, tab_mult, tab_sel, ntab, g0, g1, g2, x, y, s);
  } else {
    if (indptr_is64) dispatch_lpr<false, i64>(lanes_per_row, nrows, indptr, indices, codes, tab_mult, tab_sel, ntab, g0, g1, g2, x, y, s);
    else dispatch_lpr<false, int>(lanes_per_row, nrows, indptr, indices, codes, tab_mult, tab_sel, ntab, g0, g1, g2, x, y, s);
  }
  return check_launch("dzb_tree_spmv");
}

extern "C" int dzb_tree_ell_spmv_perm(int64_t nrows, int row_len, int64_t stride, const int32_t *indices,
                                      const uint8_t *codes, const double *tab_mult, const uint8_t *tab_sel, int ntab,
                                      const double *g0, const double *g1, const double *g2, int geom_by_col, int padded,
                                      const int32_t *perm, const double *x, double *y, void *stream);

extern "C" int dzb_tree_ell_spmv(int64_t nrows, int row_len, int64_t stride, const int32_t *indices, const uint8_t *codes,
                                 const double *tab_mult, const uint8_t *tab_sel, int ntab, const double *g0,
                                 const double *g1, const double *g2, int geom_by_col, int padded, const double *x,
                                 double *y, void *stream) {
  return dzb_tree_ell_spmv_perm(nrows, row_len, stride, indices, codes, tab_mult, tab_sel, ntab, g0, g1, g2, geom_by_col,
                                padded, nullptr, x, y, stream);
}

extern "C" int dzb_tree_ell_spmv_perm(int64_t nrows, int row_len, int64_t stride, const int32_t *indices,
                                      const uint8_t *codes, const double *tab_mult, const uint8_t *tab_sel, int ntab,
                                      const double *g0, const double *g1, const double *g2, int geom_by_col, int padded,
                                      const int32_t *perm, const double *x, double *y, void *stream) {
  if (nrows <= 0) return 0;
  if (ntab < 1 || ntab > 255 || row_len < 0 || stride < nrows) {
    set_error("dzb_tree_ell_spmv: bad table size (1..255; code 255 marks padding), row length or stride");
    return -1;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned blocks = (unsigned)((nrows + 255) / 256);
#define DZB_ELL_LAUNCH(BC, KT, PD) \
  k_tree_ell<BC, KT, PD><<<blocks, 256, 0, s>>>(nrows, row_len, stride, indices, codes, tab_mult, tab_sel, ntab, g0, g1, g2, perm, x, y)
  if (geom_by_col) {
    switch (row_len) {
      case 2: DZB_ELL_LAUNCH(true, 2, true); break;
      case 4: DZB_ELL_LAUNCH(true, 4, true); break;
      default: DZB_ELL_LAUNCH(true, 0, true); break;
    }
  } else if (padded) {
    switch (row_len) {
      case 2: DZB_ELL_LAUNCH(false, 2, true); break;
      case 3: DZB_ELL_LAUNCH(false, 3, true); break;
      case 4: DZB_ELL_LAUNCH(false, 4, true); break;
      case 6: DZB_ELL_LAUNCH(false, 6, true); break;
      case 8: DZB_ELL_LAUNCH(false, 8, true); break;
      default: DZB_ELL_LAUNCH(false, 0, true); break;
    }
  } else {
    switch (row_len) {
      case 6: DZB_ELL_LAUNCH(false, 6, false); break;
      case 8: DZB_ELL_LAUNCH(false, 8, false); break;
      case 12: DZB_ELL_LAUNCH(false, 12, false); break;
      default: DZB_ELL_LAUNCH(false, 0, false); break;
    }
  }
#undef DZB_ELL_LAUNCH
  return check_launch("dzb_tree_ell_spmv");
}

extern "C" int dzb_tree_sell_spmv(int64_t nrows, const void *slice_ptr, int slice_ptr_is64, const int32_t *indices,
                                  const uint8_t *codes, const double *tab_mult, const uint8_t *tab_sel, int ntab,
                                  const double *g0, const double *g1, const double *g2, int geom_by_col,
                                  const int32_t *perm, const double *x, double *y, void *stream) {
  if (nrows <= 0) return 0;
  if (ntab < 1 || ntab > 255) {
    set_error("dzb_tree_sell_spmv: the code table must have 1..255 entries (code 255 marks padding)");
    return -1;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const i64 nslices = (nrows + 31) / 32;
  const unsigned blocks = (unsigned)((nslices * 32 + 255) / 256);
#define DZB_SELL_LAUNCH(BC, SP) \
  k_tree_sell<BC, SP><<<blocks, 256, 0, s>>>(nrows, nslices, (const SP *)slice_ptr, indices, codes, tab_mult, tab_sel, ntab, g0, g1, g2, perm, x, y)
  if (geom_by_col) {
    if (slice_ptr_is64) DZB_SELL_LAUNCH(true, i64);
    else DZB_SELL_LAUNCH(true, int);
  } else {
    if (slice_ptr_is64) DZB_SELL_LAUNCH(false, i64);
    else DZB_SELL_LAUNCH(false, int);
  }
#undef DZB_SELL_LAUNCH
  return check_launch("dzb_tree_sell_spmv");
}

// y[rows[r]] += the entries indptr[r] .. indptr[r+1] of overflow row r (hybrid layout: call after dzb_tree_ell_spmv)
extern "C" int dzb_tree_rows_add(int64_t nrem, const int32_t *rows, const void *indptr, int indptr_is64,
                                 const int32_t *indices, const uint8_t *codes, const double *tab_mult,
                                 const uint8_t *tab_sel, int ntab, const double *g0, const double *g1, const double *g2,
                                 int geom_by_col, const double *x, double *y, void *stream) {
  if (nrem <= 0) return 0;
  if (ntab < 1 || ntab > 256) {
    set_error("dzb_tree_rows_add: the code table must have 1..256 entries");
    return -1;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned blocks = (unsigned)((nrem + 255) / 256);
#define DZB_REM_LAUNCH(BC, IP) \
  k_tree_rows_add<BC, IP><<<blocks, 256, 0, s>>>(nrem, rows, (const IP *)indptr, indices, codes, tab_mult, tab_sel, ntab, g0, g1, g2, x, y)
  if (geom_by_col) {
    if (indptr_is64) DZB_REM_LAUNCH(true, i64);
    else DZB_REM_LAUNCH(true, int);
  } else {
    if (indptr_is64) DZB_REM_LAUNCH(false, i64);
    else DZB_REM_LAUNCH(false, int);
  }
#undef DZB_REM_LAUNCH
  return check_launch("dzb_tree_rows_add");
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
