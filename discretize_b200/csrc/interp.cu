// K10/K11: TensorMesh.get_interpolation_matrix on the GPU.
//
// reference: base/base_tensor_mesh.py:684-790 -> utils/interpolation_utils.py:44-172 ->
// _extensions/interputils_cython.pyx:42-182 (_bisect_right, _get_inds_ws, _interpmat3D).
//
// Point location is three bisections over the *same float64 1-D arrays the host built*
// (node / cell-centre coordinates), with the reference's exact comparison `x < a[mid]`,
// so indices are bit-exact; weights use the reference's expression order.
#include "tm_common.cuh"

namespace dzb {

constexpr int IT = 128;  // threads (= points) per tile
#ifndef LOCATE_MINB
#define LOCATE_MINB 8   // 64 registers; capping at 40 (12 blocks) spills and runs 17 % slower
#endif

struct AxisHit {
  int i1, i2;        // positions on one axis: 32-bit (an axis has far fewer than 2^31 entries; 64-bit index arithmetic made
  double w1, w2;     // the locate kernels issue-bound: 414 instructions per point)
};

// _bisect_right + _get_inds_ws (interputils_cython.pyx:42-67).
// bisect_right(a, x) is the unique r with a[r-1] <= x < a[r] (a sorted); any search that ends on that r
// is bit-identical to the reference's bisection.  `inv_mean` = (n-1)/(a[n-1]-a[0]) gives a guess that is
// exact or off by a few entries on smooth meshes; a short walk fixes it, and a bisection over the
// remaining bracket handles arbitrary (strongly graded) arrays.
__device__ __forceinline__ AxisHit locate_axis(const double *__restrict__ a, i64 n64, double xp, double inv_mean) {
  const int n = (int)n64;
  int lo = 0, hi = n;
  if (n >= 8 && xp == xp) {
    double g = (xp - a[0]) * inv_mean;
    g = fmin(fmax(g, -1.0), (double)n);
    int r = (int)g + 1;                       // guess for bisect_right
    r = max(min(r, n), 0);
    // walk at most 4 entries towards the answer, then keep the tight bracket for the bisection
    int steps = 0;
    while (steps < 4) {
      if (r < n && !(xp < a[r])) ++r;         // a[r] <= x  -> answer is right of r
      else if (r > 0 && xp < a[r - 1]) --r;   // x < a[r-1] -> answer is left of r
      else { lo = hi = r; break; }
      ++steps;
    }
    if (lo != hi) {                           // not settled: bracket by which side we were moving to
      if (r < n && !(xp < a[r])) lo = r + 1;  // still a[r] <= x
      else if (r > 0 && xp < a[r - 1]) hi = r - 1;
      else lo = hi = r;
    }
  }
  while (lo < hi) {
    const int mid = (lo + hi) / 2;
    if (xp < a[mid]) hi = mid;
    else lo = mid + 1;
  }
  AxisHit h;
  h.i2 = max(min(lo, n - 1), 0);
  h.i1 = max(min(lo - 1, n - 1), 0);
  if (h.i1 == h.i2) h.w1 = 0.5;
  else h.w1 = (a[h.i2] - xp) / (a[h.i2] - a[h.i1]);
  h.w2 = 1 - h.w1;
  return h;
}

struct Axes {
  const double *tx, *ty, *tz;
  i64 ntx, nty, ntz;
  double ix, iy, iz;  // (n-1)/(a[n-1]-a[0]) per axis (0 disables the guess)
};

__device__ __forceinline__ void axes_means(Axes &a) {
  a.ix = (a.ntx >= 8 && a.tx[a.ntx - 1] > a.tx[0]) ? (double)(a.ntx - 1) / (a.tx[a.ntx - 1] - a.tx[0]) : 0.0;
  a.iy = (a.nty >= 8 && a.ty[a.nty - 1] > a.ty[0]) ? (double)(a.nty - 1) / (a.ty[a.nty - 1] - a.ty[0]) : 0.0;
  a.iz = (a.ntz >= 8 && a.tz[a.ntz - 1] > a.tz[0]) ? (double)(a.ntz - 1) / (a.tz[a.ntz - 1] - a.tz[0]) : 0.0;
}

// copies the three 1-D arrays into shared memory when they fit, else uses them in place
__device__ __forceinline__ Axes stage_axes(const Axes &g, double *smem, i64 smem_doubles) {
  Axes a = g;
  if (g.ntx + g.nty + g.ntz <= smem_doubles) {
    for (i64 t = threadIdx.x; t < g.ntx; t += blockDim.x) smem[t] = g.tx[t];
    for (i64 t = threadIdx.x; t < g.nty; t += blockDim.x) smem[g.ntx + t] = g.ty[t];
    for (i64 t = threadIdx.x; t < g.ntz; t += blockDim.x) smem[g.ntx + g.nty + t] = g.tz[t];
    a.tx = smem;
    a.ty = smem + g.ntx;
    a.tz = smem + g.ntx + g.nty;
    __syncthreads();
  }
  axes_means(a);
  return a;
}

// entry order of _interpmat3D (interputils_cython.pyx:145-180): e = 0..7 ->
// x: 1,1,2,2,1,1,2,2   y: 1,2,1,2,1,2,1,2   z: 1,1,1,1,2,2,2,2
__device__ __forceinline__ void entries(const AxisHit &hx, const AxisHit &hy, const AxisHit &hz, i64 nx, i64 ny,
                                        i64 col_offset, i64 (&col)[8], double (&w)[8]) {
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const bool x2 = (e >> 1) & 1, y2 = e & 1, z2 = (e >> 2) & 1;
    const i64 ix = x2 ? hx.i2 : hx.i1, iy = y2 ? hy.i2 : hy.i1, iz = z2 ? hz.i2 : hz.i1;   // 64-bit from here on
    col[e] = col_offset + ix + nx * (iy + ny * iz);
    w[e] = ((x2 ? hx.w2 : hx.w1) * (y2 ? hy.w2 : hy.w1)) * (z2 ? hz.w2 : hz.w1);
  }
}

// K10: locate + weights.  Persistent tiles of IT points; the (npts x 8) tables are transposed through shared
// memory so that they are written with fully coalesced stores.  Column indices are int32 whenever the operator has
// fewer than 2^31 columns (like scipy's index arrays): 96 instead of 128 bytes written per point.  The axes are
// read through L1 (after the index guess a point touches ~3 entries per axis), so a block needs only the two
// transposition buffers (13.8 KB) and 16 blocks = 64 warps stay resident per SM (the first version staged the
// axes in 32 KB of shared memory and ran at 25 % occupancy: ncu, profiles/r1b_*).
template <typename IDX>
__global__ void __launch_bounds__(IT, LOCATE_MINB) k_interp_locate(Axes g, const double *__restrict__ pts, i64 npts, i64 col_offset,
                                                      IDX *__restrict__ cols, double *__restrict__ w) {
  extern __shared__ double smem[];
  double *sw = smem;                          // IT*9 weights (padded rows)
  IDX *sc = (IDX *)(smem + IT * 9);           // IT*9 columns
  Axes a = g;
  axes_means(a);
  const i64 ntiles = (npts + IT - 1) / IT;
  for (i64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const i64 p = tile * IT + threadIdx.x;
    if (p < npts) {
      const AxisHit hx = locate_axis(a.tx, a.ntx, __ldg(pts + 3 * p), a.ix);
      const AxisHit hy = locate_axis(a.ty, a.nty, __ldg(pts + 3 * p + 1), a.iy);
      const AxisHit hz = locate_axis(a.tz, a.ntz, __ldg(pts + 3 * p + 2), a.iz);
      i64 c8[8];
      double w8[8];
      entries(hx, hy, hz, a.ntx, a.nty, col_offset, c8, w8);
#pragma unroll
      for (int e = 0; e < 8; ++e) {  // row pitch 9: at most 2-way bank conflicts
        sw[threadIdx.x * 9 + e] = w8[e];
        sc[threadIdx.x * 9 + e] = (IDX)c8[e];
      }
    }
    __syncthreads();
    const i64 base = tile * IT * 8;
    const i64 lim = min((i64)IT * 8, npts * 8 - base);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const i64 t = threadIdx.x + r * IT;
      if (t < lim) {
        const i64 q = (t >> 3) * 9 + (t & 7);
        w[base + t] = sw[q];
        cols[base + t] = sc[q];
      }
    }
    __syncthreads();
  }
}

// K11: y[p] = scale[p] * sum_e w[p,e] v[cols[p,e]].  Eight lanes per point: the tables are
// read fully coalesced, the gathers run 32-wide, partial sums combine with shuffles.
// With `perm` the table rows are in BUCKET order (sorted by the flat index of their first grid sample, built once by
// the host side): consecutive rows then gather from neighbouring grid entries, so the gathers stream through the grid
// vector instead of paying one 128-byte DRAM fetch per 8-byte value (ncu: 635 B / point unsorted); row p of the table
// belongs to point perm[p], where the result is scattered.  `scale` is in table order.
template <typename IDX>
__global__ void __launch_bounds__(256) k_interp_apply(const IDX *__restrict__ cols, const double *__restrict__ w,
                                                      const double *__restrict__ scale, const int *__restrict__ perm,
                                                      i64 npts, const double *__restrict__ v, double *__restrict__ y) {
  const i64 nent = npts * 8;
  const i64 stride = (i64)gridDim.x * blockDim.x;
  const i64 nent_pad = (nent + 31) / 32 * 32;
  for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < nent_pad; t += stride) {
    double part = 0.0;
    if (t < nent) part = __ldg(w + t) * __ldg(v + (i64)__ldg(cols + t));
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    part += __shfl_xor_sync(0xffffffffu, part, 4);
    if ((t & 7) == 0 && t < nent) {
      const i64 p = t >> 3;
      const double r = scale ? part * __ldg(scale + p) : part;
      y[perm ? (i64)__ldg(perm + p) : p] = r;
    }
  }
}

// dst[i] = src[idx[i]]: brings a result computed in bucket order back to point order.  A random 8-byte READ costs one
// 32-byte sector; the equivalent scatter (dst[perm[p]] = ..) costs a sector fill plus a sector write-back per value.
// 1e8 values take 2.2 ms = 46 G random sectors / s, the rate of the memory system whatever the unrolling; routing the
// values through destination blocks in two passes (scattered writes merged in L2, then reads inside an L2-resident
// window) measures 3.8 ms or more (tools/perm_probe).
__global__ void __launch_bounds__(256) k_gather_rows(i64 n, const double *__restrict__ src, const int *__restrict__ idx,
                                                     double *__restrict__ dst) {
  const i64 stride = (i64)gridDim.x * blockDim.x;
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    const int a = __ldg(idx + i), b = __ldg(idx + i + stride), c = __ldg(idx + i + 2 * stride), d = __ldg(idx + i + 3 * stride);
    const double va = __ldg(src + a), vb = __ldg(src + b), vc = __ldg(src + c), vd = __ldg(src + d);
    dst[i] = va;
    dst[i + stride] = vb;
    dst[i + 2 * stride] = vc;
    dst[i + 3 * stride] = vd;
  }
  for (; i < n; i += stride) dst[i] = __ldg(src + __ldg(idx + i));
}

// y[cols[p,e]] += scale[p] w[p,e] x[p]
template <typename IDX>
__global__ void __launch_bounds__(256) k_interp_apply_t(const IDX *__restrict__ cols, const double *__restrict__ w,
                                                        const double *__restrict__ scale, const int *__restrict__ perm,
                                                        i64 npts, const double *__restrict__ x, double *__restrict__ y) {
  const i64 nent = npts * 8;
  const i64 stride = (i64)gridDim.x * blockDim.x;
  for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < nent; t += stride) {
    const i64 p = t >> 3;
    double xv = __ldg(x + (perm ? (i64)__ldg(perm + p) : p));
    if (scale) xv *= __ldg(scale + p);
    atomicAdd(y + (i64)__ldg(cols + t), __ldg(w + t) * xv);
  }
}

// fused locate + apply: nothing but the points is read and y written (plus the gathers)
__global__ void __launch_bounds__(256) k_interp_fused(Axes g, const double *__restrict__ pts, i64 npts, i64 col_offset,
                                                      const double *__restrict__ scale, const double *__restrict__ v,
                                                      double *__restrict__ y) {
  extern __shared__ double smem[];
  const Axes a = stage_axes(g, smem, 4096);
  const i64 stride = (i64)gridDim.x * blockDim.x;
  for (i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x; p < npts; p += stride) {
    const AxisHit hx = locate_axis(a.tx, a.ntx, __ldg(pts + 3 * p), a.ix);
    const AxisHit hy = locate_axis(a.ty, a.nty, __ldg(pts + 3 * p + 1), a.iy);
    const AxisHit hz = locate_axis(a.tz, a.ntz, __ldg(pts + 3 * p + 2), a.iz);
    i64 c8[8];
    double w8[8];
    entries(hx, hy, hz, a.ntx, a.nty, col_offset, c8, w8);
    double acc = 0.0;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc += w8[e] * __ldg(v + c8[e]);
    y[p] = scale ? acc * __ldg(scale + p) : acc;
  }
}


// ---- compact tables (K10c / K11c) ---------------------------------------------------------------------------------
// The eight entries of a row are a product structure: columns base + {0,1} + {0,sy} + {0,sz}, weights
// ((wx * wy) * wz) with w2 = 1 - w1 per axis.  A compact row keeps the flat index of the first grid sample and the three
// w1 (28 bytes instead of 96); the kernels rebuild columns and weights with the reference's expressions, so the expanded
// tables are bit-identical to K10's.  An unclamped w1 = (a[i2] - x) / (a[i2] - a[i1]) with a[i1] <= x < a[i2] lies in
// (0, 1]; where the axis is clamped (i2 == i1, both weights 0.5 in the reference) the weight is stored as -0.5, the
// sign being the "no step along this axis" flag.
// One THREAD per point: its eight gathers are in flight together (the 8-lanes-per-point kernel above keeps only four
// points per warp in flight and is latency-bound: ncu long_scoreboard 76 %, DRAM 38 % busy).
__device__ __forceinline__ double compact_weight(const AxisHit &h) { return h.i1 == h.i2 ? -0.5 : h.w1; }

template <typename IDX>
__global__ void __launch_bounds__(256) k_interp_locate_c(Axes g, const double *__restrict__ pts, i64 npts, i64 col_offset,
                                                         IDX *__restrict__ base, double *__restrict__ wx,
                                                         double *__restrict__ wy, double *__restrict__ wz) {
  Axes a = g;
  axes_means(a);
  const i64 stride = (i64)gridDim.x * blockDim.x;
  for (i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x; p < npts; p += stride) {
    const AxisHit hx = locate_axis(a.tx, a.ntx, __ldg(pts + 3 * p), a.ix);
    const AxisHit hy = locate_axis(a.ty, a.nty, __ldg(pts + 3 * p + 1), a.iy);
    const AxisHit hz = locate_axis(a.tz, a.ntz, __ldg(pts + 3 * p + 2), a.iz);
    base[p] = (IDX)(col_offset + hx.i1 + a.ntx * (hy.i1 + a.nty * (i64)hz.i1));
    wx[p] = compact_weight(hx);
    wy[p] = compact_weight(hy);
    wz[p] = compact_weight(hz);
  }
}

struct CompactRow {
  i64 col[8];
  double w[8];
};

// columns and weights of a compact row in the entry order of _interpmat3D (see entries())
__device__ __forceinline__ CompactRow expand_row(i64 b, double ax, double ay, double az, i64 sy, i64 sz) {
  const double x1 = fabs(ax), y1 = fabs(ay), z1 = fabs(az);
  const double x2 = 1 - x1, y2 = 1 - y1, z2 = 1 - z1;
  const i64 dx = ax < 0 ? 0 : 1, dy = ay < 0 ? 0 : sy, dz = az < 0 ? 0 : sz;
  CompactRow r;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const bool ux = (e >> 1) & 1, uy = e & 1, uz = (e >> 2) & 1;
    r.col[e] = b + (ux ? dx : 0) + (uy ? dy : 0) + (uz ? dz : 0);
    r.w[e] = ((ux ? x2 : x1) * (uy ? y2 : y1)) * (uz ? z2 : z1);
  }
  return r;
}

template <typename IDX>
__global__ void __launch_bounds__(256) k_interp_apply_c(const IDX *__restrict__ base, const double *__restrict__ wx,
                                                        const double *__restrict__ wy, const double *__restrict__ wz,
                                                        i64 sy, i64 sz, const double *__restrict__ scale, i64 npts,
                                                        const double *__restrict__ v, double *__restrict__ y) {
  const i64 stride = (i64)gridDim.x * blockDim.x;
  for (i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x; p < npts; p += stride) {
    const CompactRow r = expand_row((i64)__ldg(base + p), __ldg(wx + p), __ldg(wy + p), __ldg(wz + p), sy, sz);
    double xv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) xv[e] = __ldg(v + r.col[e]);
    double acc = 0.0;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc += r.w[e] * xv[e];
    __stcs(y + p, scale ? acc * __ldg(scale + p) : acc);
  }
}

// y[col[p,e]] += scale[p] w[p,e] x[perm ? perm[p] : p]
template <typename IDX>
__global__ void __launch_bounds__(256) k_interp_apply_t_c(const IDX *__restrict__ base, const double *__restrict__ wx,
                                                          const double *__restrict__ wy, const double *__restrict__ wz,
                                                          i64 sy, i64 sz, const double *__restrict__ scale,
                                                          const int *__restrict__ perm, i64 npts,
                                                          const double *__restrict__ x, double *__restrict__ y) {
  const i64 stride = (i64)gridDim.x * blockDim.x;
  for (i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x; p < npts; p += stride) {
    const CompactRow r = expand_row((i64)__ldg(base + p), __ldg(wx + p), __ldg(wy + p), __ldg(wz + p), sy, sz);
    double xv = __ldg(x + (perm ? (i64)__ldg(perm + p) : p));
    if (scale) xv *= __ldg(scale + p);
#pragma unroll
    for (int e = 0; e < 8; ++e) atomicAdd(y + r.col[e], r.w[e] * xv);
  }
}

// the (npts, 8) tables of K10 from compact rows (for .tocsr() and for callers that want scipy-like arrays)
template <typename IDX, typename CIDX>
__global__ void __launch_bounds__(256) k_interp_expand(const IDX *__restrict__ base, const double *__restrict__ wx,
                                                       const double *__restrict__ wy, const double *__restrict__ wz,
                                                       i64 sy, i64 sz, i64 npts, CIDX *__restrict__ cols,
                                                       double *__restrict__ w) {
  const i64 stride = (i64)gridDim.x * blockDim.x;
  for (i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x; p < npts; p += stride) {
    const CompactRow r = expand_row((i64)__ldg(base + p), __ldg(wx + p), __ldg(wy + p), __ldg(wz + p), sy, sz);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      cols[p * 8 + e] = (CIDX)r.col[e];
      w[p * 8 + e] = r.w[e];
    }
  }
}

struct Box {
  double lo[3], hi[3], tol[3];
};

// is_inside (base_tensor_mesh.py:638-682)
__global__ void k_points_inside(const double *__restrict__ pts, i64 npts, Box b, uint8_t *__restrict__ inside) {
  const i64 stride = (i64)gridDim.x * blockDim.x;
  for (i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x; p < npts; p += stride) {
    bool in = true;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const double x = __ldg(pts + 3 * p + a);
      in = in && (x >= b.lo[a] - b.tol[a]) && (x <= b.hi[a] + b.tol[a]);
    }
    inside[p] = in ? 1 : 0;
  }
}

static int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static unsigned blocks_for(i64 work_items, int per_block, int waves) {
  const i64 need = (work_items + per_block - 1) / per_block;
  const i64 cap = (i64)sm_count() * waves;
  return (unsigned)max((i64)1, min(need, cap));
}

}  // namespace dzb

using namespace dzb;

template <typename IDX>
static void launch_locate(const Axes &g, const double *pts, i64 npts, i64 col_offset, void *cols, double *w, cudaStream_t s) {
  const size_t smem = IT * 9 * (sizeof(double) + sizeof(IDX));
  k_interp_locate<IDX><<<blocks_for(npts, IT, 16), IT, smem, s>>>(g, pts, npts, col_offset, (IDX *)cols, w);
}

extern "C" int dzb_interp_locate(const double *tx, int64_t ntx, const double *ty, int64_t nty, const double *tz,
                                 int64_t ntz, const double *pts, int64_t npts, int64_t col_offset, void *cols,
                                 int cols_is64, double *w, void *stream) {
  if (npts == 0) return 0;
  if (!cols_is64 && col_offset + ntx * nty * ntz > 2147483647LL) {
    set_error("dzb_interp_locate: column indices do not fit int32; pass an int64 table (cols_is64 = 1)");
    return -1;
  }
  Axes g{tx, ty, tz, ntx, nty, ntz, 0.0, 0.0, 0.0};
  if (cols_is64) launch_locate<i64>(g, pts, npts, col_offset, cols, w, (cudaStream_t)stream);
  else launch_locate<int>(g, pts, npts, col_offset, cols, w, (cudaStream_t)stream);
  return check_launch("dzb_interp_locate");
}

extern "C" int dzb_interp_apply_perm(const void *cols, int cols_is64, const double *w, const double *row_scale,
                                     const int32_t *perm, int64_t npts, const double *v, double *y, void *stream) {
  if (npts == 0) return 0;
  if (perm && npts >= ((int64_t)1 << 31)) {
    set_error("dzb_interp_apply_perm: bucketed tables are limited to 2^31 - 1 points");
    return -1;
  }
  const unsigned nb = blocks_for(npts * 8, 256, 16);
  if (cols_is64) k_interp_apply<i64><<<nb, 256, 0, (cudaStream_t)stream>>>((const i64 *)cols, w, row_scale, perm, npts, v, y);
  else k_interp_apply<int><<<nb, 256, 0, (cudaStream_t)stream>>>((const int *)cols, w, row_scale, perm, npts, v, y);
  return check_launch("dzb_interp_apply");
}

extern "C" int dzb_gather_rows(int64_t n, const double *src, const int32_t *idx, double *dst, void *stream) {
  if (n == 0) return 0;
  k_gather_rows<<<blocks_for(n, 256, 32), 256, 0, (cudaStream_t)stream>>>(n, src, idx, dst);
  return check_launch("dzb_gather_rows");
}

extern "C" int dzb_interp_apply(const void *cols, int cols_is64, const double *w, const double *row_scale, int64_t npts,
                                const double *v, double *y, void *stream) {
  return dzb_interp_apply_perm(cols, cols_is64, w, row_scale, nullptr, npts, v, y, stream);
}

extern "C" int dzb_interp_apply_t_perm(const void *cols, int cols_is64, const double *w, const double *row_scale,
                                       const int32_t *perm, int64_t npts, const double *x, double *y, void *stream) {
  if (npts == 0) return 0;
  if (perm && npts >= ((int64_t)1 << 31)) {
    set_error("dzb_interp_apply_t_perm: bucketed tables are limited to 2^31 - 1 points");
    return -1;
  }
  const unsigned nb = blocks_for(npts * 8, 256, 16);
  if (cols_is64) k_interp_apply_t<i64><<<nb, 256, 0, (cudaStream_t)stream>>>((const i64 *)cols, w, row_scale, perm, npts, x, y);
  else k_interp_apply_t<int><<<nb, 256, 0, (cudaStream_t)stream>>>((const int *)cols, w, row_scale, perm, npts, x, y);
  return check_launch("dzb_interp_apply_t");
}

extern "C" int dzb_interp_apply_t(const void *cols, int cols_is64, const double *w, const double *row_scale,
                                  int64_t npts, const double *x, double *y, void *stream) {
  return dzb_interp_apply_t_perm(cols, cols_is64, w, row_scale, nullptr, npts, x, y, stream);
}


// ---- compact tables ----------------------------------------------------------------------------------------------
extern "C" int dzb_interp_locate_compact(const double *tx, int64_t ntx, const double *ty, int64_t nty, const double *tz,
                                         int64_t ntz, const double *pts, int64_t npts, int64_t col_offset, void *base,
                                         int base_is64, double *wx, double *wy, double *wz, void *stream) {
  if (npts == 0) return 0;
  if (!base_is64 && col_offset + ntx * nty * ntz > 2147483647LL) {
    set_error("dzb_interp_locate_compact: column indices do not fit int32; pass an int64 table (base_is64 = 1)");
    return -1;
  }
  Axes g{tx, ty, tz, ntx, nty, ntz, 0.0, 0.0, 0.0};
  const unsigned nb = blocks_for(npts, 256, 16);
  if (base_is64) k_interp_locate_c<i64><<<nb, 256, 0, (cudaStream_t)stream>>>(g, pts, npts, col_offset, (i64 *)base, wx, wy, wz);
  else k_interp_locate_c<int><<<nb, 256, 0, (cudaStream_t)stream>>>(g, pts, npts, col_offset, (int *)base, wx, wy, wz);
  return check_launch("dzb_interp_locate_compact");
}

extern "C" int dzb_interp_apply_compact(const void *base, int base_is64, const double *wx, const double *wy,
                                        const double *wz, int64_t ntx, int64_t nty, const double *row_scale, int64_t npts,
                                        const double *v, double *y, void *stream) {
  if (npts == 0) return 0;
  const unsigned nb = blocks_for(npts, 256, 32);
  const i64 sy = ntx, sz = ntx * nty;
  if (base_is64) k_interp_apply_c<i64><<<nb, 256, 0, (cudaStream_t)stream>>>((const i64 *)base, wx, wy, wz, sy, sz, row_scale, npts, v, y);
  else k_interp_apply_c<int><<<nb, 256, 0, (cudaStream_t)stream>>>((const int *)base, wx, wy, wz, sy, sz, row_scale, npts, v, y);
  return check_launch("dzb_interp_apply_compact");
}

extern "C" int dzb_interp_apply_t_compact(const void *base, int base_is64, const double *wx, const double *wy,
                                          const double *wz, int64_t ntx, int64_t nty, const double *row_scale,
                                          const int32_t *perm, int64_t npts, const double *x, double *y, void *stream) {
  if (npts == 0) return 0;
  if (perm && npts >= ((int64_t)1 << 31)) {
    set_error("dzb_interp_apply_t_compact: bucketed tables are limited to 2^31 - 1 points");
    return -1;
  }
  const unsigned nb = blocks_for(npts, 256, 32);
  const i64 sy = ntx, sz = ntx * nty;
  if (base_is64) k_interp_apply_t_c<i64><<<nb, 256, 0, (cudaStream_t)stream>>>((const i64 *)base, wx, wy, wz, sy, sz, row_scale, perm, npts, x, y);
  else k_interp_apply_t_c<int><<<nb, 256, 0, (cudaStream_t)stream>>>((const int *)base, wx, wy, wz, sy, sz, row_scale, perm, npts, x, y);
  return check_launch("dzb_interp_apply_t_compact");
}

extern "C" int dzb_interp_expand(const void *base, int base_is64, const double *wx, const double *wy, const double *wz,
                                 int64_t ntx, int64_t nty, int64_t npts, void *cols, int cols_is64, double *w,
                                 void *stream) {
  if (npts == 0) return 0;
  if (base_is64 && !cols_is64) {
    set_error("dzb_interp_expand: int64 compact rows need an int64 column table");
    return -1;
  }
  const unsigned nb = blocks_for(npts, 256, 16);
  const i64 sy = ntx, sz = ntx * nty;
  cudaStream_t s = (cudaStream_t)stream;
  if (base_is64) k_interp_expand<i64, i64><<<nb, 256, 0, s>>>((const i64 *)base, wx, wy, wz, sy, sz, npts, (i64 *)cols, w);
  else if (cols_is64) k_interp_expand<int, i64><<<nb, 256, 0, s>>>((const int *)base, wx, wy, wz, sy, sz, npts, (i64 *)cols, w);
  else k_interp_expand<int, int><<<nb, 256, 0, s>>>((const int *)base, wx, wy, wz, sy, sz, npts, (int *)cols, w);
  return check_launch("dzb_interp_expand");
}

extern "C" int dzb_interp_fused(const double *tx, int64_t ntx, const double *ty, int64_t nty, const double *tz,
                                int64_t ntz, const double *pts, int64_t npts, int64_t col_offset,
                                const double *row_scale, const double *v, double *y, void *stream) {
  if (npts == 0) return 0;
  Axes g{tx, ty, tz, ntx, nty, ntz, 0.0, 0.0, 0.0};
  const size_t smem = 4096 * sizeof(double);
  k_interp_fused<<<blocks_for(npts, 256, 8), 256, smem, (cudaStream_t)stream>>>(g, pts, npts, col_offset, row_scale, v, y);
  return check_launch("dzb_interp_fused");
}

extern "C" int dzb_points_inside(const double *pts, int64_t npts, const double *lo_host, const double *hi_host,
                                 const double *tol_host, uint8_t *inside, void *stream) {
  if (npts == 0) return 0;
  Box b;
  for (int a = 0; a < 3; ++a) {
    b.lo[a] = lo_host[a];
    b.hi[a] = hi_host[a];
    b.tol[a] = tol_host[a];
  }
  k_points_inside<<<blocks_for(npts, 256, 16), 256, 0, (cudaStream_t)stream>>>(pts, npts, b, inside);
  return check_launch("dzb_points_inside");
}
