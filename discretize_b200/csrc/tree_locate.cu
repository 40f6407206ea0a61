// N3: the OcTree leaf that contains (or is closest to) each point, on the device.
//
// reference: _extensions/tree.cpp:758-766 (containing_cell), 1499-1516 (find_cell): the root is found with
// "x >= upper edge -> next root" (ties go high); the descent compares the point with the centre of the current cell,
// "x > centre -> high child" (ties go low), until it reaches a leaf; points outside the mesh end in a boundary cell.
//
// The centres the descent compares with are nodes of the FINEST grid, and its decisions are exactly the steps of a binary
// search over the fine nodes strictly inside the root: the fine cell the search ends in is
//     (number of interior fine nodes of the root that are < x)      -- bisect_left, ties go low,
// whatever the level the reference stops at, and the leaf it stops at is the one whose Z-order range holds that fine cell.
// So: two bisections per axis (roots: bisect_right; fine nodes: bisect_left), one Morton key, one bisection over the sorted
// leaf keys (cells are numbered root by root, x fastest, then in Z-order: tree.cpp:232-344, 949-953) -- no tree walk.
#include "tm_common.cuh"

namespace dzb {

struct TreeAxes {
  const double *nodes[3];   // fine-grid node coordinates per axis (n + 1 entries)
  int roots[3];             // root cells per axis
  int rw;                   // root width in fine cells (a power of two)
  int nbits;                // log2(rw)
};

__device__ __forceinline__ i64 spread3(i64 v, int nbits) {
  i64 out = 0;
  for (int b = 0; b < nbits; ++b) out |= ((v >> b) & 1) << (3 * b);
  return out;
}

template <typename OUT>
__global__ void __launch_bounds__(256) k_tree_locate(TreeAxes ax, const i64 *__restrict__ leaf_keys, i64 n_leaves,
                                                     const double *__restrict__ pts, i64 npts, OUT *__restrict__ out) {
  const i64 stride = (i64)gridDim.x * blockDim.x;
  const i64 rw3 = (i64)ax.rw * ax.rw * ax.rw;
  for (i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x; p < npts; p += stride) {
    i64 root_id = 0, rstride = 1, code = 0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const double x = __ldg(pts + 3 * p + d);
      const double *nd = ax.nodes[d];
      // root: number of interior root boundaries <= x (bisect_right)
      int lo = 0, hi = ax.roots[d] - 1;
      while (lo < hi) {
        const int mid = (lo + hi) / 2;
        if (x < __ldg(nd + (i64)ax.rw * (mid + 1))) hi = mid;
        else lo = mid + 1;
      }
      const int root = lo;
      // fine cell inside the root: number of interior fine nodes < x (bisect_left)
      const double *in = nd + (i64)ax.rw * root + 1;
      int a = 0, b = ax.rw - 1;
      while (a < b) {
        const int mid = (a + b) / 2;
        if (__ldg(in + mid) < x) a = mid + 1;
        else b = mid;
      }
      root_id += rstride * root;
      rstride *= ax.roots[d];
      code |= spread3(a, ax.nbits) << d;
    }
    const i64 key = root_id * rw3 + code;
    // last leaf whose key is <= key
    i64 lo = 0, hi = n_leaves;
    while (lo < hi) {
      const i64 mid = (lo + hi) / 2;
      if (__ldg(leaf_keys + mid) <= key) lo = mid + 1;
      else hi = mid;
    }
    out[p] = (OUT)(lo - 1);
  }
}

}  // namespace dzb

using namespace dzb;

extern "C" int dzb_tree_locate(const double *nodes_x, const double *nodes_y, const double *nodes_z, const int32_t *roots,
                               int root_width, const int64_t *leaf_keys, int64_t n_leaves, const double *pts,
                               int64_t npts, int64_t *cells, void *stream) {
  if (npts <= 0) return 0;
  if (root_width <= 0 || (root_width & (root_width - 1)) != 0 || n_leaves <= 0) {
    set_error("dzb_tree_locate: the root width must be a power of two and the tree must have cells");
    return -1;
  }
  TreeAxes ax;
  ax.nodes[0] = nodes_x;
  ax.nodes[1] = nodes_y;
  ax.nodes[2] = nodes_z;
  for (int d = 0; d < 3; ++d) ax.roots[d] = roots[d];
  ax.rw = root_width;
  ax.nbits = 0;
  while ((1 << ax.nbits) < root_width) ++ax.nbits;
  const i64 need = (npts + 255) / 256;
  const unsigned blocks = (unsigned)(need < 148 * 16 ? need : 148 * 16);
  k_tree_locate<i64><<<blocks, 256, 0, (cudaStream_t)stream>>>(ax, (const i64 *)leaf_keys, n_leaves, pts, npts, (i64 *)cells);
  return check_launch("dzb_tree_locate");
}
