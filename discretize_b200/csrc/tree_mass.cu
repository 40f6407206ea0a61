// K13t: full-tensor mass matrices on an OcTree, matrix-free.
//
// Reference: M = sum over the 8 cell corners n of (P_n R)^T sqrt(V/8) Sigma sqrt(V/8) (P_n R)
//   operators/inner_products.py:147-272 (the generic corner sum), _extensions/tree_ext.pyx:5651-5763 (_getFaceP /
//   _getEdgeP: which face / edge of a cell touches corner n), R = the deflation matrix (hanging -> parents).
// So  M x = R^T A (R x)  with  A = sum_n P_n^T W P_n  acting on TOTAL (hanging + non-hanging) entities, and A is a pure
// cell loop: corner n = (a, b, c) of cell C picks  u = (x[fx_a], x[fy_b], x[fz_c])  (faces) or
// u = (x[ex_{b+2c}], x[ey_{a+2c}], x[ez_{a+2b}])  (edges), forms  t = (V_C / 8) Sigma_C u  and adds t back to the same three
// entities.  One thread per cell: 6 (12) gathers, the 8 corners in registers, 6 (12) atomic adds.  R and R^T are applied
// by the CSR kernel around it (connectivity uploaded once).  Nothing is assembled; the model is read once per apply.
//
// The model derivative (operators/inner_products.py:459-565, tensorType 3) is the same loop: forward F(v) dm = the mass
// apply with dm as the model; adjoint (F(v)^T w)_C = (V_C / 8) sum_n [u_n(v) (x) u_n(w)] contracted into the 6 slots
// (xx, yy, zz, xy, xz, yz) -- a cell-wise output, no atomics.
#include "tm_common.cuh"

namespace dzb {

template <bool EDGES> struct CellEnts {
  static constexpr int PER_DIR = EDGES ? 4 : 2, N = 3 * PER_DIR;
  // local entity of direction d touching corner (a, b, c): the reference's corner tables
  __device__ static __forceinline__ int pick(int d, int a, int b, int c) {
    if (!EDGES) return d == 0 ? a : d == 1 ? b : c;
    return d == 0 ? b + 2 * c : d == 1 ? a + 2 * c : a + 2 * b;
  }
};

template <bool EDGES>
__global__ void __launch_bounds__(256) k_tree_mass_tensor(i64 nC, const int *__restrict__ ent, const double *__restrict__ model,
                                                         const double *__restrict__ vol, const double *__restrict__ xt,
                                                         double *__restrict__ yt) {
  typedef CellEnts<EDGES> E;
  const i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nC) return;
  int id[E::N];
  double x[E::N], acc[E::N];
#pragma unroll
  for (int q = 0; q < E::N; ++q) {
    id[q] = __ldg(ent + (i64)q * nC + c);     // entity tables are stored slot-major: coalesced
    x[q] = __ldg(xt + id[q]);
    acc[q] = 0.0;
  }
  const double sxx = __ldg(model + c), syy = __ldg(model + nC + c), szz = __ldg(model + 2 * nC + c);
  const double sxy = __ldg(model + 3 * nC + c), sxz = __ldg(model + 4 * nC + c), syz = __ldg(model + 5 * nC + c);
  const double w = 0.125 * __ldg(vol + c);
#pragma unroll
  for (int cc = 0; cc < 2; ++cc)
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const int ix = E::pick(0, a, b, cc), iy = E::PER_DIR + E::pick(1, a, b, cc), iz = 2 * E::PER_DIR + E::pick(2, a, b, cc);
        const double ux = x[ix], uy = x[iy], uz = x[iz];
        acc[ix] += w * (sxx * ux + sxy * uy + sxz * uz);
        acc[iy] += w * (sxy * ux + syy * uy + syz * uz);
        acc[iz] += w * (sxz * ux + syz * uy + szz * uz);
      }
#pragma unroll
  for (int q = 0; q < E::N; ++q) atomicAdd(yt + id[q], acc[q]);
}

template <bool EDGES>
__global__ void __launch_bounds__(256) k_tree_mass_tensor_adj(i64 nC, const int *__restrict__ ent, const double *__restrict__ vol,
                                                             const double *__restrict__ vt, const double *__restrict__ wt,
                                                             double *__restrict__ out) {
  typedef CellEnts<EDGES> E;
  const i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nC) return;
  double v[E::N], ww[E::N];
#pragma unroll
  for (int q = 0; q < E::N; ++q) {
    const int id = __ldg(ent + (i64)q * nC + c);
    v[q] = __ldg(vt + id);
    ww[q] = __ldg(wt + id);
  }
  double d[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int cc = 0; cc < 2; ++cc)
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const int ix = E::pick(0, a, b, cc), iy = E::PER_DIR + E::pick(1, a, b, cc), iz = 2 * E::PER_DIR + E::pick(2, a, b, cc);
        d[0] += v[ix] * ww[ix];
        d[1] += v[iy] * ww[iy];
        d[2] += v[iz] * ww[iz];
        d[3] += v[ix] * ww[iy] + v[iy] * ww[ix];
        d[4] += v[ix] * ww[iz] + v[iz] * ww[ix];
        d[5] += v[iy] * ww[iz] + v[iz] * ww[iy];
      }
  const double w = 0.125 * __ldg(vol + c);
#pragma unroll
  for (int q = 0; q < 6; ++q) out[(i64)q * nC + c] = w * d[q];
}

}  // namespace dzb

using namespace dzb;

extern "C" int dzb_tree_mass_tensor(int64_t n_cells, int edges, const int32_t *ent, const double *model6,
                                    const double *cell_volumes, const double *x_total, double *y_total, void *stream) {
  if (n_cells <= 0) return 0;
  const unsigned nb = (unsigned)((n_cells + 255) / 256);
  if (edges) k_tree_mass_tensor<true><<<nb, 256, 0, (cudaStream_t)stream>>>(n_cells, ent, model6, cell_volumes, x_total, y_total);
  else k_tree_mass_tensor<false><<<nb, 256, 0, (cudaStream_t)stream>>>(n_cells, ent, model6, cell_volumes, x_total, y_total);
  return check_launch("dzb_tree_mass_tensor");
}

extern "C" int dzb_tree_mass_tensor_adj(int64_t n_cells, int edges, const int32_t *ent, const double *cell_volumes,
                                        const double *v_total, const double *w_total, double *out6, void *stream) {
  if (n_cells <= 0) return 0;
  const unsigned nb = (unsigned)((n_cells + 255) / 256);
  if (edges) k_tree_mass_tensor_adj<true><<<nb, 256, 0, (cudaStream_t)stream>>>(n_cells, ent, cell_volumes, v_total, w_total, out6);
  else k_tree_mass_tensor_adj<false><<<nb, 256, 0, (cudaStream_t)stream>>>(n_cells, ent, cell_volumes, v_total, w_total, out6);
  return check_launch("dzb_tree_mass_tensor_adj");
}
