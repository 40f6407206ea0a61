// Mass matrices (InnerProducts mixin) generated on the fly from the cell model.
//
//   K5  diagonal mass matrices  (scalar / isotropic / diagonally-anisotropic model)
//       reference: base/base_tensor_mesh.py:792-863 (_fastInnerProduct):
//         iso   : M = diag( 3 * Av^T (V sigma) )      Av = ave{F,E}2CC  (entries 1/6, 1/12)
//         aniso : M = diag( AvV^T (V sigma_cc) )      AvV = ave{F,E}2CCV (entries 1/2, 1/4)
//       i.e. face: 1/2 * sum of V sigma over the (<=2) adjacent cells,
//            edge: 1/4 * sum of V sigma over the (<=4) adjacent cells.
//   K6  full-tensor mass matrices (<= 9 entries per row)
//       reference: operators/inner_products.py:147-272, 671-738, 792-863:
//         M = sum over the 8 cell corners n of P_n^T sqrt(V/8) Sigma sqrt(V/8) P_n.
//   K7  model derivatives  F(v) = d(M(m) v)/dm  applied forward / adjoint
//       reference: base/base_tensor_mesh.py:865-988, inner_products.py:405-565.
#include "tm_mass_common.cuh"
#include "tm_rowkernels.cuh"
#include "tm_mass_march.cuh"
#include "tm_mass_tensor_march.cuh"

namespace dzb {

__device__ __forceinline__ int sym_idx(int a, int b) {  // (a,b) -> position in xx,yy,zz,xy,xz,yz
  return a == b ? a : (a + b + 2);                        // (0,1)->3 (0,2)->4 (1,2)->5
}

// mode 0: out = diag ; mode 1: out = diag * x.  One launch covers the three components (the cell model is
// read once instead of once per component; neighbouring dofs share it through L1).
template <bool EDGE, int COMP, int MODE>
__device__ __forceinline__ void mass_diag_one(const TM &m, const Model &md, bool inv_mat, i64 i, i64 j, i64 k,
                                              const double *__restrict__ x, double *__restrict__ out) {
  constexpr int LOC = DofLoc<EDGE, COMP>::LOC;
  if (i >= m.sx<LOC>() || j >= m.sy<LOC>() || k >= m.sz<LOC>()) return;
  double d = mass_diag_entry<EDGE, COMP>(m, md, i, j, k);
  if (inv_mat) d = 1.0 / d;
  const i64 row = m.idx<LOC>(i, j, k);
  out[row] = MODE == 0 ? d : d * __ldg(x + row);
}

template <bool EDGE, int MODE>
__global__ void __launch_bounds__(BX *BY *BZ)
    k_mass_diag(TM m, Model md, bool inv_mat, const double *__restrict__ x, double *__restrict__ out) {
  const i64 i = (i64)blockIdx.x * BX + threadIdx.x;
  const i64 j = (i64)blockIdx.y * BY + threadIdx.y;
  const i64 k = (i64)blockIdx.z * BZ + threadIdx.z;
  mass_diag_one<EDGE, 0, MODE>(m, md, inv_mat, i, j, k, x, out);
  mass_diag_one<EDGE, 1, MODE>(m, md, inv_mat, i, j, k, x, out);
  mass_diag_one<EDGE, 2, MODE>(m, md, inv_mat, i, j, k, x, out);
}

int g_mass_tensor_march = 1;  // K6: 0 = generic row generator (cross-check), 1 = plane-marching kernels (> 1: sweep)
int g_mass_march = 1;  // 0: generic kernel (kept as cross-check), 1: plane-marching kernel (> 1: sweep configurations)

template <bool EDGE, int MODE>
static void launch_mass_diag(const TM &m, const Model &md, bool inv_mat, const double *x, double *out, cudaStream_t s) {
  if (g_mass_march) {
    launch_mass_diag_march<EDGE, MODE>(m, md, inv_mat, x, out, s);
    return;
  }
  k_mass_diag<EDGE, MODE><<<grid_for<LOC_N>(m), dim3(BX, BY, BZ), 0, s>>>(m, md, inv_mat, x, out);
}

// ---- full tensor rows ------------------------------------------------------------------------
// flat index of the face / edge of direction T at integer position q[3]
template <bool EDGE, int T> __device__ __forceinline__ i64 dof_index(const TM &m, const i64 (&q)[3]) {
  return m.idx<AveSrc<EDGE, T>::LOC>(q[0], q[1], q[2]);
}
template <bool EDGE> __device__ __forceinline__ i64 dof_index_rt(const TM &m, int t, const i64 (&q)[3]) {
  return t == 0 ? dof_index<EDGE, 0>(m, q) : t == 1 ? dof_index<EDGE, 1>(m, q) : dof_index<EDGE, 2>(m, q);
}

template <bool EDGE> struct RowsMassTensor {
  Model md;
  static constexpr int NCOMP = 3;
  static constexpr int MAXNNZ = 9;
  template <int COMP> static constexpr int loc() { return AveSrc<EDGE, COMP>::LOC; }
  template <int COMP> __host__ __device__ static i64 row_off(const TM &m) { return m.off<loc<COMP>()>(); }

  template <int D, class E> __device__ void row(const TM &m, i64 i, i64 j, i64 k, E &&emit) const {
    const i64 n[3] = {m.nx, m.ny, m.nz};
    const i64 p[3] = {i, j, k};
    constexpr int T1 = (D + 1) % 3 < (D + 2) % 3 ? (D + 1) % 3 : (D + 2) % 3;
    constexpr int T2 = (D + 1) % 3 < (D + 2) % 3 ? (D + 2) % 3 : (D + 1) % 3;
    double diag = 0.0;
    if (!EDGE) {
      // face normal to D at node index p[D]; adjacent cells p[D]-1 (its + face) and p[D] (its - face)
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        i64 q[3] = {p[0], p[1], p[2]};
        q[D] = p[D] - s;
        if (q[D] < 0 || q[D] >= n[D]) continue;
        double sg[6];
        md.tensor(m.lidx<LOC_C>(q[0], q[1], q[2]), sg);
        const double v8 = 0.125 * cell_volume(m, q[0], q[1], q[2]);
        diag += 4.0 * v8 * sg[D];
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          i64 r1[3] = {q[0], q[1], q[2]};
          r1[T1] = q[T1] + b;
          emit(dof_index<false, T1>(m, r1), 2.0 * v8 * sg[sym_idx(D, T1)]);
          i64 r2[3] = {q[0], q[1], q[2]};
          r2[T2] = q[T2] + b;
          emit(dof_index<false, T2>(m, r2), 2.0 * v8 * sg[sym_idx(D, T2)]);
        }
      }
    } else {
      // edge along D: cell index p[D] along D, node indices p[T1], p[T2] across
      double acc1[2][2] = {{0, 0}, {0, 0}}, acc2[2][2] = {{0, 0}, {0, 0}};  // [a][s1], [a][s2]
      bool ok1[2] = {false, false}, ok2[2] = {false, false};
#pragma unroll
      for (int s1 = 0; s1 < 2; ++s1) {
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2) {
          i64 q[3];
          q[D] = p[D]; q[T1] = p[T1] - s1; q[T2] = p[T2] - s2;
          if (q[T1] < 0 || q[T1] >= n[T1] || q[T2] < 0 || q[T2] >= n[T2]) continue;
          ok1[s1] = true; ok2[s2] = true;
          double sg[6];
          md.tensor(m.lidx<LOC_C>(q[0], q[1], q[2]), sg);
          const double v8 = 0.125 * cell_volume(m, q[0], q[1], q[2]);
          diag += 2.0 * v8 * sg[D];
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            acc1[a][s1] += v8 * sg[sym_idx(D, T1)];
            acc2[a][s2] += v8 * sg[sym_idx(D, T2)];
          }
        }
      }
#pragma unroll
      for (int a = 0; a < 2; ++a) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          if (ok1[s]) {  // edge along T1: own index p[T1]-s, node index p[D]+a along D, p[T2] along T2
            i64 r[3];
            r[T1] = p[T1] - s; r[D] = p[D] + a; r[T2] = p[T2];
            emit(dof_index<true, T1>(m, r), acc1[a][s]);
          }
          if (ok2[s]) {
            i64 r[3];
            r[T2] = p[T2] - s; r[D] = p[D] + a; r[T1] = p[T1];
            emit(dof_index<true, T2>(m, r), acc2[a][s]);
          }
        }
      }
    }
    emit(m.idx<AveSrc<EDGE, D>::LOC>(i, j, k), diag);
  }
};

// ---- derivatives, fast path (scalar handled by the caller as a broadcast iso model) -----------
// row scaling r[f]: 1, or -MI^2 (invert_matrix), or +MI^2 (both); column scaling c: 1, -1/m^2, +1/m^2
template <bool EDGE, int COMP>
__device__ __forceinline__ double deriv_row_scale(const TM &m, const Model &md, bool inv_mat, i64 i, i64 j, i64 k) {
  if (!inv_mat) return 1.0;
  const double d = mass_diag_entry<EDGE, COMP>(m, md, i, j, k);
  const double mi = 1.0 / d;
  return md.inv ? mi * mi : -(mi * mi);
}
__device__ __forceinline__ double deriv_col_scale(const Model &md, bool inv_mat, int comp, i64 cell) {
  if (!md.inv) return 1.0;
  const double v = md.raw(comp, cell);
  return inv_mat ? 1.0 / (v * v) : -1.0 / (v * v);
}

// forward: y[f] = v[f] r[f] sum_cells w V c[cell] dm[block(comp) + cell]
template <bool EDGE, int COMP>
__global__ void __launch_bounds__(BX *BY *BZ) k_deriv_fwd(TM m, Model md, bool inv_mat, const double *__restrict__ v,
                                                           const double *__restrict__ dm, double *__restrict__ y) {
  constexpr int LOC = DofLoc<EDGE, COMP>::LOC;
  typedef AveSrc<EDGE, COMP> S;
  const i64 i = (i64)blockIdx.x * BX + threadIdx.x;
  const i64 j = (i64)blockIdx.y * BY + threadIdx.y;
  const i64 k = (i64)blockIdx.z * BZ + threadIdx.z;
  if (i >= m.sx<LOC>() || j >= m.sy<LOC>() || k >= m.sz<LOC>()) return;
  const i64 blk = md.kind == DZB_MODEL_ANISO ? COMP * md.nC : 0;
  double acc = 0.0;
  term<S::TX, S::TY, S::TZ>(m, i, j, k, 1.0, [&](i64 a, i64 b, i64 c, double w) {
    const i64 cell = m.lidx<LOC_C>(a, b, c);
    acc += w * cell_volume(m, a, b, c) * deriv_col_scale(md, inv_mat, COMP, cell) * __ldg(dm + blk + cell);
  });
  const i64 row = m.idx<LOC>(i, j, k);
  y[row] = __ldg(v + row) * deriv_row_scale<EDGE, COMP>(m, md, inv_mat, i, j, k) * acc;
}

// adjoint: y[block(comp) + cell] (+)= c[cell] V sum_{dofs of comp around cell} w r[f] v[f] x[f]
template <bool EDGE, int COMP, bool ACCUM>
__global__ void __launch_bounds__(BX *BY *BZ) k_deriv_adj(TM m, Model md, bool inv_mat, const double *__restrict__ v,
                                                           const double *__restrict__ x, double *__restrict__ y) {
  typedef AveSrc<EDGE, COMP> S;
  constexpr int LOC = S::LOC;
  const i64 i = (i64)blockIdx.x * BX + threadIdx.x;
  const i64 j = (i64)blockIdx.y * BY + threadIdx.y;
  const i64 k = (i64)blockIdx.z * BZ + threadIdx.z;
  if (i >= m.nx || j >= m.ny || k >= m.nz) return;
  double acc = 0.0;
  term<S::OX, S::OY, S::OZ>(m, i, j, k, 1.0, [&](i64 a, i64 b, i64 c, double w) {
    const i64 f = m.idx<LOC>(a, b, c);
    acc += w * deriv_row_scale<EDGE, COMP>(m, md, inv_mat, a, b, c) * __ldg(v + f) * __ldg(x + f);
  });
  const i64 cell = m.lidx<LOC_C>(i, j, k);
  const i64 blk = md.kind == DZB_MODEL_ANISO ? COMP * md.nC : 0;
  const double r = deriv_col_scale(md, inv_mat, COMP, cell) * cell_volume(m, i, j, k) * acc;
  if (ACCUM) y[blk + cell] += r;
  else y[blk + cell] = r;
}

// ---- derivatives, full tensor (inner_products.py:520-563): n x 6nC ------------------------------
// member dof of direction T at corner (c0,c1,c2) of cell q
template <bool EDGE> __device__ __forceinline__ i64 corner_member(const TM &m, int t, const i64 (&q)[3], const int (&cb)[3]) {
  i64 r[3];
  if (!EDGE) {
    r[0] = q[0]; r[1] = q[1]; r[2] = q[2];
    r[t] = q[t] + cb[t];
  } else {
    r[0] = q[0] + cb[0]; r[1] = q[1] + cb[1]; r[2] = q[2] + cb[2];
    r[t] = q[t];
  }
  return dof_index_rt<EDGE>(m, t, r);
}

// forward, one thread per dof of component D:
//   y[f] = sum_{cells, corners touching f} V/8 * sum_b v[member_b] * dm[block(D,b) + cell]
template <bool EDGE, int D>
__global__ void __launch_bounds__(BX *BY *BZ)
    k_deriv_tensor_fwd(TM m, i64 nC, const double *__restrict__ v, const double *__restrict__ dm, double *__restrict__ y) {
  constexpr int LOC = DofLoc<EDGE, D>::LOC;
  const i64 i = (i64)blockIdx.x * BX + threadIdx.x;
  const i64 j = (i64)blockIdx.y * BY + threadIdx.y;
  const i64 k = (i64)blockIdx.z * BZ + threadIdx.z;
  if (i >= m.sx<LOC>() || j >= m.sy<LOC>() || k >= m.sz<LOC>()) return;
  const i64 n[3] = {m.nx, m.ny, m.nz};
  const i64 p[3] = {i, j, k};
  double acc = 0.0;
  // enumerate (cell, corner) pairs whose member of direction D is this dof
  for (int c2 = 0; c2 < 2; ++c2)
    for (int c1 = 0; c1 < 2; ++c1)
      for (int c0 = 0; c0 < 2; ++c0) {
        const int cb[3] = {c0, c1, c2};
        i64 q[3];
        bool ok = true;
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          // faces: member D sits at q[D]+cb[D], other coordinates are the cell's
          // edges: member D sits at cell index along D, corner node coords across
          if (!EDGE) q[t] = (t == D) ? p[t] - cb[t] : p[t];
          else q[t] = (t == D) ? p[t] : p[t] - cb[t];
          ok = ok && q[t] >= 0 && q[t] < n[t];
        }
        if (!ok) continue;
        const i64 cell = m.lidx<LOC_C>(q[0], q[1], q[2]);
        const double v8 = 0.125 * cell_volume(m, q[0], q[1], q[2]);
        double s = 0.0;
#pragma unroll
        for (int b = 0; b < 3; ++b) {
          s += __ldg(v + corner_member<EDGE>(m, b, q, cb)) * __ldg(dm + sym_idx(D, b) * nC + cell);
        }
        acc += v8 * s;
      }
  y[m.idx<LOC>(i, j, k)] = acc;
}

// adjoint, one thread per cell: y[blk*nC + cell] = sum_corners V/8 * sum_{(a,b) in blk} v[member_b] w[member_a]
template <bool EDGE>
__global__ void __launch_bounds__(BX *BY *BZ)
    k_deriv_tensor_adj(TM m, i64 nC, const double *__restrict__ v, const double *__restrict__ x, double *__restrict__ y) {
  const i64 i = (i64)blockIdx.x * BX + threadIdx.x;
  const i64 j = (i64)blockIdx.y * BY + threadIdx.y;
  const i64 k = (i64)blockIdx.z * BZ + threadIdx.z;
  if (i >= m.nx || j >= m.ny || k >= m.nz) return;
  const i64 q[3] = {i, j, k};
  double acc[6] = {0, 0, 0, 0, 0, 0};
  for (int c2 = 0; c2 < 2; ++c2)
    for (int c1 = 0; c1 < 2; ++c1)
      for (int c0 = 0; c0 < 2; ++c0) {
        const int cb[3] = {c0, c1, c2};
        double vv[3], xx[3];
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          const i64 f = corner_member<EDGE>(m, t, q, cb);
          vv[t] = __ldg(v + f);
          xx[t] = __ldg(x + f);
        }
        acc[0] += vv[0] * xx[0];
        acc[1] += vv[1] * xx[1];
        acc[2] += vv[2] * xx[2];
        acc[3] += vv[1] * xx[0] + vv[0] * xx[1];
        acc[4] += vv[2] * xx[0] + vv[0] * xx[2];
        acc[5] += vv[2] * xx[1] + vv[1] * xx[2];
      }
  const double v8 = 0.125 * cell_volume(m, i, j, k);
  const i64 cell = m.lidx<LOC_C>(i, j, k);
#pragma unroll
  for (int b = 0; b < 6; ++b) y[b * nC + cell] = v8 * acc[b];
}

static int make_model(const TM &m, int flags, int kind, const double *model, double scalar, Model *out) {
  if (kind < 0 || kind > 3) {
    set_error("unknown model kind");
    return -1;
  }
  if (kind != DZB_MODEL_SCALAR && model == nullptr) {
    set_error("model pointer is null");
    return -1;
  }
  out->kind = kind;
  out->p = model;
  out->scalar = scalar;
  out->nC = m.nC();
  out->inv = (flags & DZB_INVERT_MODEL) != 0;
  return 0;
}

}  // namespace dzb

using namespace dzb;

// experiment knob (not part of the stable ABI)
extern "C" void dzb_tune_mass(int march) { dzb::g_mass_march = march; }
extern "C" void dzb_tune_mass_tensor(int march) { dzb::g_mass_tensor_march = march; }

// out = the inverses of the nC symmetric 3x3 tensors in `model` (six columns xx, yy, zz, xy, xz, yz of nC entries, like the
// reference's (nC, 6) property in Fortran order; utils/matrix_utils.py:734-754 closed form)
extern "C" int dzb_tm_tensor_invert(int64_t n_cells, const double *model, double *out, void *stream) {
  if (n_cells <= 0) return 0;
  const i64 need = (n_cells + 255) / 256;
  k_tensor_invert<<<(unsigned)(need < 148 * 16 ? need : 148 * 16), 256, 0, (cudaStream_t)stream>>>(n_cells, model, out);
  return check_launch("dzb_tm_tensor_invert");
}

extern "C" int dzb_tm_mass_diag(const DzbTensorMesh *mm, int flags, int model_kind, const double *model,
                                double model_scalar, double *diag, void *stream) {
  if (!valid_mesh(mm)) return -1;
  TM m(*mm);
  Model md;
  if (make_model(m, flags, model_kind, model, model_scalar, &md)) return -1;
  if (model_kind == DZB_MODEL_TENSOR) {
    set_error("dzb_tm_mass_diag: full-tensor models have no diagonal mass matrix");
    return -1;
  }
  const bool inv_mat = (flags & DZB_INVERT_MATRIX) != 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (flags & DZB_EDGES) launch_mass_diag<true, 0>(m, md, inv_mat, nullptr, diag, s);
  else launch_mass_diag<false, 0>(m, md, inv_mat, nullptr, diag, s);
  return check_launch("dzb_tm_mass_diag");
}

extern "C" int dzb_tm_mass_apply(const DzbTensorMesh *mm, int flags, int model_kind, const double *model,
                                 double model_scalar, const double *x, double *y, void *stream) {
  if (!valid_mesh(mm)) return -1;
  TM m(*mm);
  Model md;
  if (make_model(m, flags, model_kind, model, model_scalar, &md)) return -1;
  const bool inv_mat = (flags & DZB_INVERT_MATRIX) != 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (model_kind == DZB_MODEL_TENSOR) {
    if (inv_mat) {
      set_error("Solver needed to invert A.");
      return -2;
    }
    if (g_mass_tensor_march && !md.inv && m.nN() < ((i64)1 << 62)) {
      // (invert_model: the host side inverts the tensors once with dzb_tm_tensor_invert and calls this without the flag)
      if (flags & DZB_EDGES) launch_mass_tensor_march<true>(m, md.p, x, y, s);
      else launch_mass_tensor_march<false>(m, md.p, x, y, s);
      return check_launch("dzb_tm_mass_apply");
    }
    Args a{x, y, nullptr, nullptr, nullptr, nullptr};
    if (flags & DZB_EDGES) launch_all(RowsMassTensor<true>{md}, MODE_APPLY, m, a, s);
    else launch_all(RowsMassTensor<false>{md}, MODE_APPLY, m, a, s);
  } else {
    if (flags & DZB_EDGES) launch_mass_diag<true, 1>(m, md, inv_mat, x, y, s);
    else launch_mass_diag<false, 1>(m, md, inv_mat, x, y, s);
  }
  return check_launch("dzb_tm_mass_apply");
}

extern "C" int dzb_tm_mass_tensor_csr_count(const DzbTensorMesh *mm, int flags, int64_t *row_nnz, void *stream) {
  if (!valid_mesh(mm)) return -1;
  TM m(*mm);
  Model md{DZB_MODEL_SCALAR, nullptr, 1.0, m.nC(), false};  // the pattern does not depend on the values
  Args a{nullptr, nullptr, (i64 *)row_nnz, nullptr, nullptr, nullptr};
  cudaStream_t s = (cudaStream_t)stream;
  if (flags & DZB_EDGES) launch_all(RowsMassTensor<true>{md}, MODE_COUNT, m, a, s);
  else launch_all(RowsMassTensor<false>{md}, MODE_COUNT, m, a, s);
  return check_launch("dzb_tm_mass_tensor_csr_count");
}

extern "C" int dzb_tm_mass_tensor_csr_fill(const DzbTensorMesh *mm, int flags, const double *model,
                                           const int64_t *indptr, int64_t *indices, double *data, void *stream) {
  if (!valid_mesh(mm)) return -1;
  TM m(*mm);
  Model md;
  if (make_model(m, flags, DZB_MODEL_TENSOR, model, 0.0, &md)) return -1;
  Args a{nullptr, nullptr, nullptr, (const i64 *)indptr, (i64 *)indices, data};
  cudaStream_t s = (cudaStream_t)stream;
  if (flags & DZB_EDGES) launch_all(RowsMassTensor<true>{md}, MODE_FILL, m, a, s);
  else launch_all(RowsMassTensor<false>{md}, MODE_FILL, m, a, s);
  return check_launch("dzb_tm_mass_tensor_csr_fill");
}

template <bool EDGE>
static void launch_deriv(const TM &m, const Model &md, bool inv_mat, const double *v, int transpose, const double *x,
                         double *y, cudaStream_t s) {
  dim3 b(BX, BY, BZ);
  if (md.kind == DZB_MODEL_TENSOR) {
    if (!transpose) {
      k_deriv_tensor_fwd<EDGE, 0><<<grid_for<DofLoc<EDGE, 0>::LOC>(m), b, 0, s>>>(m, md.nC, v, x, y);
      k_deriv_tensor_fwd<EDGE, 1><<<grid_for<DofLoc<EDGE, 1>::LOC>(m), b, 0, s>>>(m, md.nC, v, x, y);
      k_deriv_tensor_fwd<EDGE, 2><<<grid_for<DofLoc<EDGE, 2>::LOC>(m), b, 0, s>>>(m, md.nC, v, x, y);
    } else {
      k_deriv_tensor_adj<EDGE><<<grid_for<LOC_C>(m), b, 0, s>>>(m, md.nC, v, x, y);
    }
    return;
  }
  if (g_mass_march && !inv_mat && !md.inv) {
    // no inversions: F(v) dm = v o diag(M(dm)) -- the diagonal mass kernel with dm as the model; and its adjoint
    if (!transpose) {
      Model dmodel = md;
      dmodel.p = x;
      launch_mass_diag_march<EDGE, 1>(m, dmodel, false, v, y, s);
    } else {
      launch_deriv_adj_march<EDGE>(m, md, v, x, y, s);
    }
    return;
  }
  if (!transpose) {
    k_deriv_fwd<EDGE, 0><<<grid_for<DofLoc<EDGE, 0>::LOC>(m), b, 0, s>>>(m, md, inv_mat, v, x, y);
    k_deriv_fwd<EDGE, 1><<<grid_for<DofLoc<EDGE, 1>::LOC>(m), b, 0, s>>>(m, md, inv_mat, v, x, y);
    k_deriv_fwd<EDGE, 2><<<grid_for<DofLoc<EDGE, 2>::LOC>(m), b, 0, s>>>(m, md, inv_mat, v, x, y);
  } else {
    dim3 g = grid_for<LOC_C>(m);
    if (md.kind == DZB_MODEL_ANISO) {
      k_deriv_adj<EDGE, 0, false><<<g, b, 0, s>>>(m, md, inv_mat, v, x, y);
      k_deriv_adj<EDGE, 1, false><<<g, b, 0, s>>>(m, md, inv_mat, v, x, y);
      k_deriv_adj<EDGE, 2, false><<<g, b, 0, s>>>(m, md, inv_mat, v, x, y);
    } else {
      k_deriv_adj<EDGE, 0, false><<<g, b, 0, s>>>(m, md, inv_mat, v, x, y);
      k_deriv_adj<EDGE, 1, true><<<g, b, 0, s>>>(m, md, inv_mat, v, x, y);
      k_deriv_adj<EDGE, 2, true><<<g, b, 0, s>>>(m, md, inv_mat, v, x, y);
    }
  }
}

extern "C" int dzb_tm_mass_deriv_apply(const DzbTensorMesh *mm, int flags, int model_kind, const double *model,
                                       double model_scalar, const double *v, int transpose, const double *x, double *y,
                                       void *stream) {
  if (!valid_mesh(mm)) return -1;
  TM m(*mm);
  Model md;
  if (make_model(m, flags, model_kind, model, model_scalar, &md)) return -1;
  if (model_kind == DZB_MODEL_SCALAR) {
    set_error("dzb_tm_mass_deriv_apply: pass scalar models as a broadcast isotropic model");
    return -1;
  }
  const bool inv_mat = (flags & DZB_INVERT_MATRIX) != 0;
  if (model_kind == DZB_MODEL_TENSOR && (inv_mat || md.inv)) {
    set_error("inverting the property or the matrix is not yet implemented for this mesh/tensorType.");
    return -3;
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (flags & DZB_EDGES) launch_deriv<true>(m, md, inv_mat, v, transpose, x, y, s);
  else launch_deriv<false>(m, md, inv_mat, v, transpose, x, y, s);
  return check_launch("dzb_tm_mass_deriv_apply");
}
