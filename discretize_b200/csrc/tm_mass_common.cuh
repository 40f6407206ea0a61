// Model access and diagonal mass entries shared by the mass and fused kernels.
#pragma once
#include "tm_rows.cuh"

namespace dzb {

struct Model {
  int kind;          // DzbModelKind
  const double *p;   // device, Fortran-ordered columns
  double scalar;
  i64 nC;
  bool inv;          // invert_model

  // diagonal component c (0,1,2) of the (possibly inverted) property in `cell`; kinds 0..2
  __device__ __forceinline__ double diag(int c, i64 cell) const {
    double v = kind == DZB_MODEL_SCALAR ? scalar : kind == DZB_MODEL_ISO ? __ldg(p + cell) : __ldg(p + c * nC + cell);
    return inv ? 1.0 / v : v;
  }
  // raw (never inverted) value, for derivative chain rules
  __device__ __forceinline__ double raw(int c, i64 cell) const {
    return kind == DZB_MODEL_SCALAR ? scalar : kind == DZB_MODEL_ISO ? __ldg(p + cell) : __ldg(p + c * nC + cell);
  }
  // symmetric 3x3 tensor of `cell` as s[6] = xx,yy,zz,xy,xz,yz (any kind), inverted if asked
  __device__ __forceinline__ void tensor(i64 cell, double (&s)[6]) const {
    if (kind == DZB_MODEL_TENSOR) {
#pragma unroll
      for (int q = 0; q < 6; ++q) s[q] = __ldg(p + q * nC + cell);
      if (inv) {  // closed-form inverse (utils/matrix_utils.py:734-754)
        const double a11 = s[0], a22 = s[1], a33 = s[2], a12 = s[3], a13 = s[4], a23 = s[5];
        const double det = a11 * a22 * a33 + a12 * a23 * a13 + a13 * a12 * a23 - a13 * a22 * a13 - a12 * a12 * a33 -
                           a11 * a23 * a23;
        s[0] = (a22 * a33 - a23 * a23) / det;
        s[3] = -(a12 * a33 - a13 * a23) / det;
        s[4] = (a12 * a23 - a13 * a22) / det;
        s[1] = (a11 * a33 - a13 * a13) / det;
        s[5] = -(a11 * a23 - a13 * a12) / det;
        s[2] = (a11 * a22 - a12 * a12) / det;
      }
    } else {
      s[0] = diag(0, cell); s[1] = diag(1, cell); s[2] = diag(2, cell);
      s[3] = s[4] = s[5] = 0.0;
    }
  }
};

__device__ __forceinline__ double cell_volume(const TM &m, i64 a, i64 b, i64 c) {
  return (__ldg(m.hx + a) * __ldg(m.hy + b)) * __ldg(m.hz + c);  // (hx*hy)*hz like tensor_mesh.py:296-310
}

// location of component COMP of a face / edge vector
template <bool EDGE, int COMP> struct DofLoc {
  static constexpr int LOC = AveSrc<EDGE, COMP>::LOC;
};

// diagonal entry of the mass matrix at dof (i,j,k) of component COMP (before invert_matrix)
template <bool EDGE, int COMP>
__device__ __forceinline__ double mass_diag_entry(const TM &m, const Model &md, i64 i, i64 j, i64 k) {
  typedef AveSrc<EDGE, COMP> S;
  double acc = 0.0;
  term<S::TX, S::TY, S::TZ>(m, i, j, k, 1.0, [&](i64 a, i64 b, i64 c, double w) {
    acc += w * (cell_volume(m, a, b, c) * md.diag(COMP, m.lidx<LOC_C>(a, b, c)));
  });
  return acc;
}

}  // namespace dzb
