// Row generators for the elementary TensorMesh operators.
//
// Every operator of the reference's DiffOperators mixin on a tensor mesh is a sum
// of tensor products of 1-D two-point stencils (the reference literally builds
// them as kron3(...) of ddx / av matrices: utils/matrix_utils.py:163-253,
// operators/differential_operators.py:161-223, 593-656, 2091-2167, 2478-3253).
// Here a row of such an operator is *generated* on the fly: `term<SRC, OX, OY, OZ>`
// enumerates the (<= 8) entries one tensor-product term contributes to the row at
// (i,j,k) and hands (ix, iy, iz, value) of the source location to a callable.
// The same generators drive y = A x, y = A^T x and the CSR materialiser, so the
// matrix-free apply and `.tocsr()` cannot drift apart.
#pragma once
#include "tm_common.cuh"

namespace dzb {

// 1-D two-point stencils between a node-extent axis (n+1) and a cell-extent axis (n)
enum Ax : int {
  AX_ID = 0,  // same extent on both sides
  AX_D = 1,   // cell i  <- nodes i, i+1 : (-1/h_i, +1/h_i)      (ddx)
  AX_DT = 2,  // node i  <- cells i-1, i : (+1/h_{i-1}, -1/h_i)   (ddx^T)
  AX_A = 3,   // cell i  <- nodes i, i+1 : (1/2, 1/2)             (av)
  AX_AT = 4   // node i  <- cells i-1, i : (1/2, 1/2)             (av^T)
};

template <int OP>
__device__ __forceinline__ int ax_entries(i64 i, i64 n, const double *__restrict__ rh, i64 (&ix)[2], double (&w)[2]) {
  if (OP == AX_ID) {
    ix[0] = i; w[0] = 1.0;
    return 1;
  } else if (OP == AX_D) {
    double r = __ldg(rh + i);
    ix[0] = i; w[0] = -r;
    ix[1] = i + 1; w[1] = r;
    return 2;
  } else if (OP == AX_A) {
    ix[0] = i; w[0] = 0.5;
    ix[1] = i + 1; w[1] = 0.5;
    return 2;
  } else if (OP == AX_DT) {
    int c = 0;
    if (i >= 1) { ix[c] = i - 1; w[c] = __ldg(rh + i - 1); ++c; }
    if (i < n) { ix[c] = i; w[c] = -__ldg(rh + i); ++c; }
    return c;
  } else {  // AX_AT
    int c = 0;
    if (i >= 1) { ix[c] = i - 1; w[c] = 0.5; ++c; }
    if (i < n) { ix[c] = i; w[c] = 0.5; ++c; }
    return c;
  }
}

// Entries of one tensor-product term; emit(ix, iy, iz, value) in source-location indices.
template <int OX, int OY, int OZ, class E>
__device__ __forceinline__ void term(const TM &m, i64 i, i64 j, i64 k, double scale, E &&emit) {
  i64 xs[2], ys[2], zs[2];
  double wx[2], wy[2], wz[2];
  const int cx = ax_entries<OX>(i, m.nx, m.rhx, xs, wx);
  const int cy = ax_entries<OY>(j, m.ny, m.rhy, ys, wy);
  const int cz = ax_entries<OZ>(k, m.nz, m.rhz, zs, wz);
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    if (c >= cz) break;
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      if (b >= cy) break;
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        if (a >= cx) break;
        emit(xs[a], ys[b], zs[c], ((scale * wx[a]) * wy[b]) * wz[c]);
      }
    }
  }
}

#define DZB_THIRD (1.0 / 3.0)

// Rows<OP, TR>: NCOMP row blocks; block c lives on location loc(c) with rows starting at
// row_off(c).  row<COMP>(m,i,j,k,emit) calls emit(flat_column, value).
template <int OP, bool TR> struct Rows;

#define DZB_EMIT(SRC, COLOFF) [&](i64 a, i64 b, i64 c, double v) { emit((COLOFF) + m.lidx<SRC>(a, b, c), v); }

// ---- face_divergence ---------------------------------------------------------------
template <> struct Rows<DZB_OP_DIV, false> {
  static constexpr int NCOMP = 1;
  static constexpr int MAXNNZ = 6;
  template <int COMP> static constexpr int loc() { return LOC_C; }
  template <int COMP> __host__ __device__ static i64 row_off(const TM &m) { return 0; }
  __host__ __device__ static i64 nrows(const TM &m) { return m.nC(); }
  __host__ __device__ static i64 ncols(const TM &m) { return m.nF(); }
  template <int COMP, class E> __device__ static void row(const TM &m, i64 i, i64 j, i64 k, E &&emit) {
    term<AX_D, AX_ID, AX_ID>(m, i, j, k, 1.0, DZB_EMIT(LOC_FX, m.off<LOC_FX>()));
    term<AX_ID, AX_D, AX_ID>(m, i, j, k, 1.0, DZB_EMIT(LOC_FY, m.off<LOC_FY>()));
    term<AX_ID, AX_ID, AX_D>(m, i, j, k, 1.0, DZB_EMIT(LOC_FZ, m.off<LOC_FZ>()));
  }
};
template <> struct Rows<DZB_OP_DIV, true> {
  static constexpr int NCOMP = 3;
  static constexpr int MAXNNZ = 2;
  template <int COMP> static constexpr int loc() { return COMP == 0 ? LOC_FX : COMP == 1 ? LOC_FY : LOC_FZ; }
  template <int COMP> __host__ __device__ static i64 row_off(const TM &m) { return m.off<loc<COMP>()>(); }
  __host__ __device__ static i64 nrows(const TM &m) { return m.nF(); }
  __host__ __device__ static i64 ncols(const TM &m) { return m.nC(); }
  template <int COMP, class E> __device__ static void row(const TM &m, i64 i, i64 j, i64 k, E &&emit) {
    if (COMP == 0) term<AX_DT, AX_ID, AX_ID>(m, i, j, k, 1.0, DZB_EMIT(LOC_C, 0));
    if (COMP == 1) term<AX_ID, AX_DT, AX_ID>(m, i, j, k, 1.0, DZB_EMIT(LOC_C, 0));
    if (COMP == 2) term<AX_ID, AX_ID, AX_DT>(m, i, j, k, 1.0, DZB_EMIT(LOC_C, 0));
  }
};

// ---- edge_curl -----------------------------------------------------------------------
template <> struct Rows<DZB_OP_CURL, false> {
  static constexpr int NCOMP = 3;
  static constexpr int MAXNNZ = 4;
  template <int COMP> static constexpr int loc() { return COMP == 0 ? LOC_FX : COMP == 1 ? LOC_FY : LOC_FZ; }
  template <int COMP> __host__ __device__ static i64 row_off(const TM &m) { return m.off<loc<COMP>()>(); }
  __host__ __device__ static i64 nrows(const TM &m) { return m.nF(); }
  __host__ __device__ static i64 ncols(const TM &m) { return m.nE(); }
  template <int COMP, class E> __device__ static void row(const TM &m, i64 i, i64 j, i64 k, E &&emit) {
    if (COMP == 0) {  // d_y e_z - d_z e_y
      term<AX_ID, AX_ID, AX_D>(m, i, j, k, -1.0, DZB_EMIT(LOC_EY, m.off<LOC_EY>()));
      term<AX_ID, AX_D, AX_ID>(m, i, j, k, 1.0, DZB_EMIT(LOC_EZ, m.off<LOC_EZ>()));
    }
    if (COMP == 1) {  // d_z e_x - d_x e_z
      term<AX_ID, AX_ID, AX_D>(m, i, j, k, 1.0, DZB_EMIT(LOC_EX, m.off<LOC_EX>()));
      term<AX_D, AX_ID, AX_ID>(m, i, j, k, -1.0, DZB_EMIT(LOC_EZ, m.off<LOC_EZ>()));
    }
    if (COMP == 2) {  // d_x e_y - d_y e_x
      term<AX_ID, AX_D, AX_ID>(m, i, j, k, -1.0, DZB_EMIT(LOC_EX, m.off<LOC_EX>()));
      term<AX_D, AX_ID, AX_ID>(m, i, j, k, 1.0, DZB_EMIT(LOC_EY, m.off<LOC_EY>()));
    }
  }
};
template <> struct Rows<DZB_OP_CURL, true> {
  static constexpr int NCOMP = 3;
  static constexpr int MAXNNZ = 4;
  template <int COMP> static constexpr int loc() { return COMP == 0 ? LOC_EX : COMP == 1 ? LOC_EY : LOC_EZ; }
  template <int COMP> __host__ __device__ static i64 row_off(const TM &m) { return m.off<loc<COMP>()>(); }
  __host__ __device__ static i64 nrows(const TM &m) { return m.nE(); }
  __host__ __device__ static i64 ncols(const TM &m) { return m.nF(); }
  template <int COMP, class E> __device__ static void row(const TM &m, i64 i, i64 j, i64 k, E &&emit) {
    if (COMP == 0) {
      term<AX_ID, AX_ID, AX_DT>(m, i, j, k, 1.0, DZB_EMIT(LOC_FY, m.off<LOC_FY>()));
      term<AX_ID, AX_DT, AX_ID>(m, i, j, k, -1.0, DZB_EMIT(LOC_FZ, m.off<LOC_FZ>()));
    }
    if (COMP == 1) {
      term<AX_ID, AX_ID, AX_DT>(m, i, j, k, -1.0, DZB_EMIT(LOC_FX, m.off<LOC_FX>()));
      term<AX_DT, AX_ID, AX_ID>(m, i, j, k, 1.0, DZB_EMIT(LOC_FZ, m.off<LOC_FZ>()));
    }
    if (COMP == 2) {
      term<AX_ID, AX_DT, AX_ID>(m, i, j, k, 1.0, DZB_EMIT(LOC_FX, m.off<LOC_FX>()));
      term<AX_DT, AX_ID, AX_ID>(m, i, j, k, -1.0, DZB_EMIT(LOC_FY, m.off<LOC_FY>()));
    }
  }
};

// ---- nodal_gradient ------------------------------------------------------------------
template <> struct Rows<DZB_OP_GRAD, false> {
  static constexpr int NCOMP = 3;
  static constexpr int MAXNNZ = 2;
  template <int COMP> static constexpr int loc() { return COMP == 0 ? LOC_EX : COMP == 1 ? LOC_EY : LOC_EZ; }
  template <int COMP> __host__ __device__ static i64 row_off(const TM &m) { return m.off<loc<COMP>()>(); }
  __host__ __device__ static i64 nrows(const TM &m) { return m.nE(); }
  __host__ __device__ static i64 ncols(const TM &m) { return m.nN(); }
  template <int COMP, class E> __device__ static void row(const TM &m, i64 i, i64 j, i64 k, E &&emit) {
    if (COMP == 0) term<AX_D, AX_ID, AX_ID>(m, i, j, k, 1.0, DZB_EMIT(LOC_N, 0));
    if (COMP == 1) term<AX_ID, AX_D, AX_ID>(m, i, j, k, 1.0, DZB_EMIT(LOC_N, 0));
    if (COMP == 2) term<AX_ID, AX_ID, AX_D>(m, i, j, k, 1.0, DZB_EMIT(LOC_N, 0));
  }
};
template <> struct Rows<DZB_OP_GRAD, true> {
  static constexpr int NCOMP = 1;
  static constexpr int MAXNNZ = 6;
  template <int COMP> static constexpr int loc() { return LOC_N; }
  template <int COMP> __host__ __device__ static i64 row_off(const TM &m) { return 0; }
  __host__ __device__ static i64 nrows(const TM &m) { return m.nN(); }
  __host__ __device__ static i64 ncols(const TM &m) { return m.nE(); }
  template <int COMP, class E> __device__ static void row(const TM &m, i64 i, i64 j, i64 k, E &&emit) {
    term<AX_DT, AX_ID, AX_ID>(m, i, j, k, 1.0, DZB_EMIT(LOC_EX, m.off<LOC_EX>()));
    term<AX_ID, AX_DT, AX_ID>(m, i, j, k, 1.0, DZB_EMIT(LOC_EY, m.off<LOC_EY>()));
    term<AX_ID, AX_ID, AX_DT>(m, i, j, k, 1.0, DZB_EMIT(LOC_EZ, m.off<LOC_EZ>()));
  }
};

// ---- averaging: faces -> cells -------------------------------------------------------
// SC: overall scale; STACK: 0 = scalar average of the three components (hstack, x 1/3),
// 1 = vector average (block diagonal), 2/3/4 = single component x/y/z (local columns).
template <bool EDGE, int COMP> struct AveSrc {
  static constexpr int LOC = EDGE ? (COMP == 0 ? LOC_EX : COMP == 1 ? LOC_EY : LOC_EZ)
                                  : (COMP == 0 ? LOC_FX : COMP == 1 ? LOC_FY : LOC_FZ);
  // node-extent axes of the source get averaged
  static constexpr int OX = (LOC & 1) ? AX_A : AX_ID;
  static constexpr int OY = (LOC & 2) ? AX_A : AX_ID;
  static constexpr int OZ = (LOC & 4) ? AX_A : AX_ID;
  static constexpr int TX = (LOC & 1) ? AX_AT : AX_ID;
  static constexpr int TY = (LOC & 2) ? AX_AT : AX_ID;
  static constexpr int TZ = (LOC & 4) ? AX_AT : AX_ID;
};

// scalar average: nC x n{F,E}
template <bool EDGE> struct RowsAveScalar {
  static constexpr int NCOMP = 1;
  static constexpr int MAXNNZ = EDGE ? 12 : 6;
  template <int COMP> static constexpr int loc() { return LOC_C; }
  template <int COMP> __host__ __device__ static i64 row_off(const TM &m) { return 0; }
  __host__ __device__ static i64 nrows(const TM &m) { return m.nC(); }
  __host__ __device__ static i64 ncols(const TM &m) { return EDGE ? m.nE() : m.nF(); }
  template <int COMP, class E> __device__ static void row(const TM &m, i64 i, i64 j, i64 k, E &&emit) {
    typedef AveSrc<EDGE, 0> S0; typedef AveSrc<EDGE, 1> S1; typedef AveSrc<EDGE, 2> S2;
    term<S0::OX, S0::OY, S0::OZ>(m, i, j, k, DZB_THIRD, DZB_EMIT(S0::LOC, m.off<S0::LOC>()));
    term<S1::OX, S1::OY, S1::OZ>(m, i, j, k, DZB_THIRD, DZB_EMIT(S1::LOC, m.off<S1::LOC>()));
    term<S2::OX, S2::OY, S2::OZ>(m, i, j, k, DZB_THIRD, DZB_EMIT(S2::LOC, m.off<S2::LOC>()));
  }
};
template <bool EDGE> struct RowsAveScalarT {
  static constexpr int NCOMP = 3;
  static constexpr int MAXNNZ = EDGE ? 4 : 2;
  template <int COMP> static constexpr int loc() { return AveSrc<EDGE, COMP>::LOC; }
  template <int COMP> __host__ __device__ static i64 row_off(const TM &m) { return m.off<loc<COMP>()>(); }
  __host__ __device__ static i64 nrows(const TM &m) { return EDGE ? m.nE() : m.nF(); }
  __host__ __device__ static i64 ncols(const TM &m) { return m.nC(); }
  template <int COMP, class E> __device__ static void row(const TM &m, i64 i, i64 j, i64 k, E &&emit) {
    typedef AveSrc<EDGE, COMP> S;
    term<S::TX, S::TY, S::TZ>(m, i, j, k, DZB_THIRD, DZB_EMIT(LOC_C, 0));
  }
};
// vector average: 3nC x n{F,E}, block diagonal
template <bool EDGE> struct RowsAveVector {
  static constexpr int NCOMP = 3;
  static constexpr int MAXNNZ = EDGE ? 4 : 2;
  template <int COMP> static constexpr int loc() { return LOC_C; }
  template <int COMP> __host__ __device__ static i64 row_off(const TM &m) { return COMP * m.nC(); }
  __host__ __device__ static i64 nrows(const TM &m) { return 3 * m.nC(); }
  __host__ __device__ static i64 ncols(const TM &m) { return EDGE ? m.nE() : m.nF(); }
  template <int COMP, class E> __device__ static void row(const TM &m, i64 i, i64 j, i64 k, E &&emit) {
    typedef AveSrc<EDGE, COMP> S;
    term<S::OX, S::OY, S::OZ>(m, i, j, k, 1.0, DZB_EMIT(S::LOC, m.off<S::LOC>()));
  }
};
template <bool EDGE> struct RowsAveVectorT {
  static constexpr int NCOMP = 3;
  static constexpr int MAXNNZ = EDGE ? 4 : 2;
  template <int COMP> static constexpr int loc() { return AveSrc<EDGE, COMP>::LOC; }
  template <int COMP> __host__ __device__ static i64 row_off(const TM &m) { return m.off<loc<COMP>()>(); }
  __host__ __device__ static i64 nrows(const TM &m) { return EDGE ? m.nE() : m.nF(); }
  __host__ __device__ static i64 ncols(const TM &m) { return 3 * m.nC(); }
  template <int COMP, class E> __device__ static void row(const TM &m, i64 i, i64 j, i64 k, E &&emit) {
    typedef AveSrc<EDGE, COMP> S;
    term<S::TX, S::TY, S::TZ>(m, i, j, k, 1.0, DZB_EMIT(LOC_C, COMP * m.nC()));
  }
};
// single component: nC x n{F,E}{x,y,z} with component-local columns
template <bool EDGE, int C> struct RowsAveComp {
  static constexpr int NCOMP = 1;
  static constexpr int MAXNNZ = EDGE ? 4 : 2;
  template <int COMP> static constexpr int loc() { return LOC_C; }
  template <int COMP> __host__ __device__ static i64 row_off(const TM &m) { return 0; }
  __host__ __device__ static i64 nrows(const TM &m) { return m.nC(); }
  __host__ __device__ static i64 ncols(const TM &m) { return m.count<AveSrc<EDGE, C>::LOC>(); }
  template <int COMP, class E> __device__ static void row(const TM &m, i64 i, i64 j, i64 k, E &&emit) {
    typedef AveSrc<EDGE, C> S;
    term<S::OX, S::OY, S::OZ>(m, i, j, k, 1.0, DZB_EMIT(S::LOC, 0));
  }
};
template <bool EDGE, int C> struct RowsAveCompT {
  static constexpr int NCOMP = 1;
  static constexpr int MAXNNZ = EDGE ? 4 : 2;
  template <int COMP> static constexpr int loc() { return AveSrc<EDGE, C>::LOC; }
  template <int COMP> __host__ __device__ static i64 row_off(const TM &m) { return 0; }
  __host__ __device__ static i64 nrows(const TM &m) { return m.count<AveSrc<EDGE, C>::LOC>(); }
  __host__ __device__ static i64 ncols(const TM &m) { return m.nC(); }
  template <int COMP, class E> __device__ static void row(const TM &m, i64 i, i64 j, i64 k, E &&emit) {
    typedef AveSrc<EDGE, C> S;
    term<S::TX, S::TY, S::TZ>(m, i, j, k, 1.0, DZB_EMIT(LOC_C, 0));
  }
};

template <> struct Rows<DZB_OP_AVE_F2CC, false> : RowsAveScalar<false> {};
template <> struct Rows<DZB_OP_AVE_F2CC, true> : RowsAveScalarT<false> {};
template <> struct Rows<DZB_OP_AVE_E2CC, false> : RowsAveScalar<true> {};
template <> struct Rows<DZB_OP_AVE_E2CC, true> : RowsAveScalarT<true> {};
template <> struct Rows<DZB_OP_AVE_F2CCV, false> : RowsAveVector<false> {};
template <> struct Rows<DZB_OP_AVE_F2CCV, true> : RowsAveVectorT<false> {};
template <> struct Rows<DZB_OP_AVE_E2CCV, false> : RowsAveVector<true> {};
template <> struct Rows<DZB_OP_AVE_E2CCV, true> : RowsAveVectorT<true> {};
template <> struct Rows<DZB_OP_AVE_FX2CC, false> : RowsAveComp<false, 0> {};
template <> struct Rows<DZB_OP_AVE_FX2CC, true> : RowsAveCompT<false, 0> {};
template <> struct Rows<DZB_OP_AVE_FY2CC, false> : RowsAveComp<false, 1> {};
template <> struct Rows<DZB_OP_AVE_FY2CC, true> : RowsAveCompT<false, 1> {};
template <> struct Rows<DZB_OP_AVE_FZ2CC, false> : RowsAveComp<false, 2> {};
template <> struct Rows<DZB_OP_AVE_FZ2CC, true> : RowsAveCompT<false, 2> {};
template <> struct Rows<DZB_OP_AVE_EX2CC, false> : RowsAveComp<true, 0> {};
template <> struct Rows<DZB_OP_AVE_EX2CC, true> : RowsAveCompT<true, 0> {};
template <> struct Rows<DZB_OP_AVE_EY2CC, false> : RowsAveComp<true, 1> {};
template <> struct Rows<DZB_OP_AVE_EY2CC, true> : RowsAveCompT<true, 1> {};
template <> struct Rows<DZB_OP_AVE_EZ2CC, false> : RowsAveComp<true, 2> {};
template <> struct Rows<DZB_OP_AVE_EZ2CC, true> : RowsAveCompT<true, 2> {};

// ---- nodes -> cells ------------------------------------------------------------------
template <> struct Rows<DZB_OP_AVE_N2CC, false> {
  static constexpr int NCOMP = 1;
  static constexpr int MAXNNZ = 8;
  template <int COMP> static constexpr int loc() { return LOC_C; }
  template <int COMP> __host__ __device__ static i64 row_off(const TM &m) { return 0; }
  __host__ __device__ static i64 nrows(const TM &m) { return m.nC(); }
  __host__ __device__ static i64 ncols(const TM &m) { return m.nN(); }
  template <int COMP, class E> __device__ static void row(const TM &m, i64 i, i64 j, i64 k, E &&emit) {
    term<AX_A, AX_A, AX_A>(m, i, j, k, 1.0, DZB_EMIT(LOC_N, 0));
  }
};
template <> struct Rows<DZB_OP_AVE_N2CC, true> {
  static constexpr int NCOMP = 1;
  static constexpr int MAXNNZ = 8;
  template <int COMP> static constexpr int loc() { return LOC_N; }
  template <int COMP> __host__ __device__ static i64 row_off(const TM &m) { return 0; }
  __host__ __device__ static i64 nrows(const TM &m) { return m.nN(); }
  __host__ __device__ static i64 ncols(const TM &m) { return m.nC(); }
  template <int COMP, class E> __device__ static void row(const TM &m, i64 i, i64 j, i64 k, E &&emit) {
    term<AX_AT, AX_AT, AX_AT>(m, i, j, k, 1.0, DZB_EMIT(LOC_C, 0));
  }
};

}  // namespace dzb
