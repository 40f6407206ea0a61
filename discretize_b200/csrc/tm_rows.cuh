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
  AX_AT = 4,  // node i  <- cells i-1, i : (1/2, 1/2)             (av^T)
  AX_AE = 5,  // node i  <- cells i-1, i : (1/2, 1/2), weight 1 on the two boundary nodes  (av_extrap)
  AX_AET = 6, // cell i  <- nodes i, i+1 : av_extrap^T
  AX_GC = 7,  // node i  <- cells i-1, i : cell gradient (-s, +s) with ghost-point boundary rows (_ddxCellGrad)
  AX_GCT = 8, // cell i  <- nodes i, i+1 : its transpose
  AX_D2 = 9,  // cell i  <- nodes i, i+1 : (1/h_i^2, 1/h_i^2)          entrywise square of ddx
  AX_DT2 = 10 // node i  <- cells i-1, i : (1/h_{i-1}^2, 1/h_i^2)      entrywise square of ddx^T
};

// flags of AX_GC / AX_GCT for one axis
enum GcFlags : int {
  GC_DIRICHLET_LO = 1,  // boundary node 0: +2 s (else Neumann: 0)
  GC_DIRICHLET_HI = 2,  // boundary node n: -2 s
  GC_SCALED = 4         // s = S/V_face = 1 / (distance between the two cell centres); else s = 1 (the stencil)
};

// value of the cell-gradient stencil between node i and cell c (c = i-1 or c = i); *keep says whether the
// reference stores the entry (Neumann boundary rows: the stencil keeps an explicit zero, the scaled
// gradient -- a csr_matmat product -- drops it)
__device__ __forceinline__ double gc_value(i64 i, i64 c, i64 n, const double *__restrict__ h,
                                           const double *__restrict__ rh, int flags, bool *keep) {
  const bool scaled = flags & GC_SCALED;
  *keep = true;
  if (i == 0) {
    if (!(flags & GC_DIRICHLET_LO)) { *keep = !scaled; return 0.0; }
    return scaled ? 2.0 * __ldg(rh) : 2.0;
  }
  if (i == n) {
    if (!(flags & GC_DIRICHLET_HI)) { *keep = !scaled; return 0.0; }
    return scaled ? -2.0 * __ldg(rh + n - 1) : -2.0;
  }
  const double s = scaled ? 1.0 / (0.5 * __ldg(h + i - 1) + 0.5 * __ldg(h + i)) : 1.0;
  return c == i ? s : -s;
}

// Entries of a 1-D stencil at row i: two FIXED slots (slot 0 = lower source index, slot 1 = upper) and a bit
// mask saying which slots exist.  Fixed slots keep everything in registers: indexing the slots with a run-time
// count puts the arrays in local memory, which made every transposed operator L1-bound (ncu, round 1).
template <int OP>
__device__ __forceinline__ int ax_entries(i64 i, i64 n, const double *__restrict__ rh, const double *__restrict__ h,
                                          int flags, i64 (&ix)[2], double (&w)[2]) {
  if (OP == AX_ID) {
    ix[0] = i; w[0] = 1.0;
    ix[1] = i; w[1] = 0.0;
    return 1;
  } else if (OP == AX_D) {
    double r = __ldg(rh + i);
    ix[0] = i; w[0] = -r;
    ix[1] = i + 1; w[1] = r;
    return 3;
  } else if (OP == AX_A) {
    ix[0] = i; w[0] = 0.5;
    ix[1] = i + 1; w[1] = 0.5;
    return 3;
  } else if (OP == AX_DT) {
    const bool lo = i >= 1, hi = i < n;
    ix[0] = i - 1; w[0] = lo ? __ldg(rh + i - 1) : 0.0;
    ix[1] = i; w[1] = hi ? -__ldg(rh + i) : 0.0;
    return (lo ? 1 : 0) | (hi ? 2 : 0);
  } else if (OP == AX_D2) {
    const double r = __ldg(rh + i);
    ix[0] = i; w[0] = r * r;
    ix[1] = i + 1; w[1] = r * r;
    return 3;
  } else if (OP == AX_DT2) {
    const bool lo = i >= 1, hi = i < n;
    const double a = lo ? __ldg(rh + i - 1) : 0.0, b = hi ? __ldg(rh + i) : 0.0;
    ix[0] = i - 1; w[0] = a * a;
    ix[1] = i; w[1] = b * b;
    return (lo ? 1 : 0) | (hi ? 2 : 0);
  } else if (OP == AX_AT) {
    ix[0] = i - 1; w[0] = 0.5;
    ix[1] = i; w[1] = 0.5;
    return (i >= 1 ? 1 : 0) | (i < n ? 2 : 0);
  } else if (OP == AX_AE) {
    const bool lo = i >= 1, hi = i < n;
    ix[0] = i - 1; w[0] = hi ? 0.5 : 1.0;
    ix[1] = i; w[1] = lo ? 0.5 : 1.0;
    return (lo ? 1 : 0) | (hi ? 2 : 0);
  } else if (OP == AX_AET) {
    ix[0] = i; w[0] = i == 0 ? 1.0 : 0.5;
    ix[1] = i + 1; w[1] = i + 1 == n ? 1.0 : 0.5;
    return 3;
  } else if (OP == AX_GC) {
    bool k0 = false, k1 = false;
    ix[0] = i - 1; w[0] = i >= 1 ? gc_value(i, i - 1, n, h, rh, flags, &k0) : 0.0;
    ix[1] = i; w[1] = i < n ? gc_value(i, i, n, h, rh, flags, &k1) : 0.0;
    return ((i >= 1 && k0) ? 1 : 0) | ((i < n && k1) ? 2 : 0);
  } else {  // AX_GCT: cell i <- node i (entry G[i,i]) and node i+1 (entry G[i+1,i])
    bool k0, k1;
    ix[0] = i; w[0] = gc_value(i, i, n, h, rh, flags, &k0);
    ix[1] = i + 1; w[1] = gc_value(i + 1, i, n, h, rh, flags, &k1);
    return (k0 ? 1 : 0) | (k1 ? 2 : 0);
  }
}

// Entries of one tensor-product term; emit(ix, iy, iz, value) in source-location indices.
template <int OX, int OY, int OZ, class E>
__device__ __forceinline__ void term(const TM &m, i64 i, i64 j, i64 k, double scale, E &&emit, int fx = 0, int fy = 0,
                                     int fz = 0) {
  i64 xs[2], ys[2], zs[2];
  double wx[2], wy[2], wz[2];
  const int mx = ax_entries<OX>(i, m.nx, m.rhx, m.hx, fx, xs, wx);
  const int my = ax_entries<OY>(j, m.ny, m.rhy, m.hy, fy, ys, wy);
  const int mz = ax_entries<OZ>(k, m.nz, m.rhz, m.hz, fz, zs, wz);
#pragma unroll
  for (int c = 0; c < (OZ == AX_ID ? 1 : 2); ++c) {
#pragma unroll
    for (int b = 0; b < (OY == AX_ID ? 1 : 2); ++b) {
#pragma unroll
      for (int a = 0; a < (OX == AX_ID ? 1 : 2); ++a) {
        if (((mx >> a) & 1) && ((my >> b) & 1) && ((mz >> c) & 1)) emit(xs[a], ys[b], zs[c], ((scale * wx[a]) * wy[b]) * wz[c]);
      }
    }
  }
}

#define DZB_THIRD (1.0 / 3.0)

// Rows<OP, TR>: NCOMP row blocks; block c lives on location loc(c) with rows starting at
// row_off(c).  row<COMP>(m,i,j,k,emit) calls emit(flat_column, value).
struct RowsBase {
  int flags = 0;  // runtime options of an operator (cell-gradient boundary conditions); 0 for most
};
template <int OP, bool TR> struct Rows;

#define DZB_EMIT(SRC, COLOFF) [&](i64 a, i64 b, i64 c, double v) { emit((COLOFF) + m.lidx<SRC>(a, b, c), v); }

// ---- face_divergence ---------------------------------------------------------------
template <> struct Rows<DZB_OP_DIV, false> : RowsBase {
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
template <> struct Rows<DZB_OP_DIV, true> : RowsBase {
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
template <> struct Rows<DZB_OP_CURL, false> : RowsBase {
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
template <> struct Rows<DZB_OP_CURL, true> : RowsBase {
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
template <> struct Rows<DZB_OP_GRAD, false> : RowsBase {
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
template <> struct Rows<DZB_OP_GRAD, true> : RowsBase {
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
template <bool EDGE> struct RowsAveScalar : RowsBase {
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
template <bool EDGE> struct RowsAveScalarT : RowsBase {
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
template <bool EDGE> struct RowsAveVector : RowsBase {
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
template <bool EDGE> struct RowsAveVectorT : RowsBase {
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
template <bool EDGE, int C> struct RowsAveComp : RowsBase {
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
template <bool EDGE, int C> struct RowsAveCompT : RowsBase {
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
template <> struct Rows<DZB_OP_AVE_N2CC, false> : RowsBase {
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
template <> struct Rows<DZB_OP_AVE_N2CC, true> : RowsBase {
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

// =====================================================================================================
// cell-centred operators
// =====================================================================================================
template <int C> struct FaceOf { static constexpr int LOC = C == 0 ? LOC_FX : C == 1 ? LOC_FY : LOC_FZ; };
template <int C> struct EdgeOf { static constexpr int LOC = C == 0 ? LOC_EX : C == 1 ? LOC_EY : LOC_EZ; };

// ---- average_cell_to_face / average_cell_vector_to_face (differential_operators.py:2789-2865) -------
// rows on faces: av_extrap along the face normal.  VEC: block-diagonal columns (3 nC).
template <bool VEC> struct RowsCC2F : RowsBase {
  static constexpr int NCOMP = 3;
  static constexpr int MAXNNZ = 2;
  template <int COMP> static constexpr int loc() { return FaceOf<COMP>::LOC; }
  template <int COMP> __host__ __device__ static i64 row_off(const TM &m) { return m.off<loc<COMP>()>(); }
  __host__ __device__ static i64 nrows(const TM &m) { return m.nF(); }
  __host__ __device__ static i64 ncols(const TM &m) { return (VEC ? 3 : 1) * m.nC(); }
  template <int COMP, class E> __device__ void row(const TM &m, i64 i, i64 j, i64 k, E &&emit) const {
    const i64 co = VEC ? COMP * m.nC() : 0;
    term<(COMP == 0 ? AX_AE : AX_ID), (COMP == 1 ? AX_AE : AX_ID), (COMP == 2 ? AX_AE : AX_ID)>(m, i, j, k, 1.0,
                                                                                                 DZB_EMIT(LOC_C, co));
  }
};
template <bool VEC> struct RowsCC2FT : RowsBase {
  static constexpr int NCOMP = VEC ? 3 : 1;
  static constexpr int MAXNNZ = VEC ? 2 : 6;
  template <int COMP> static constexpr int loc() { return LOC_C; }
  template <int COMP> __host__ __device__ static i64 row_off(const TM &m) { return VEC ? COMP * m.nC() : 0; }
  __host__ __device__ static i64 nrows(const TM &m) { return (VEC ? 3 : 1) * m.nC(); }
  __host__ __device__ static i64 ncols(const TM &m) { return m.nF(); }
  template <int COMP, class E> __device__ void row(const TM &m, i64 i, i64 j, i64 k, E &&emit) const {
    if (!VEC || COMP == 0) term<AX_AET, AX_ID, AX_ID>(m, i, j, k, 1.0, DZB_EMIT(LOC_FX, m.off<LOC_FX>()));
    if (!VEC || COMP == 1) term<AX_ID, AX_AET, AX_ID>(m, i, j, k, 1.0, DZB_EMIT(LOC_FY, m.off<LOC_FY>()));
    if (!VEC || COMP == 2) term<AX_ID, AX_ID, AX_AET>(m, i, j, k, 1.0, DZB_EMIT(LOC_FZ, m.off<LOC_FZ>()));
  }
};
template <> struct Rows<DZB_OP_AVE_CC2F, false> : RowsCC2F<false> {};
template <> struct Rows<DZB_OP_AVE_CC2F, true> : RowsCC2FT<false> {};
template <> struct Rows<DZB_OP_AVE_CCV2F, false> : RowsCC2F<true> {};
template <> struct Rows<DZB_OP_AVE_CCV2F, true> : RowsCC2FT<true> {};

// ---- average_cell_to_edge (:2867-2892): av_extrap on the two axes transverse to the edge --------------
template <> struct Rows<DZB_OP_AVE_CC2E, false> : RowsBase {
  static constexpr int NCOMP = 3;
  static constexpr int MAXNNZ = 4;
  template <int COMP> static constexpr int loc() { return EdgeOf<COMP>::LOC; }
  template <int COMP> __host__ __device__ static i64 row_off(const TM &m) { return m.off<loc<COMP>()>(); }
  __host__ __device__ static i64 nrows(const TM &m) { return m.nE(); }
  __host__ __device__ static i64 ncols(const TM &m) { return m.nC(); }
  template <int COMP, class E> __device__ void row(const TM &m, i64 i, i64 j, i64 k, E &&emit) const {
    term<(COMP == 0 ? AX_ID : AX_AE), (COMP == 1 ? AX_ID : AX_AE), (COMP == 2 ? AX_ID : AX_AE)>(m, i, j, k, 1.0,
                                                                                                 DZB_EMIT(LOC_C, 0));
  }
};
template <> struct Rows<DZB_OP_AVE_CC2E, true> : RowsBase {
  static constexpr int NCOMP = 1;
  static constexpr int MAXNNZ = 12;
  template <int COMP> static constexpr int loc() { return LOC_C; }
  template <int COMP> __host__ __device__ static i64 row_off(const TM &m) { return 0; }
  __host__ __device__ static i64 nrows(const TM &m) { return m.nC(); }
  __host__ __device__ static i64 ncols(const TM &m) { return m.nE(); }
  template <int COMP, class E> __device__ void row(const TM &m, i64 i, i64 j, i64 k, E &&emit) const {
    term<AX_ID, AX_AET, AX_AET>(m, i, j, k, 1.0, DZB_EMIT(LOC_EX, m.off<LOC_EX>()));
    term<AX_AET, AX_ID, AX_AET>(m, i, j, k, 1.0, DZB_EMIT(LOC_EY, m.off<LOC_EY>()));
    term<AX_AET, AX_AET, AX_ID>(m, i, j, k, 1.0, DZB_EMIT(LOC_EZ, m.off<LOC_EZ>()));
  }
};

// ---- average_edge_to_face (:3205-3234): each face takes the mean of its four bounding edges -----------
// face axis f, edge axis e != f: average along the third axis t (node extent on the edge, cell extent on the face)
template <> struct Rows<DZB_OP_AVE_E2F, false> : RowsBase {
  static constexpr int NCOMP = 3;
  static constexpr int MAXNNZ = 4;
  template <int COMP> static constexpr int loc() { return FaceOf<COMP>::LOC; }
  template <int COMP> __host__ __device__ static i64 row_off(const TM &m) { return m.off<loc<COMP>()>(); }
  __host__ __device__ static i64 nrows(const TM &m) { return m.nF(); }
  __host__ __device__ static i64 ncols(const TM &m) { return m.nE(); }
  template <int COMP, class E> __device__ void row(const TM &m, i64 i, i64 j, i64 k, E &&emit) const {
    if (COMP == 0) {
      term<AX_ID, AX_ID, AX_A>(m, i, j, k, 0.5, DZB_EMIT(LOC_EY, m.off<LOC_EY>()));
      term<AX_ID, AX_A, AX_ID>(m, i, j, k, 0.5, DZB_EMIT(LOC_EZ, m.off<LOC_EZ>()));
    }
    if (COMP == 1) {
      term<AX_ID, AX_ID, AX_A>(m, i, j, k, 0.5, DZB_EMIT(LOC_EX, m.off<LOC_EX>()));
      term<AX_A, AX_ID, AX_ID>(m, i, j, k, 0.5, DZB_EMIT(LOC_EZ, m.off<LOC_EZ>()));
    }
    if (COMP == 2) {
      term<AX_ID, AX_A, AX_ID>(m, i, j, k, 0.5, DZB_EMIT(LOC_EX, m.off<LOC_EX>()));
      term<AX_A, AX_ID, AX_ID>(m, i, j, k, 0.5, DZB_EMIT(LOC_EY, m.off<LOC_EY>()));
    }
  }
};
template <> struct Rows<DZB_OP_AVE_E2F, true> : RowsBase {
  static constexpr int NCOMP = 3;
  static constexpr int MAXNNZ = 4;
  template <int COMP> static constexpr int loc() { return EdgeOf<COMP>::LOC; }
  template <int COMP> __host__ __device__ static i64 row_off(const TM &m) { return m.off<loc<COMP>()>(); }
  __host__ __device__ static i64 nrows(const TM &m) { return m.nE(); }
  __host__ __device__ static i64 ncols(const TM &m) { return m.nF(); }
  template <int COMP, class E> __device__ void row(const TM &m, i64 i, i64 j, i64 k, E &&emit) const {
    if (COMP == 0) {  // Ex feeds Fy (averaged along z) and Fz (along y)
      term<AX_ID, AX_ID, AX_AT>(m, i, j, k, 0.5, DZB_EMIT(LOC_FY, m.off<LOC_FY>()));
      term<AX_ID, AX_AT, AX_ID>(m, i, j, k, 0.5, DZB_EMIT(LOC_FZ, m.off<LOC_FZ>()));
    }
    if (COMP == 1) {  // Ey feeds Fx (along z) and Fz (along x)
      term<AX_ID, AX_ID, AX_AT>(m, i, j, k, 0.5, DZB_EMIT(LOC_FX, m.off<LOC_FX>()));
      term<AX_AT, AX_ID, AX_ID>(m, i, j, k, 0.5, DZB_EMIT(LOC_FZ, m.off<LOC_FZ>()));
    }
    if (COMP == 2) {  // Ez feeds Fx (along y) and Fy (along x)
      term<AX_ID, AX_AT, AX_ID>(m, i, j, k, 0.5, DZB_EMIT(LOC_FX, m.off<LOC_FX>()));
      term<AX_AT, AX_ID, AX_ID>(m, i, j, k, 0.5, DZB_EMIT(LOC_FY, m.off<LOC_FY>()));
    }
  }
};

// ---- average_node_to_edge (:3254-3319) / average_node_to_face (:3321-3382) -----------------------------
// FACE = false: mean of the two end nodes of an edge; FACE = true: mean of the four corner nodes of a face
template <bool FACE> struct RowsN2X : RowsBase {
  static constexpr int NCOMP = 3;
  static constexpr int MAXNNZ = FACE ? 4 : 2;
  template <int COMP> static constexpr int loc() { return FACE ? FaceOf<COMP>::LOC : EdgeOf<COMP>::LOC; }
  template <int COMP> __host__ __device__ static i64 row_off(const TM &m) { return m.off<loc<COMP>()>(); }
  __host__ __device__ static i64 nrows(const TM &m) { return FACE ? m.nF() : m.nE(); }
  __host__ __device__ static i64 ncols(const TM &m) { return m.nN(); }
  template <int COMP, class E> __device__ void row(const TM &m, i64 i, i64 j, i64 k, E &&emit) const {
    constexpr int L = loc<COMP>();  // average along the cell-extent axes of the row location
    term<((L & 1) ? AX_ID : AX_A), ((L & 2) ? AX_ID : AX_A), ((L & 4) ? AX_ID : AX_A)>(m, i, j, k, 1.0, DZB_EMIT(LOC_N, 0));
  }
};
template <bool FACE> struct RowsN2XT : RowsBase {
  static constexpr int NCOMP = 1;
  static constexpr int MAXNNZ = FACE ? 12 : 6;
  template <int COMP> static constexpr int loc() { return LOC_N; }
  template <int COMP> __host__ __device__ static i64 row_off(const TM &m) { return 0; }
  __host__ __device__ static i64 nrows(const TM &m) { return m.nN(); }
  __host__ __device__ static i64 ncols(const TM &m) { return FACE ? m.nF() : m.nE(); }
  template <int C, class E> __device__ static void comp(const TM &m, i64 i, i64 j, i64 k, E &&emit) {
    constexpr int L = FACE ? FaceOf<C>::LOC : EdgeOf<C>::LOC;
    term<((L & 1) ? AX_ID : AX_AT), ((L & 2) ? AX_ID : AX_AT), ((L & 4) ? AX_ID : AX_AT)>(m, i, j, k, 1.0,
                                                                                         DZB_EMIT(L, m.off<L>()));
  }
  template <int COMP, class E> __device__ void row(const TM &m, i64 i, i64 j, i64 k, E &&emit) const {
    comp<0>(m, i, j, k, emit);
    comp<1>(m, i, j, k, emit);
    comp<2>(m, i, j, k, emit);
  }
};
template <> struct Rows<DZB_OP_AVE_N2E, false> : RowsN2X<false> {};
template <> struct Rows<DZB_OP_AVE_N2E, true> : RowsN2XT<false> {};
template <> struct Rows<DZB_OP_AVE_N2F, false> : RowsN2X<true> {};
template <> struct Rows<DZB_OP_AVE_N2F, true> : RowsN2XT<true> {};

// ---- cell gradient (:41-82, 1413-1443, 1445-1595): faces <- cells with ghost-point boundary rows ------
// flags: bits 0-1 x-axis GcFlags (lower, upper Dirichlet), bits 2-3 y, bits 4-5 z, bit 6 scaled by S / V_face.
// AXIS = -1: all three components stacked (nF x nC); AXIS = 0/1/2: one component block (nF_axis x nC).
#define DZB_CG_SCALED 64
template <int AXIS> struct RowsCellGrad : RowsBase {
  static constexpr int NCOMP = AXIS < 0 ? 3 : 1;
  static constexpr int MAXNNZ = 2;
  template <int COMP> static constexpr int comp() { return AXIS < 0 ? COMP : AXIS; }
  template <int COMP> static constexpr int loc() { return FaceOf<comp<COMP>()>::LOC; }
  template <int COMP> __host__ __device__ static i64 row_off(const TM &m) { return AXIS < 0 ? m.off<loc<COMP>()>() : 0; }
  __host__ __device__ static i64 nrows(const TM &m) {
    return AXIS < 0 ? m.nF() : AXIS == 0 ? m.count<LOC_FX>() : AXIS == 1 ? m.count<LOC_FY>() : m.count<LOC_FZ>();
  }
  __host__ __device__ static i64 ncols(const TM &m) { return m.nC(); }
  __device__ int axis_flags(int a) const { return ((flags >> (2 * a)) & 3) | ((flags & DZB_CG_SCALED) ? GC_SCALED : 0); }
  template <int COMP, class E> __device__ void row(const TM &m, i64 i, i64 j, i64 k, E &&emit) const {
    constexpr int A = comp<COMP>();
    term<(A == 0 ? AX_GC : AX_ID), (A == 1 ? AX_GC : AX_ID), (A == 2 ? AX_GC : AX_ID)>(
        m, i, j, k, 1.0, DZB_EMIT(LOC_C, 0), axis_flags(0), axis_flags(1), axis_flags(2));
  }
};
template <int AXIS> struct RowsCellGradT : RowsBase {
  static constexpr int NCOMP = 1;
  static constexpr int MAXNNZ = AXIS < 0 ? 6 : 2;
  template <int COMP> static constexpr int loc() { return LOC_C; }
  template <int COMP> __host__ __device__ static i64 row_off(const TM &m) { return 0; }
  __host__ __device__ static i64 nrows(const TM &m) { return m.nC(); }
  __host__ __device__ static i64 ncols(const TM &m) { return RowsCellGrad<AXIS>::nrows(m); }
  __device__ int axis_flags(int a) const { return ((flags >> (2 * a)) & 3) | ((flags & DZB_CG_SCALED) ? GC_SCALED : 0); }
  template <int A, class E> __device__ void comp(const TM &m, i64 i, i64 j, i64 k, E &&emit) const {
    constexpr int L = FaceOf<A>::LOC;
    const i64 co = AXIS < 0 ? m.off<L>() : 0;
    term<(A == 0 ? AX_GCT : AX_ID), (A == 1 ? AX_GCT : AX_ID), (A == 2 ? AX_GCT : AX_ID)>(
        m, i, j, k, 1.0, DZB_EMIT(L, co), axis_flags(0), axis_flags(1), axis_flags(2));
  }
  template <int COMP, class E> __device__ void row(const TM &m, i64 i, i64 j, i64 k, E &&emit) const {
    if (AXIS < 0 || AXIS == 0) comp<0>(m, i, j, k, emit);
    if (AXIS < 0 || AXIS == 1) comp<1>(m, i, j, k, emit);
    if (AXIS < 0 || AXIS == 2) comp<2>(m, i, j, k, emit);
  }
};
template <> struct Rows<DZB_OP_CELL_GRAD, false> : RowsCellGrad<-1> {};
template <> struct Rows<DZB_OP_CELL_GRAD, true> : RowsCellGradT<-1> {};
template <> struct Rows<DZB_OP_CELL_GRAD_X, false> : RowsCellGrad<0> {};
template <> struct Rows<DZB_OP_CELL_GRAD_X, true> : RowsCellGradT<0> {};
template <> struct Rows<DZB_OP_CELL_GRAD_Y, false> : RowsCellGrad<1> {};
template <> struct Rows<DZB_OP_CELL_GRAD_Y, true> : RowsCellGradT<1> {};
template <> struct Rows<DZB_OP_CELL_GRAD_Z, false> : RowsCellGrad<2> {};
template <> struct Rows<DZB_OP_CELL_GRAD_Z, true> : RowsCellGradT<2> {};

// ---- face_x/y/z_divergence (:236-591): one component block of the divergence (nC x nF_axis) -----------
template <int AXIS> struct RowsDivComp : RowsBase {
  static constexpr int NCOMP = 1;
  static constexpr int MAXNNZ = 2;
  template <int COMP> static constexpr int loc() { return LOC_C; }
  template <int COMP> __host__ __device__ static i64 row_off(const TM &m) { return 0; }
  __host__ __device__ static i64 nrows(const TM &m) { return m.nC(); }
  __host__ __device__ static i64 ncols(const TM &m) { return m.count<FaceOf<AXIS>::LOC>(); }
  template <int COMP, class E> __device__ void row(const TM &m, i64 i, i64 j, i64 k, E &&emit) const {
    term<(AXIS == 0 ? AX_D : AX_ID), (AXIS == 1 ? AX_D : AX_ID), (AXIS == 2 ? AX_D : AX_ID)>(
        m, i, j, k, 1.0, DZB_EMIT(FaceOf<AXIS>::LOC, 0));
  }
};
template <int AXIS> struct RowsDivCompT : RowsBase {
  static constexpr int NCOMP = 1;
  static constexpr int MAXNNZ = 2;
  template <int COMP> static constexpr int loc() { return FaceOf<AXIS>::LOC; }
  template <int COMP> __host__ __device__ static i64 row_off(const TM &m) { return 0; }
  __host__ __device__ static i64 nrows(const TM &m) { return m.count<FaceOf<AXIS>::LOC>(); }
  __host__ __device__ static i64 ncols(const TM &m) { return m.nC(); }
  template <int COMP, class E> __device__ void row(const TM &m, i64 i, i64 j, i64 k, E &&emit) const {
    term<(AXIS == 0 ? AX_DT : AX_ID), (AXIS == 1 ? AX_DT : AX_ID), (AXIS == 2 ? AX_DT : AX_ID)>(m, i, j, k, 1.0,
                                                                                                 DZB_EMIT(LOC_C, 0));
  }
};
template <> struct Rows<DZB_OP_DIV_X, false> : RowsDivComp<0> {};
template <> struct Rows<DZB_OP_DIV_X, true> : RowsDivCompT<0> {};
template <> struct Rows<DZB_OP_DIV_Y, false> : RowsDivComp<1> {};
template <> struct Rows<DZB_OP_DIV_Y, true> : RowsDivCompT<1> {};
template <> struct Rows<DZB_OP_DIV_Z, false> : RowsDivComp<2> {};
template <> struct Rows<DZB_OP_DIV_Z, true> : RowsDivCompT<2> {};

// ---- entrywise squares of nodal_gradient and edge_curl --------------------------------------------------------------
// diag(G^T W G) = (G o G)^T w  and  diag(C^T W C) = (C o C)^T w  for diagonal W: the Jacobi preconditioners of the fused
// DC and Maxwell operators are ONE apply of these with the mass diagonal as input.
template <> struct Rows<DZB_OP_GRAD_SQ, false> : RowsBase {
  static constexpr int NCOMP = 3;
  static constexpr int MAXNNZ = 2;
  template <int COMP> static constexpr int loc() { return EdgeOf<COMP>::LOC; }
  template <int COMP> __host__ __device__ static i64 row_off(const TM &m) { return m.off<loc<COMP>()>(); }
  __host__ __device__ static i64 nrows(const TM &m) { return m.nE(); }
  __host__ __device__ static i64 ncols(const TM &m) { return m.nN(); }
  template <int COMP, class E> __device__ void row(const TM &m, i64 i, i64 j, i64 k, E &&emit) const {
    term<(COMP == 0 ? AX_D2 : AX_ID), (COMP == 1 ? AX_D2 : AX_ID), (COMP == 2 ? AX_D2 : AX_ID)>(m, i, j, k, 1.0,
                                                                                                 DZB_EMIT(LOC_N, 0));
  }
};
template <> struct Rows<DZB_OP_GRAD_SQ, true> : RowsBase {
  static constexpr int NCOMP = 1;
  static constexpr int MAXNNZ = 6;
  template <int COMP> static constexpr int loc() { return LOC_N; }
  template <int COMP> __host__ __device__ static i64 row_off(const TM &m) { return 0; }
  __host__ __device__ static i64 nrows(const TM &m) { return m.nN(); }
  __host__ __device__ static i64 ncols(const TM &m) { return m.nE(); }
  template <int COMP, class E> __device__ void row(const TM &m, i64 i, i64 j, i64 k, E &&emit) const {
    term<AX_DT2, AX_ID, AX_ID>(m, i, j, k, 1.0, DZB_EMIT(LOC_EX, m.off<LOC_EX>()));
    term<AX_ID, AX_DT2, AX_ID>(m, i, j, k, 1.0, DZB_EMIT(LOC_EY, m.off<LOC_EY>()));
    term<AX_ID, AX_ID, AX_DT2>(m, i, j, k, 1.0, DZB_EMIT(LOC_EZ, m.off<LOC_EZ>()));
  }
};
template <> struct Rows<DZB_OP_CURL_SQ, false> : RowsBase {
  static constexpr int NCOMP = 3;
  static constexpr int MAXNNZ = 4;
  template <int COMP> static constexpr int loc() { return FaceOf<COMP>::LOC; }
  template <int COMP> __host__ __device__ static i64 row_off(const TM &m) { return m.off<loc<COMP>()>(); }
  __host__ __device__ static i64 nrows(const TM &m) { return m.nF(); }
  __host__ __device__ static i64 ncols(const TM &m) { return m.nE(); }
  template <int COMP, class E> __device__ void row(const TM &m, i64 i, i64 j, i64 k, E &&emit) const {
    if (COMP == 0) {
      term<AX_ID, AX_ID, AX_D2>(m, i, j, k, 1.0, DZB_EMIT(LOC_EY, m.off<LOC_EY>()));
      term<AX_ID, AX_D2, AX_ID>(m, i, j, k, 1.0, DZB_EMIT(LOC_EZ, m.off<LOC_EZ>()));
    }
    if (COMP == 1) {
      term<AX_ID, AX_ID, AX_D2>(m, i, j, k, 1.0, DZB_EMIT(LOC_EX, m.off<LOC_EX>()));
      term<AX_D2, AX_ID, AX_ID>(m, i, j, k, 1.0, DZB_EMIT(LOC_EZ, m.off<LOC_EZ>()));
    }
    if (COMP == 2) {
      term<AX_ID, AX_D2, AX_ID>(m, i, j, k, 1.0, DZB_EMIT(LOC_EX, m.off<LOC_EX>()));
      term<AX_D2, AX_ID, AX_ID>(m, i, j, k, 1.0, DZB_EMIT(LOC_EY, m.off<LOC_EY>()));
    }
  }
};
template <> struct Rows<DZB_OP_CURL_SQ, true> : RowsBase {
  static constexpr int NCOMP = 3;
  static constexpr int MAXNNZ = 4;
  template <int COMP> static constexpr int loc() { return EdgeOf<COMP>::LOC; }
  template <int COMP> __host__ __device__ static i64 row_off(const TM &m) { return m.off<loc<COMP>()>(); }
  __host__ __device__ static i64 nrows(const TM &m) { return m.nE(); }
  __host__ __device__ static i64 ncols(const TM &m) { return m.nF(); }
  template <int COMP, class E> __device__ void row(const TM &m, i64 i, i64 j, i64 k, E &&emit) const {
    if (COMP == 0) {
      term<AX_ID, AX_ID, AX_DT2>(m, i, j, k, 1.0, DZB_EMIT(LOC_FY, m.off<LOC_FY>()));
      term<AX_ID, AX_DT2, AX_ID>(m, i, j, k, 1.0, DZB_EMIT(LOC_FZ, m.off<LOC_FZ>()));
    }
    if (COMP == 1) {
      term<AX_ID, AX_ID, AX_DT2>(m, i, j, k, 1.0, DZB_EMIT(LOC_FX, m.off<LOC_FX>()));
      term<AX_DT2, AX_ID, AX_ID>(m, i, j, k, 1.0, DZB_EMIT(LOC_FZ, m.off<LOC_FZ>()));
    }
    if (COMP == 2) {
      term<AX_ID, AX_DT2, AX_ID>(m, i, j, k, 1.0, DZB_EMIT(LOC_FX, m.off<LOC_FX>()));
      term<AX_DT2, AX_ID, AX_ID>(m, i, j, k, 1.0, DZB_EMIT(LOC_FY, m.off<LOC_FY>()));
    }
  }
};

}  // namespace dzb
