// K9: y = C^T Mf(mui) C e + Me(sigma) e  (Maxwell, E-formulation) without temporaries.
//
// Reference composition (tests/boundaries/test_boundary_maxwell.py:95-99):
//   C  = edge_curl                      (operators/differential_operators.py:2169-2181)
//   Mf = get_face_inner_product(mui)    (operators/inner_products.py:35-67, isotropic)
//   Me = get_edge_inner_product(sigma)  (operators/inner_products.py:69-99, isotropic)
// A 13-point edge stencil: every edge couples to the edges of its (<=4) adjacent faces.
//
// Variant "gather": one thread per output edge evaluates the whole chain for its row
// from global memory (L1/L2 serve the re-reads).  It is the simple, obviously-correct
// form and the cross-check for the marching variant.
#include "tm_mass_common.cuh"
#include "tm_rowkernels.cuh"

namespace dzb {

// Mf(f) * (C e)(f) for the face of component FC at (a,b,c)
template <int FC>
__device__ __forceinline__ double face_flux(const TM &m, const Model &mu, const double *__restrict__ e, i64 a, i64 b,
                                            i64 c) {
  double ce = 0.0;
  Rows<DZB_OP_CURL, false>::row<FC>(m, a, b, c, [&](i64 col, double v) { ce += v * __ldg(e + col); });
  return mass_diag_entry<false, FC>(m, mu, a, b, c) * ce;
}

template <int COMP>
__global__ void __launch_bounds__(BX *BY *BZ)
    k_maxwell_gather(TM m, Model mu, Model sg, const double *__restrict__ e, double *__restrict__ y) {
  constexpr int LOC = DofLoc<true, COMP>::LOC;
  const i64 i = (i64)blockIdx.x * BX + threadIdx.x;
  const i64 j = (i64)blockIdx.y * BY + threadIdx.y;
  const i64 k = (i64)blockIdx.z * BZ + threadIdx.z;
  if (i >= m.sx<LOC>() || j >= m.sy<LOC>() || k >= m.sz<LOC>()) return;
  const i64 row = m.idx<LOC>(i, j, k);
  double acc = mass_diag_entry<true, COMP>(m, sg, i, j, k) * __ldg(e + row);
  // rows of C^T (cf. Rows<DZB_OP_CURL, true>)
  if (COMP == 0) {
    term<AX_ID, AX_ID, AX_DT>(m, i, j, k, 1.0, [&](i64 a, i64 b, i64 c, double w) { acc += w * face_flux<1>(m, mu, e, a, b, c); });
    term<AX_ID, AX_DT, AX_ID>(m, i, j, k, -1.0, [&](i64 a, i64 b, i64 c, double w) { acc += w * face_flux<2>(m, mu, e, a, b, c); });
  } else if (COMP == 1) {
    term<AX_ID, AX_ID, AX_DT>(m, i, j, k, -1.0, [&](i64 a, i64 b, i64 c, double w) { acc += w * face_flux<0>(m, mu, e, a, b, c); });
    term<AX_DT, AX_ID, AX_ID>(m, i, j, k, 1.0, [&](i64 a, i64 b, i64 c, double w) { acc += w * face_flux<2>(m, mu, e, a, b, c); });
  } else {
    term<AX_ID, AX_DT, AX_ID>(m, i, j, k, 1.0, [&](i64 a, i64 b, i64 c, double w) { acc += w * face_flux<0>(m, mu, e, a, b, c); });
    term<AX_DT, AX_ID, AX_ID>(m, i, j, k, -1.0, [&](i64 a, i64 b, i64 c, double w) { acc += w * face_flux<1>(m, mu, e, a, b, c); });
  }
  y[row] = acc;
}

}  // namespace dzb

using namespace dzb;

extern "C" int dzb_tm_maxwell(const DzbTensorMesh *mm, const double *mui, const double *sigma, const double *e,
                              double *y, void *stream) {
  if (!valid_mesh(mm)) return -1;
  TM m(*mm);
  Model mu{DZB_MODEL_ISO, mui, 0.0, m.nC(), false};
  Model sg{DZB_MODEL_ISO, sigma, 0.0, m.nC(), false};
  cudaStream_t s = (cudaStream_t)stream;
  dim3 b(BX, BY, BZ);
  k_maxwell_gather<0><<<grid_for<LOC_EX>(m), b, 0, s>>>(m, mu, sg, e, y);
  k_maxwell_gather<1><<<grid_for<LOC_EY>(m), b, 0, s>>>(m, mu, sg, e, y);
  k_maxwell_gather<2><<<grid_for<LOC_EZ>(m), b, 0, s>>>(m, mu, sg, e, y);
  return check_launch("dzb_tm_maxwell");
}
