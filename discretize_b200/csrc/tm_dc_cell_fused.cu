// Fused cell-centred DC operator  y = D Mf(rho)^-1 D^T x  (SimPEG's default DC formulation; the reference builds it
// from face_divergence and get_face_inner_product(rho, invert_matrix=True): examples/plot_dc_resistivity.py:23-30,
// operators/differential_operators.py:225-235, operators/inner_products.py:35-99) as ONE plane-marching kernel.
//
// Unfused it is three applies -- D^T (4.3 GB at 512^3), the inverse diagonal mass matrix (9.7 GB), D (4.3 GB) -- and
// two face-sized temporaries; fused, the compulsory traffic is x, rho in and y out: 24 B per cell (3.2 GB).
//
//   flux through the face between cells L and R (along an axis with widths h):
//       g = x_L / h_L - x_R / h_R                     (D^T)
//       M = 1/2 (V_L rho_L + V_R rho_R)               (diagonal face mass matrix, V = (hx hy) hz)
//       F = g / M
//   boundary faces have one neighbour (Dirichlet ghost value 0) or, with DZB_DC_CELL_NEUMANN, carry no flux;
//   y_c = sum over axes of (F_hi - F_lo) / h_c         (D)
//
// A thread owns the (i, j) cell column and walks KC layers: x and rho of the layer above are loaded once and carried,
// the z-flux through the bottom face is the top flux of the previous step; the four lateral neighbours come from L1.
// Five FP64 divisions per cell (the lateral fluxes are evaluated by both neighbours -- cheaper than a shuffle /
// shared-memory exchange at this arithmetic intensity).
#include "tm_march.cuh"

namespace dzb {

template <bool NEUMANN, class CFG>
__global__ void DZB_MARCH_BOUNDS(CFG)
    k_dc_cell_march(TM m, const double *__restrict__ rho, const double *__restrict__ x, double *__restrict__ y) {
  MarchGeom g;
  const i64 nx = m.nx, ny = m.ny, nz = m.nz;
  if (!g.template init<CFG>(nx, ny, nz)) return;
  const i64 i = g.i, j = g.j;
  const bool lx = i >= 1, hx = i + 1 < nx, ly = j >= 1, hy = j + 1 < ny;
  const double hxc = __ldg(m.hx + i), hyc = __ldg(m.hy + j);
  const double hxl = lx ? __ldg(m.hx + i - 1) : 0.0, hxr = hx ? __ldg(m.hx + i + 1) : 0.0;
  const double hyl = ly ? __ldg(m.hy + j - 1) : 0.0, hyr = hy ? __ldg(m.hy + j + 1) : 0.0;
  const double rxc = __ldg(m.rhx + i), ryc = __ldg(m.rhy + j);
  const double rxl = lx ? __ldg(m.rhx + i - 1) : 0.0, rxr = hx ? __ldg(m.rhx + i + 1) : 0.0;
  const double ryl = ly ? __ldg(m.rhy + j - 1) : 0.0, ryr = hy ? __ldg(m.rhy + j + 1) : 0.0;
  // (hx * hy) of the five cell columns: centre, x-left, x-right, y-left, y-right
  const double ac = hxc * hyc, axl = hxl * hyc, axr = hxr * hyc, ayl = hxc * hyl, ayr = hxc * hyr;
  const i64 plC = nx * ny;
  const i64 off = i + nx * j + plC * g.k0;
  const double *px = x + off, *pr = rho + off;
  double *py = y + off;
  // flux between the low-side cell L (value xl, V rho = vl, 1/h = rl) and the high-side cell R.  A missing cell enters
  // with x = V rho = 1/h = 0, which IS the boundary formula (one-sided mass, ghost value 0): no branch needed.
  // `both` = the face is interior (Neumann drops the others).
  auto flux = [&](bool both, double xl, double vl, double rl, double xr, double vr, double rr) -> double {
    const double f = (xl * rl - xr * rr) / (0.5 * vl + 0.5 * vr);
    return (NEUMANN && !both) ? 0.0 : f;
  };
  // layer k0: centre values, and the flux through its bottom face
  double xc = __ldg(px), rc = __ldg(pr);
  double hzc = __ldg(m.hz + g.k0), rzc = __ldg(m.rhz + g.k0);
  double fz_lo;
  {
    const bool has = g.k0 >= 1;
    const double xb = has ? __ldg(px - plC) : 0.0, rb = has ? __ldg(pr - plC) : 0.0;
    const double hzb = has ? __ldg(m.hz + g.k0 - 1) : 0.0, rzb = has ? __ldg(m.rhz + g.k0 - 1) : 0.0;
    fz_lo = flux(has, xb, (ac * hzb) * rb, rzb, xc, (ac * hzc) * rc, rzc);
  }
#pragma unroll 2
  for (i64 k = g.k0; k < g.k1; ++k) {
    const bool ht = k + 1 < nz;
    const double xt = ht ? __ldg(px + plC) : 0.0, rt = ht ? __ldg(pr + plC) : 0.0;
    const double hzt = ht ? __ldg(m.hz + k + 1) : 0.0, rzt = ht ? __ldg(m.rhz + k + 1) : 0.0;
    const double xL = lx ? __ldg(px - 1) : 0.0, rL = lx ? __ldg(pr - 1) : 0.0;
    const double xR = hx ? __ldg(px + 1) : 0.0, rR = hx ? __ldg(pr + 1) : 0.0;
    const double xD = ly ? __ldg(px - nx) : 0.0, rD = ly ? __ldg(pr - nx) : 0.0;
    const double xU = hy ? __ldg(px + nx) : 0.0, rU = hy ? __ldg(pr + nx) : 0.0;
    const double vc = (ac * hzc) * rc;
    const double fx_lo = flux(lx, xL, (axl * hzc) * rL, rxl, xc, vc, rxc);
    const double fx_hi = flux(hx, xc, vc, rxc, xR, (axr * hzc) * rR, rxr);
    const double fy_lo = flux(ly, xD, (ayl * hzc) * rD, ryl, xc, vc, ryc);
    const double fy_hi = flux(hy, xc, vc, ryc, xU, (ayr * hzc) * rU, ryr);
    const double fz_hi = flux(ht, xc, vc, rzc, xt, (ac * hzt) * rt, rzt);
    *py = (fx_hi - fx_lo) * rxc + (fy_hi - fy_lo) * ryc + (fz_hi - fz_lo) * rzc;
    fz_lo = fz_hi; xc = xt; rc = rt; hzc = hzt; rzc = rzt;
    px += plC; pr += plC; py += plC;
  }
}

// The same operator with the inverse face mass diagonal precomputed (mi = 1 / diag(Mf(rho)), nF values: what
// get_face_inner_product(rho, invert_matrix=True) caches for its own applies anyway): no divisions, no neighbour rho --
// 40 B per cell (x, three face values, y) instead of 24 B + five FP64 divisions.
template <bool NEUMANN, class CFG>
__global__ void DZB_MARCH_BOUNDS(CFG)
    k_dc_cell_mi_march(TM m, const double *__restrict__ mi, const double *__restrict__ x, double *__restrict__ y) {
  MarchGeom g;
  const i64 nx = m.nx, ny = m.ny, nz = m.nz;
  if (!g.template init<CFG>(nx, ny, nz)) return;
  const i64 i = g.i, j = g.j;
  const bool lx = i >= 1, hx = i + 1 < nx, ly = j >= 1, hy = j + 1 < ny;
  const double rxc = __ldg(m.rhx + i), ryc = __ldg(m.rhy + j);
  const double rxl = lx ? __ldg(m.rhx + i - 1) : 0.0, rxr = hx ? __ldg(m.rhx + i + 1) : 0.0;
  const double ryl = ly ? __ldg(m.rhy + j - 1) : 0.0, ryr = hy ? __ldg(m.rhy + j + 1) : 0.0;
  // Neumann: boundary faces carry no flux -> their weight is zeroed once per thread
  const double wxl = (NEUMANN && !lx) ? 0.0 : 1.0, wxr = (NEUMANN && !hx) ? 0.0 : 1.0;
  const double wyl = (NEUMANN && !ly) ? 0.0 : 1.0, wyr = (NEUMANN && !hy) ? 0.0 : 1.0;
  const i64 plC = nx * ny, pFx = nx + 1, plFx = pFx * ny, plFy = nx * (ny + 1);
  const double *px = x + i + nx * j + plC * g.k0;
  const double *fx = mi + i + pFx * j + plFx * g.k0;                         // x-face (i, j, k); (i+1, j, k) = fx + 1
  const double *fy = mi + m.off<LOC_FY>() + i + nx * j + plFy * g.k0;        // y-face (i, j, k); (i, j+1, k) = fy + nx
  const double *fz = mi + m.off<LOC_FZ>() + i + nx * j + plC * g.k0;         // z-face (i, j, k); (i, j, k+1) = fz + plC
  double *py = y + i + nx * j + plC * g.k0;
  double xc = __ldg(px);
  double rzc = __ldg(m.rhz + g.k0);
  double fz_lo;
  {
    const bool has = g.k0 >= 1;
    const double xb = has ? __ldg(px - plC) : 0.0, rzb = has ? __ldg(m.rhz + g.k0 - 1) : 0.0;
    fz_lo = (NEUMANN && !has) ? 0.0 : (xb * rzb - xc * rzc) * __ldg(fz);
  }
  // layers with a layer above them run without any k-dependent predicate (the unrolled loads issue together); the
  // top layer of the mesh, if it is in this chunk, is one peeled step
  auto step = [&](auto top_tag, i64 k) {
    constexpr bool HT = decltype(top_tag)::value;     // a layer k + 1 exists
    double xt = 0.0, rzt = 0.0;
    if (HT) {
      xt = __ldg(px + plC);
      rzt = __ldg(m.rhz + k + 1);
    }
    const double xL = lx ? __ldg(px - 1) : 0.0, xR = hx ? __ldg(px + 1) : 0.0;
    const double xD = ly ? __ldg(px - nx) : 0.0, xU = hy ? __ldg(px + nx) : 0.0;
    const double fx_lo = wxl * ((xL * rxl - xc * rxc) * __ldg(fx)), fx_hi = wxr * ((xc * rxc - xR * rxr) * __ldg(fx + 1));
    const double fy_lo = wyl * ((xD * ryl - xc * ryc) * __ldg(fy)), fy_hi = wyr * ((xc * ryc - xU * ryr) * __ldg(fy + nx));
    const double fz_hi = (NEUMANN && !HT) ? 0.0 : (xc * rzc - xt * rzt) * __ldg(fz + plC);
    *py = (fx_hi - fx_lo) * rxc + (fy_hi - fy_lo) * ryc + (fz_hi - fz_lo) * rzc;
    fz_lo = fz_hi; xc = xt; rzc = rzt;
    px += plC; fx += plFx; fy += plFy; fz += plC; py += plC;
  };
  {
    i64 k = g.k0;
    const i64 kmain = min(g.k1, nz - 1);
#pragma unroll 4
    for (; k < kmain; ++k) step(std::true_type{}, k);
    if (k < g.k1) step(std::false_type{}, k);
  }
}

}  // namespace dzb

using namespace dzb;

extern "C" int dzb_tm_dc_cell_mi(const DzbTensorMesh *mm, const double *mi, int flags, const double *x, double *y,
                                 void *stream) {
  if (!valid_mesh(mm)) return -1;
  if (!mi) {
    set_error("dzb_tm_dc_cell_mi: the inverse face mass diagonal is null");
    return -1;
  }
  TM m(*mm);
  typedef MarchCfg<8, 16, 4> CFG;
  const dim3 gc = march_grid<CFG>(m.nx, m.ny, m.nz), b(32, CFG::BY, 1);
  if (flags & DZB_DC_CELL_NEUMANN) k_dc_cell_mi_march<true, CFG><<<gc, b, 0, (cudaStream_t)stream>>>(m, mi, x, y);
  else k_dc_cell_mi_march<false, CFG><<<gc, b, 0, (cudaStream_t)stream>>>(m, mi, x, y);
  return check_launch("dzb_tm_dc_cell_mi");
}


extern "C" int dzb_tm_dc_cell(const DzbTensorMesh *mm, const double *rho, int flags, const double *x, double *y,
                              void *stream) {
  if (!valid_mesh(mm)) return -1;
  if (!rho) {
    set_error("dzb_tm_dc_cell: rho is null");
    return -1;
  }
  TM m(*mm);
  typedef MarchCfg<8, 16, 0> CFG;
  const dim3 gc = march_grid<CFG>(m.nx, m.ny, m.nz), b(32, CFG::BY, 1);
  if (flags & DZB_DC_CELL_NEUMANN) k_dc_cell_march<true, CFG><<<gc, b, 0, (cudaStream_t)stream>>>(m, rho, x, y);
  else k_dc_cell_march<false, CFG><<<gc, b, 0, (cudaStream_t)stream>>>(m, rho, x, y);
  return check_launch("dzb_tm_dc_cell");
}
