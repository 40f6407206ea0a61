// K5, plane-marching form: diagonal mass matrices (scalar / isotropic / anisotropic model) for faces and edges,
//   out = diag(M)  (MODE 0)   or   out = M x  (MODE 1),   optionally with 1/model and 1/M
// reference: base/base_tensor_mesh.py:792-863 (_fastInnerProduct): face entry = 1/2 sum of V sigma over the
// (<= 2) adjacent cells, edge entry = 1/4 sum over the (<= 4) adjacent cells, component c uses sigma_c.
//
// The generic kernel (tm_mass.cu: k_mass_diag) recomputes the cell volume and the 64-bit cell index for every
// (dof, adjacent cell) pair: 2.47e9 warp instructions at 512^3 -- issue-bound at 2.5 TB/s.  Here a thread owns
// the (i, j) column: the x-y part of the volume of its four neighbouring cell columns is computed once, cell
// values of the layer below are carried in registers, pointers advance by constant plane strides.
#pragma once
#include "tm_march.cuh"
#include "tm_mass_common.cuh"

namespace dzb {

// INV: bit 0 = invert_model, bit 1 = invert_matrix -- compile-time, so that the plain case has no branch (and no
// division slow path) between its loads and stores
template <bool EDGE, int MODE, int KIND, int INV, class CFG>
__global__ void DZB_MARCH_BOUNDS(CFG)
    k_mass_diag_march(TM m, Model md, const double *__restrict__ x, double *__restrict__ out) {
  MarchGeom g;
  const i64 nx = m.nx, ny = m.ny, nz = m.nz;
  if (!g.template init<CFG>(nx + 1, ny + 1, nz + 1)) return;
  const i64 i = g.i, j = g.j;
  const bool hx = i < nx, hy = j < ny, lx = i >= 1, ly = j >= 1;
  // cell columns around node column (i, j): A = (i, j), B = (i-1, j), C = (i, j-1), D = (i-1, j-1)
  const bool okA = hx && hy, okB = lx && hy, okC = hx && ly, okD = lx && ly;
  const double hxi = hx ? __ldg(m.hx + i) : 0.0, hxm = lx ? __ldg(m.hx + i - 1) : 0.0;
  const double hyj = hy ? __ldg(m.hy + j) : 0.0, hym = ly ? __ldg(m.hy + j - 1) : 0.0;
  const double aA = hxi * hyj, aB = hxm * hyj, aC = hxi * hym, aD = hxm * hym;   // (hx * hy) of tensor_mesh.py:296-310
  const i64 pN = nx + 1, plC = nx * ny, nC = md.nC;
  const i64 plX = EDGE ? nx * (ny + 1) : pN * ny;        // plane stride of the x-component (Ex | Fx)
  const i64 plY = EDGE ? pN * ny : nx * (ny + 1);        // Ey | Fy
  const i64 plZ = EDGE ? pN * (ny + 1) : plC;            // Ez | Fz
  const i64 rowX = EDGE ? nx : pN, rowY = EDGE ? pN : nx, rowZ = EDGE ? pN : nx;
  const i64 offY = EDGE ? m.off<LOC_EY>() : m.off<LOC_FY>(), offZ = EDGE ? m.off<LOC_EZ>() : m.off<LOC_FZ>();
  i64 qx = i + rowX * j + plX * g.k0, qy = offY + i + rowY * j + plY * g.k0, qz = offZ + i + rowZ * j + plZ * g.k0;
  const double *sA = md.p + i + nx * j + plC * g.k0;     // model value of cell A in layer k (component 0)
  const i64 c1 = KIND == DZB_MODEL_ANISO ? nC : 0, c2 = KIND == DZB_MODEL_ANISO ? 2 * nC : 0;
  constexpr bool inv = (INV & 1) != 0, inv_mat = (INV & 2) != 0;
  const double sc = md.scalar;
  // (possibly inverted) model value; `ok` guards the load
  auto val = [&](const double *p, bool ok) -> double {
    if (!ok) return 0.0;
    const double v = KIND == DZB_MODEL_SCALAR ? sc : __ldg(p);
    return inv ? 1.0 / v : v;
  };
  auto put = [&](i64 q, double d) {
    if (inv_mat) d = 1.0 / d;
    out[q] = MODE == 0 ? d : d * __ldg(x + q);
  };
  // carried from the layer below: faces: V sigma_z of cell A; edges: the x- and y-component pair sums
  double lo0 = 0.0, lo1 = 0.0;
  if (g.k0 >= 1) {
    const double hz = __ldg(m.hz + g.k0 - 1);
    const double *sL = sA - plC;
    if (!EDGE) {
      lo0 = (aA * hz) * val(sL + c2, okA);
    } else {
      lo0 = (aC * hz) * val(sL - nx, okC) + (aA * hz) * val(sL, okA);              // x-edges: cells C, A
      lo1 = (aB * hz) * val(sL - 1 + c1, okB) + (aA * hz) * val(sL + c1, okA);     // y-edges: cells B, A
    }
  }
  auto step = [&](auto hz_tag, i64 k) {
    constexpr bool HZ = decltype(hz_tag)::value;
    double hi0 = 0.0, hi1 = 0.0;
    if (!EDGE) {
      if (HZ) {
        const double hz = __ldg(m.hz + k);
        const double vA = aA * hz, vB = aB * hz, vC = aC * hz;
        const double ax = val(sA, okA), bx = val(sA - 1, okB);
        const double ay = KIND == DZB_MODEL_ANISO ? val(sA + c1, okA) : ax, cy = val(sA - nx + c1, okC);
        const double az = KIND == DZB_MODEL_ANISO ? val(sA + c2, okA) : ax;
        if (hy) put(qx, 0.5 * (vB * bx) + 0.5 * (vA * ax));
        if (hx) put(qy, 0.5 * (vC * cy) + 0.5 * (vA * ay));
        hi0 = vA * az;
      }
      if (okA) put(qz, 0.5 * lo0 + 0.5 * hi0);
    } else {
      double ez = 0.0;
      if (HZ) {
        const double hz = __ldg(m.hz + k);
        const double vA = aA * hz, vB = aB * hz, vC = aC * hz, vD = aD * hz;
        const double ax = val(sA, okA), cx = val(sA - nx, okC);
        const double ay = KIND == DZB_MODEL_ANISO ? val(sA + c1, okA) : ax, by = val(sA - 1 + c1, okB);
        const double az = KIND == DZB_MODEL_ANISO ? val(sA + c2, okA) : ax;
        const double bz = KIND == DZB_MODEL_ANISO ? val(sA - 1 + c2, okB) : by;
        const double cz = KIND == DZB_MODEL_ANISO ? val(sA - nx + c2, okC) : cx;
        const double dz = val(sA - nx - 1 + c2, okD);
        hi0 = vC * cx + vA * ax;
        hi1 = vB * by + vA * ay;
        ez = 0.25 * (((vD * dz + vC * cz) + vB * bz) + vA * az);
        put(qz, ez);
      }
      if (hx) put(qx, 0.25 * (lo0 + hi0));
      if (hy) put(qy, 0.25 * (lo1 + hi1));
    }
    lo0 = hi0; lo1 = hi1;
    sA += plC; qx += plX; qy += plY; qz += plZ;
  };
  DZB_MARCH_LOOP(step)
}

template <bool EDGE, int MODE, int INV, class CFG>
static void launch_mass_diag_march_cfg(const TM &m, const Model &md, const double *x, double *out, cudaStream_t s) {
  const dim3 gn = march_grid<CFG>(m.nx + 1, m.ny + 1, m.nz + 1), b(32, CFG::BY, 1);
  if (md.kind == DZB_MODEL_SCALAR) k_mass_diag_march<EDGE, MODE, DZB_MODEL_SCALAR, INV, CFG><<<gn, b, 0, s>>>(m, md, x, out);
  else if (md.kind == DZB_MODEL_ISO) k_mass_diag_march<EDGE, MODE, DZB_MODEL_ISO, INV, CFG><<<gn, b, 0, s>>>(m, md, x, out);
  else k_mass_diag_march<EDGE, MODE, DZB_MODEL_ANISO, INV, CFG><<<gn, b, 0, s>>>(m, md, x, out);
}

extern int g_mass_march;

template <bool EDGE, int MODE, int INV>
static void launch_mass_diag_march_inv(const TM &m, const Model &md, const double *x, double *out, cudaStream_t s) {
  switch (g_mass_march) {
#ifdef DZB_APPLY_SWEEP
    case 2: launch_mass_diag_march_cfg<EDGE, MODE, INV, MarchCfg<8, 16, 4>>(m, md, x, out, s); break;
    case 3: launch_mass_diag_march_cfg<EDGE, MODE, INV, MarchCfg<4, 32, 0>>(m, md, x, out, s); break;
    case 4: launch_mass_diag_march_cfg<EDGE, MODE, INV, MarchCfg<8, 32, 0>>(m, md, x, out, s); break;
    case 5: launch_mass_diag_march_cfg<EDGE, MODE, INV, MarchCfg<8, 16, 5>>(m, md, x, out, s); break;
    case 6: launch_mass_diag_march_cfg<EDGE, MODE, INV, MarchCfg<8, 8, 0>>(m, md, x, out, s); break;
    case 7: launch_mass_diag_march_cfg<EDGE, MODE, INV, MarchCfg<8, 16, 3>>(m, md, x, out, s); break;
    case 8: launch_mass_diag_march_cfg<EDGE, MODE, INV, MarchCfg<8, 16, 0>>(m, md, x, out, s); break;
#endif
    default:
      // 512^3 sweep (profiles/r1b_sweep_mass_march.log): plain iso / scalar models stream best with 4 x 32 blocks
      // over 32 planes; the anisotropic kernels (5-9 model loads per step) want ~80 registers; the variants
      // with divisions (1/model, 1/M) do best capped at 48 registers
      if (INV != 0) launch_mass_diag_march_cfg<EDGE, MODE, INV, MarchCfg<8, 16, 5>>(m, md, x, out, s);
      else if (md.kind == DZB_MODEL_ANISO) launch_mass_diag_march_cfg<EDGE, MODE, INV, MarchCfg<8, 16, 3>>(m, md, x, out, s);
      else launch_mass_diag_march_cfg<EDGE, MODE, INV, MarchCfg<4, 32, 0>>(m, md, x, out, s);
      break;
  }
}

template <bool EDGE, int MODE>
static void launch_mass_diag_march(const TM &m, const Model &md, bool inv_mat, const double *x, double *out, cudaStream_t s) {
  switch ((md.inv ? 1 : 0) | (inv_mat ? 2 : 0)) {
    case 0: launch_mass_diag_march_inv<EDGE, MODE, 0>(m, md, x, out, s); break;
    case 1: launch_mass_diag_march_inv<EDGE, MODE, 1>(m, md, x, out, s); break;
    case 2: launch_mass_diag_march_inv<EDGE, MODE, 2>(m, md, x, out, s); break;
    default: launch_mass_diag_march_inv<EDGE, MODE, 3>(m, md, x, out, s); break;
  }
}

}  // namespace dzb

namespace dzb {

// K7 adjoint without inversions, plane-marching: y[cell] (+ comp * nC) = V_cell * sum over the cell's faces (x 1/2) or
// edges (x 1/4) of v[f] w[f]  (base/base_tensor_mesh.py:865-988: F(v)^T w = (k Av^T diag(V))^T (v o w));
// isotropic models sum the three components, anisotropic ones keep them apart.  The forward product without
// inversions IS the diagonal mass kernel with the model replaced by dm (F(v) dm = v o diag(M(dm))).
template <bool EDGE, bool ANISO, class CFG>
__global__ void DZB_MARCH_BOUNDS(CFG)
    k_deriv_adj_march(TM m, i64 nC, const double *__restrict__ v, const double *__restrict__ w, double *__restrict__ y) {
  MarchGeom g;
  const i64 nx = m.nx, ny = m.ny, nz = m.nz;
  if (!g.template init<CFG>(nx, ny, nz)) return;
  const i64 i = g.i, j = g.j;
  const i64 pN = nx + 1, plC = nx * ny;
  const double axy = __ldg(m.hx + i) * __ldg(m.hy + j);
  double *o = y + i + nx * j + plC * g.k0;
  auto pr = [&](i64 q) { return __ldg(v + q) * __ldg(w + q); };
  if (!EDGE) {
    const i64 plFx = pN * ny, plFy = nx * (ny + 1);
    i64 qx = i + pN * j + plFx * g.k0, qy = m.off<LOC_FY>() + i + nx * j + plFy * g.k0,
        qz = m.off<LOC_FZ>() + i + nx * j + plC * g.k0;
    double z0 = pr(qz);
#pragma unroll 4
    for (i64 k = g.k0; k < g.k1; ++k) {
      const double z1 = pr(qz + plC);
      const double sx = 0.5 * pr(qx) + 0.5 * pr(qx + 1), sy = 0.5 * pr(qy) + 0.5 * pr(qy + nx), sz = 0.5 * z0 + 0.5 * z1;
      const double V = axy * __ldg(m.hz + k);
      if (ANISO) {
        o[0] = V * sx; o[nC] = V * sy; o[2 * nC] = V * sz;
      } else {
        o[0] = V * ((sx + sy) + sz);
      }
      z0 = z1;
      qx += plFx; qy += plFy; qz += plC; o += plC;
    }
  } else {
    const i64 plN = pN * (ny + 1), plEx = nx * (ny + 1), plEy = pN * ny;
    i64 qx = i + nx * j + plEx * g.k0, qy = m.off<LOC_EY>() + i + pN * j + plEy * g.k0,
        qz = m.off<LOC_EZ>() + i + pN * j + plN * g.k0;
    double sx0 = pr(qx) + pr(qx + nx), sy0 = pr(qy) + pr(qy + 1);
#pragma unroll 2
    for (i64 k = g.k0; k < g.k1; ++k) {
      const double sx1 = pr(qx + plEx) + pr(qx + plEx + nx), sy1 = pr(qy + plEy) + pr(qy + plEy + 1);
      const double sz = (pr(qz) + pr(qz + 1)) + (pr(qz + pN) + pr(qz + pN + 1));
      const double V = axy * __ldg(m.hz + k);
      const double ex = 0.25 * (sx0 + sx1), ey = 0.25 * (sy0 + sy1), ez = 0.25 * sz;
      if (ANISO) {
        o[0] = V * ex; o[nC] = V * ey; o[2 * nC] = V * ez;
      } else {
        o[0] = V * ((ex + ey) + ez);
      }
      sx0 = sx1; sy0 = sy1;
      qx += plEx; qy += plEy; qz += plN; o += plC;
    }
  }
}

template <bool EDGE>
static void launch_deriv_adj_march(const TM &m, const Model &md, const double *v, const double *w, double *y, cudaStream_t s) {
  typedef MarchCfg<8, 16, 0> CFG;
  const dim3 gc = march_grid<CFG>(m.nx, m.ny, m.nz), b(32, CFG::BY, 1);
  if (md.kind == DZB_MODEL_ANISO) k_deriv_adj_march<EDGE, true, CFG><<<gc, b, 0, s>>>(m, md.nC, v, w, y);
  else k_deriv_adj_march<EDGE, false, CFG><<<gc, b, 0, s>>>(m, md.nC, v, w, y);
}

}  // namespace dzb
