// K6, plane-marching form: y = M(Sigma) x for a FULL-TENSOR cell property (6 components per cell), faces and edges.
// reference: operators/inner_products.py:147-272, 671-738 -- M = sum over the 8 corners n of every cell of
//   P_n^T (V/8 Sigma) P_n,   P_n = the three faces (edges) that meet at corner n.
//
// Summed over the corners of ONE cell (v8 = V/8; s = Sigma of the cell):
//   faces  y_fx[c0]     += v8 ( 4 sxx fx[c0]     + 2 sxy (fy[0] + fy[1])         + 2 sxz (fz[0] + fz[1]) )
//   edges  y_ex[c1][c2] += v8 ( 2 sxx ex[c1][c2] +   sxy (ey[0][c2] + ey[1][c2]) +   sxz (ez[0][c1] + ez[1][c1]) )
// and cyclically (fx[c0] = the cell's x-face on side c0; ex[c1][c2] = its x-edge at corner (c1, c2) of the y-z
// section, ey[c0][c2], ez[c0][c1] likewise), i.e. the rows the generic generator RowsMassTensor (tm_mass.cu) emits, term
// for term.  That generator spends ~500 instructions per row on 64-bit index arithmetic (issue-bound: 0.24 / 0.14 of
// HBM at 512^3); here a thread owns the node column (i, j), walks the planes with pointers that advance by constant
// strides, reads its x / y neighbours directly (L1 hits: every value comes from DRAM once) and carries what the next
// plane needs in registers: the z-face (z-edge) part of the layer below and the contribution of the layer below to the
// plane it shares with the current one.
#pragma once
#include "tm_march.cuh"
#include "tm_mass_common.cuh"

namespace dzb {

// like DZB_MARCH_LOOP without unrolling: one step of these kernels already has 25-45 loads in flight, and four unrolled
// steps need 164 (faces) / 255 + spills (edges) registers
#define DZB_MARCH_LOOP_1(STEP)                                \
  {                                                           \
    i64 k = g.k0;                                             \
    const i64 kmain = min(g.k1, nz);                          \
    _Pragma("unroll 1") for (; k < kmain; ++k) STEP(std::true_type{}, k);  \
    if (k < g.k1) STEP(std::false_type{}, k);                 \
  }

// ---- faces -----------------------------------------------------------------------------------------------------------
template <class CFG>
__global__ void DZB_MARCH_BOUNDS(CFG)
    k_mass_tensor_face_march(TM m, const double *__restrict__ sg, i64 nC, const double *__restrict__ x, double *__restrict__ y) {
  MarchGeom g;
  const i64 nx = m.nx, ny = m.ny, nz = m.nz;
  if (!g.template init<CFG>(nx + 1, ny + 1, nz + 1)) return;
  const i64 i = g.i, j = g.j;
  const bool hx = i < nx, hy = j < ny, lx = i >= 1, ly = j >= 1;
  // cell columns around node column (i, j): A = (i, j), B = (i-1, j), C = (i, j-1)
  const bool okA = hx && hy, okB = lx && hy, okC = hx && ly;
  const double hxi = hx ? __ldg(m.hx + i) : 0.0, hxm = lx ? __ldg(m.hx + i - 1) : 0.0;
  const double hyj = hy ? __ldg(m.hy + j) : 0.0, hym = ly ? __ldg(m.hy + j - 1) : 0.0;
  const double aA = okA ? hxi * hyj : 0.0, aB = okB ? hxm * hyj : 0.0, aC = okC ? hxi * hym : 0.0;
  const i64 pN = nx + 1, plC = nx * ny, plX = pN * ny, plY = nx * (ny + 1);
  const i64 offY = m.off<LOC_FY>(), offZ = m.off<LOC_FZ>();
  const i64 qx0 = i + pN * j + plX * g.k0, qy0 = offY + i + nx * j + plY * g.k0, qz0 = offZ + i + nx * j + plC * g.k0;
  const double *fx = x + qx0, *fy = x + qy0, *fz = x + qz0;
  double *ox = y + qx0, *oy = y + qy0, *oz = y + qz0;
  const double *sA = sg + i + nx * j + plC * g.k0;   // component xx of cell A in layer k; yy, zz, xy, xz, yz follow at nC strides
  const i64 XX = 0, YY = nC, ZZ = 2 * nC, XY = 3 * nC, XZ = 4 * nC, YZ = 5 * nC;

  // z-faces on the current plane (the upper ones of a layer become the lower ones of the next)
  double zA = ldp(fz, okA), zB = ldp(fz - 1, okB), zC = ldp(fz - nx, okC);
  // carried from the layer below: its contribution to the z-face between the layers = loc + lod * fz
  double loc = 0.0, lod = 0.0;
  if (g.k0 >= 1 && okA) {
    const double v8 = 0.125 * (aA * __ldg(m.hz + g.k0 - 1));
    const double *sL = sA - plC;
    const double ax = __ldg(fx - plX) + __ldg(fx - plX + 1), ay = __ldg(fy - plY) + __ldg(fy - plY + nx);
    loc = v8 * (2.0 * __ldg(sL + XZ) * ax + 2.0 * __ldg(sL + YZ) * ay);
    lod = 4.0 * v8 * __ldg(sL + ZZ);
  }
  auto step = [&](auto hz_tag, i64 k) {
    constexpr bool HZ = decltype(hz_tag)::value;
    double hic = 0.0, hid = 0.0;
    if (HZ) {
      const double hz = __ldg(m.hz + k);
      const double vA = 0.125 * (aA * hz), vB = 0.125 * (aB * hz), vC = 0.125 * (aC * hz);
      const double x0 = ldp(fx, hy), xA1 = ldp(fx + 1, okA), xC0 = ldp(fx - pN, ly), xC1 = ldp(fx - pN + 1, okC);
      const double y0 = ldp(fy, hx), yA1 = ldp(fy + nx, okA), yB0 = ldp(fy - 1, lx), yB1 = ldp(fy - 1 + nx, okB);
      const double zA1 = ldp(fz + plC, okA), zB1 = ldp(fz + plC - 1, okB), zC1 = ldp(fz + plC - nx, okC);
      const double axA = x0 + xA1, axC = xC0 + xC1, ayA = y0 + yA1, ayB = yB0 + yB1;
      const double azA = zA + zA1, azB = zB + zB1, azC = zC + zC1;
      const double sxxA = ldp(sA + XX, okA), syyA = ldp(sA + YY, okA), szzA = ldp(sA + ZZ, okA);
      const double sxyA = ldp(sA + XY, okA), sxzA = ldp(sA + XZ, okA), syzA = ldp(sA + YZ, okA);
      const double sxxB = ldp(sA - 1 + XX, okB), sxyB = ldp(sA - 1 + XY, okB), sxzB = ldp(sA - 1 + XZ, okB);
      const double syyC = ldp(sA - nx + YY, okC), sxyC = ldp(sA - nx + XY, okC), syzC = ldp(sA - nx + YZ, okC);
      if (hy) *ox = vB * (4.0 * sxxB * x0 + 2.0 * sxyB * ayB + 2.0 * sxzB * azB) +
                    vA * (4.0 * sxxA * x0 + 2.0 * sxyA * ayA + 2.0 * sxzA * azA);
      if (hx) *oy = vC * (4.0 * syyC * y0 + 2.0 * sxyC * axC + 2.0 * syzC * azC) +
                    vA * (4.0 * syyA * y0 + 2.0 * sxyA * axA + 2.0 * syzA * azA);
      hic = vA * (2.0 * sxzA * axA + 2.0 * syzA * ayA);
      hid = 4.0 * vA * szzA;
      if (okA) *oz = (loc + lod * zA) + (hic + hid * zA);
      zA = zA1; zB = zB1; zC = zC1;
    } else {
      if (okA) *oz = loc + lod * zA;
    }
    loc = hic; lod = hid;
    sA += plC; fx += plX; fy += plY; fz += plC; ox += plX; oy += plY; oz += plC;
  };
  DZB_MARCH_LOOP_1(step)
}

// ---- edges -----------------------------------------------------------------------------------------------------------
// Edge values a thread needs on one node plane: ex at (i, j), (i, j+1), (i-1, j), (i-1, j+1); ey at (i, j), (i+1, j),
// (i, j-1), (i+1, j-1).
struct EdgePlane {
  double x00, x01, xm0, xm1, y00, y10, y0m, y1m;
};

template <class CFG>
__global__ void DZB_MARCH_BOUNDS(CFG)
    k_mass_tensor_edge_march(TM m, const double *__restrict__ sg, i64 nC, const double *__restrict__ x, double *__restrict__ y) {
  MarchGeom g;
  const i64 nx = m.nx, ny = m.ny, nz = m.nz;
  if (!g.template init<CFG>(nx + 1, ny + 1, nz + 1)) return;
  const i64 i = g.i, j = g.j;
  const bool hx = i < nx, hy = j < ny, lx = i >= 1, ly = j >= 1;
  // cell columns around node column (i, j): A = (i, j), B = (i-1, j), C = (i, j-1), D = (i-1, j-1)
  const bool okA = hx && hy, okB = lx && hy, okC = hx && ly, okD = lx && ly;
  const double hxi = hx ? __ldg(m.hx + i) : 0.0, hxm = lx ? __ldg(m.hx + i - 1) : 0.0;
  const double hyj = hy ? __ldg(m.hy + j) : 0.0, hym = ly ? __ldg(m.hy + j - 1) : 0.0;
  const double aA = okA ? hxi * hyj : 0.0, aB = okB ? hxm * hyj : 0.0, aC = okC ? hxi * hym : 0.0, aD = okD ? hxm * hym : 0.0;
  const i64 pN = nx + 1, plC = nx * ny, plN = pN * (ny + 1), plX = nx * (ny + 1), plY = pN * ny;
  const i64 offY = m.off<LOC_EY>(), offZ = m.off<LOC_EZ>();
  const i64 qx0 = i + nx * j + plX * g.k0, qy0 = offY + i + pN * j + plY * g.k0, qz0 = offZ + i + pN * j + plN * g.k0;
  const double *ex = x + qx0, *ey = x + qy0, *ez = x + qz0;
  double *ox = y + qx0, *oy = y + qy0, *oz = y + qz0;
  const double *sA = sg + i + nx * j + plC * g.k0;
  const i64 XX = 0, YY = nC, ZZ = 2 * nC, XY = 3 * nC, XZ = 4 * nC, YZ = 5 * nC;

  auto load_plane = [&](const double *px, const double *py) {
    EdgePlane p;
    p.x00 = ldp(px, hx); p.x01 = ldp(px + nx, okA); p.xm0 = ldp(px - 1, lx); p.xm1 = ldp(px - 1 + nx, okB);
    p.y00 = ldp(py, hy); p.y10 = ldp(py + 1, okA); p.y0m = ldp(py - pN, ly); p.y1m = ldp(py - pN + 1, okC);
    return p;
  };
  // what the cells of ONE layer (model pointer s, thickness hz, z-edges pz) add to ex / ey on a node plane P that bounds
  // the layer (both planes get the same expression with their own edge values)
  struct LayerCoef {
    double xxA, xyA, xzA, xxC, xyC, xzC;   // x-edges: cells A and C, premultiplied by V/8
    double yyA, yzA, yyB, xyB, yzB;        // y-edges: cells A (xy shared with the x part) and B
    double zx, zy;                         // z-edge sums of the layer: ez(i,j) + ez(i+1,j), ez(i,j) + ez(i,j+1)
  };
  auto to_ex = [&](const LayerCoef &c, const EdgePlane &p) {
    return c.xxA * (2.0 * p.x00) + c.xyA * (p.y00 + p.y10) + c.xzA * c.zx +
           c.xxC * (2.0 * p.x00) + c.xyC * (p.y0m + p.y1m) + c.xzC * c.zx;
  };
  auto to_ey = [&](const LayerCoef &c, const EdgePlane &p) {
    return c.yyA * (2.0 * p.y00) + c.xyA * (p.x00 + p.x01) + c.yzA * c.zy +
           c.yyB * (2.0 * p.y00) + c.xyB * (p.xm0 + p.xm1) + c.yzB * c.zy;
  };
  auto coef = [&](const double *s, double hz, const double *pz, double &z00) {
    LayerCoef c;
    const double vA = 0.125 * (aA * hz), vB = 0.125 * (aB * hz), vC = 0.125 * (aC * hz);
    z00 = __ldg(pz);
    c.zx = z00 + ldp(pz + 1, hx);
    c.zy = z00 + ldp(pz + pN, hy);
    c.xxA = vA * ldp(s + XX, okA); c.xyA = vA * ldp(s + XY, okA); c.xzA = vA * ldp(s + XZ, okA);
    c.yyA = vA * ldp(s + YY, okA); c.yzA = vA * ldp(s + YZ, okA);
    c.xxC = vC * ldp(s - nx + XX, okC); c.xyC = vC * ldp(s - nx + XY, okC); c.xzC = vC * ldp(s - nx + XZ, okC);
    c.yyB = vB * ldp(s - 1 + YY, okB); c.xyB = vB * ldp(s - 1 + XY, okB); c.yzB = vB * ldp(s - 1 + YZ, okB);
    return c;
  };

  EdgePlane lo = load_plane(ex, ey);     // node plane k
  double cx = 0.0, cy = 0.0;             // contribution of the layer below to ex / ey on plane k
  if (g.k0 >= 1) {
    double z00;
    const LayerCoef c = coef(sA - plC, __ldg(m.hz + g.k0 - 1), ez - plN, z00);
    cx = to_ex(c, lo);
    cy = to_ey(c, lo);
  }
  auto step = [&](auto hz_tag, i64 k) {
    constexpr bool HZ = decltype(hz_tag)::value;
    if (HZ) {
      const double hz = __ldg(m.hz + k);
      double z00;
      const LayerCoef c = coef(sA, hz, ez, z00);
      const EdgePlane hi = load_plane(ex + plX, ey + plY);   // node plane k + 1
      if (hx) *ox = cx + to_ex(c, lo);
      if (hy) *oy = cy + to_ey(c, lo);
      cx = to_ex(c, hi);
      cy = to_ey(c, hi);
      // z-edge (i, j) of layer k: cells A, B, C, D
      const double vA = 0.125 * (aA * hz), vB = 0.125 * (aB * hz), vC = 0.125 * (aC * hz), vD = 0.125 * (aD * hz);
      const double zzA = ldp(sA + ZZ, okA), zzB = ldp(sA - 1 + ZZ, okB), zzC = ldp(sA - nx + ZZ, okC), zzD = ldp(sA - nx - 1 + ZZ, okD);
      const double xzB = ldp(sA - 1 + XZ, okB), yzC = ldp(sA - nx + YZ, okC);
      const double xzD = ldp(sA - nx - 1 + XZ, okD), yzD = ldp(sA - nx - 1 + YZ, okD);
      const double ex0 = lo.x00 + hi.x00, exm = lo.xm0 + hi.xm0, ey0 = lo.y00 + hi.y00, eym = lo.y0m + hi.y0m;
      const double two_z = 2.0 * z00;
      *oz = (c.xzA * ex0 + c.yzA * ey0 + vA * zzA * two_z) + (vB * (xzB * exm) + c.yzB * ey0 + vB * zzB * two_z) +
            (c.xzC * ex0 + vC * (yzC * eym) + vC * zzC * two_z) + (vD * (xzD * exm + yzD * eym) + vD * zzD * two_z);
      lo = hi;
    } else {
      if (hx) *ox = cx;
      if (hy) *oy = cy;
    }
    sA += plC; ex += plX; ey += plY; ez += plN; ox += plX; oy += plY; oz += plN;
  };
  DZB_MARCH_LOOP_1(step)
}

// sigma^-1 of every cell (invert_model for full tensors), with the closed form of Model::tensor so that the marching
// kernels and the generic generator see the same numbers
__global__ void k_tensor_invert(i64 nC, const double *__restrict__ sg, double *__restrict__ out) {
  const i64 stride = (i64)gridDim.x * blockDim.x;
  Model md{DZB_MODEL_TENSOR, sg, 0.0, nC, true};
  for (i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x; c < nC; c += stride) {
    double s[6];
    md.tensor(c, s);
#pragma unroll
    for (int q = 0; q < 6; ++q) out[q * nC + c] = s[q];
  }
}

extern int g_mass_tensor_march;   // 0: generic generator; 1: default configuration; > 1: sweep configurations

template <bool EDGE, class CFG>
static void launch_mass_tensor_march_cfg(const TM &m, const double *sg, const double *x, double *y, cudaStream_t s) {
  const dim3 gn = march_grid<CFG>(m.nx + 1, m.ny + 1, m.nz + 1), b(32, CFG::BY, 1);
  if (EDGE) k_mass_tensor_edge_march<CFG><<<gn, b, 0, s>>>(m, sg, m.nC(), x, y);
  else k_mass_tensor_face_march<CFG><<<gn, b, 0, s>>>(m, sg, m.nC(), x, y);
}

template <bool EDGE>
static void launch_mass_tensor_march(const TM &m, const double *sg, const double *x, double *y, cudaStream_t s) {
  switch (g_mass_tensor_march) {
    case 2: launch_mass_tensor_march_cfg<EDGE, MarchCfg<8, 16, 2>>(m, sg, x, y, s); break;
    case 3: launch_mass_tensor_march_cfg<EDGE, MarchCfg<8, 16, 3>>(m, sg, x, y, s); break;
    case 4: launch_mass_tensor_march_cfg<EDGE, MarchCfg<4, 32, 4>>(m, sg, x, y, s); break;
    case 5: launch_mass_tensor_march_cfg<EDGE, MarchCfg<8, 16, 0>>(m, sg, x, y, s); break;
    case 6: launch_mass_tensor_march_cfg<EDGE, MarchCfg<4, 16, 0>>(m, sg, x, y, s); break;
    case 7: launch_mass_tensor_march_cfg<EDGE, MarchCfg<8, 32, 3>>(m, sg, x, y, s); break;
    case 8: launch_mass_tensor_march_cfg<EDGE, MarchCfg<8, 64, 0>>(m, sg, x, y, s); break;
    case 9: launch_mass_tensor_march_cfg<EDGE, MarchCfg<4, 32, 0>>(m, sg, x, y, s); break;
    case 10: launch_mass_tensor_march_cfg<EDGE, MarchCfg<16, 32, 0>>(m, sg, x, y, s); break;
    case 11: launch_mass_tensor_march_cfg<EDGE, MarchCfg<4, 64, 5>>(m, sg, x, y, s); break;
    // 512^3 sweep (tools/sweep_tensor_mass.py): faces 8 x 32 threads over 32 planes, uncapped registers (124): 2.90 ms;
    // edges 4 x 32 threads over 64 planes at 96 registers (5 blocks / SM): 3.07 ms (uncapped, 116 registers: 3.42 ms;
    // capped at 80: spills, 3.9 ms)
    default:
      if (EDGE) launch_mass_tensor_march_cfg<EDGE, MarchCfg<4, 64, 5>>(m, sg, x, y, s);
      else launch_mass_tensor_march_cfg<EDGE, MarchCfg<8, 32, 0>>(m, sg, x, y, s);
      break;
  }
}

}  // namespace dzb
