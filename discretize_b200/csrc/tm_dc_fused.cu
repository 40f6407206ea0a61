// K8: y = G^T Me(sigma) G phi in ONE kernel (DC resistivity, nodal formulation).
//
// Reference composition (tutorials/pde/3_nodal_dirichlet_poisson.py:146-160):
//   G  = nodal_gradient            (operators/differential_operators.py:658-664)
//   Me = get_edge_inner_product(s) (operators/inner_products.py:69-99 ->
//                                   base/base_tensor_mesh.py:792-863, isotropic)
// which is a 7-point nodal stencil in flux form:
//   y[n] = sum_{edges e into n} f_e - sum_{edges e out of n} f_e,
//   f_e  = w_e (phi[head] - phi[tail]),  w_e = Me_e / L_e^2,
//   Me_e = 1/4 * sum over the (<=4) cells c around e of  V_c sigma_c.
// With V = hx hy hz this gives, for an x-edge at (i+1/2, j, k),
//   w = 1/(4 hx_i) * sum_{jj in {j-1,j}, kk in {k-1,k}} sigma(i,jj,kk) hy_jj hz_kk
// and cyclically for y and z; cells outside the mesh contribute nothing, which is
// exactly what the reference's Av^T (V sigma) does on boundary edges.
//
// Mapping (HBM-bound, FP64, no tensor cores -- nothing here is a contraction):
//   * one warp owns a strip of 32 consecutive node columns in x (30 outputs + one
//     halo lane on each side; x-neighbours are exchanged with warp shuffles),
//   * each lane keeps RY consecutive y-rows (+ halo rows) in registers,
//   * the warp marches along z keeping the previous plane / cell layer in
//     registers, so every phi and sigma value is loaded from HBM/L2 once per strip
//     and every output written once: algorithmic bytes = 8 (2 nN + nC).
//   * loads are 256-byte coalesced row segments; there is no shared memory and no
//     block-level synchronisation at all.
#include "tm_common.cuh"

namespace dzb {

constexpr int DC_WARPS = 4;      // warps per block, consecutive x strips
constexpr int DC_XOUT = 30;      // outputs per warp strip (lanes 1..30)

template <int RY>
__global__ void __launch_bounds__(DC_WARPS * 32)
    k_dc_nodal(TM m, const double *__restrict__ sigma, const double *__restrict__ phi, double *__restrict__ y,
               int kchunk) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const i64 nNx = m.nx + 1, nNy = m.ny + 1, nNz = m.nz + 1;
  const i64 X0 = ((i64)blockIdx.x * DC_WARPS + warp) * DC_XOUT;
  if (X0 >= nNx) return;
  const i64 i = X0 + lane - 1;          // node column of this lane (lane 0 / 31 are halo)
  const i64 j0 = (i64)blockIdx.y * RY;  // first output row
  const i64 kb = (i64)blockIdx.z * kchunk;
  const i64 ke = min(kb + (i64)kchunk, nNz);

  const bool node_ok = (i >= 0) && (i < nNx);
  const bool cell_ok = (i >= 0) && (i < m.nx) && (lane < 31);  // lane 31's cells are never needed
  const double hx = cell_ok ? __ldg(m.hx + i) : 0.0;
  const double qx = cell_ok ? 0.25 * __ldg(m.rhx + i) : 0.0;

  // row constants.  phi rows: j0-1 .. j0+RY  (RY+2);  cell rows: j0-1 .. j0+RY-1 (RY+1)
  double hy[RY + 1], qy[RY + 1];
  bool prow[RY + 2], crow[RY + 1];
#pragma unroll
  for (int r = 0; r < RY + 2; ++r) prow[r] = node_ok && (j0 - 1 + r >= 0) && (j0 - 1 + r < nNy);
#pragma unroll
  for (int r = 0; r < RY + 1; ++r) {
    const i64 jc = j0 - 1 + r;
    const bool ok = (jc >= 0) && (jc < m.ny);
    crow[r] = cell_ok && ok;
    hy[r] = ok ? __ldg(m.hy + jc) : 0.0;
    qy[r] = ok ? 0.25 * __ldg(m.rhy + jc) : 0.0;
  }

  const i64 pstride = nNx * nNy;     // node plane
  const i64 cstride = m.nx * m.ny;   // cell layer
  // pointers to row j0-1 of the current plane / layer (may be formally out of range; guarded)
  const double *pphi = phi + i + nNx * (j0 - 1);
  const double *psig = sigma + i + m.nx * (j0 - 1);

  double ph0[RY + 2], ph1[RY + 2];   // phi on planes k and k+1
  double Ap[RY], Bp[RY + 1];         // layer k-1: hz * (y-pair of sigma hy), hz * (x-pair of sigma hx)
  double fzp[RY];                    // z-flux between planes k-1 and k

  // ---- preamble: plane kb-1 / layer kb-1 ------------------------------------------------
  {
    const bool have = kb >= 1;
    double phm[RY + 2], sg[RY + 1];
#pragma unroll
    for (int r = 0; r < RY + 2; ++r) {
      phm[r] = (have && prow[r]) ? __ldg(pphi + (kb - 1) * pstride + r * nNx) : 0.0;
      ph0[r] = prow[r] ? __ldg(pphi + kb * pstride + r * nNx) : 0.0;
    }
#pragma unroll
    for (int r = 0; r < RY + 1; ++r) sg[r] = (have && crow[r]) ? __ldg(psig + (kb - 1) * cstride + r * m.nx) : 0.0;
    const double hz = have ? __ldg(m.hz + kb - 1) : 0.0;
    const double qz = have ? 0.25 * __ldg(m.rhz + kb - 1) : 0.0;
    double p[RY + 1];
#pragma unroll
    for (int r = 0; r < RY + 1; ++r) {
      p[r] = sg[r] * hy[r];
      const double t = sg[r] * hx;
      const double tl = __shfl_up_sync(0xffffffffu, t, 1);
      Bp[r] = hz * (tl + t);
    }
#pragma unroll
    for (int r = 0; r < RY; ++r) {
      const double a = p[r] + p[r + 1];
      Ap[r] = hz * a;
      const double ha = hx * a;
      const double hal = __shfl_up_sync(0xffffffffu, ha, 1);
      const double wz = qz * (hal + ha);
      fzp[r] = wz * (ph0[r + 1] - phm[r + 1]);
    }
  }

  // ---- march ------------------------------------------------------------------------------
  for (i64 k = kb; k < ke; ++k) {
    const bool havep = (k + 1) < nNz;   // plane k+1 exists
    const bool havec = k < m.nz;        // layer k exists
    double sg[RY + 1];
#pragma unroll
    for (int r = 0; r < RY + 2; ++r) ph1[r] = (havep && prow[r]) ? __ldg(pphi + (k + 1) * pstride + r * nNx) : 0.0;
#pragma unroll
    for (int r = 0; r < RY + 1; ++r) sg[r] = (havec && crow[r]) ? __ldg(psig + k * cstride + r * m.nx) : 0.0;
    const double hz = havec ? __ldg(m.hz + k) : 0.0;
    const double qz = havec ? 0.25 * __ldg(m.rhz + k) : 0.0;

    double p[RY + 1], B[RY + 1], fy[RY + 1];
#pragma unroll
    for (int r = 0; r < RY + 1; ++r) {
      p[r] = sg[r] * hy[r];
      const double t = sg[r] * hx;
      const double tl = __shfl_up_sync(0xffffffffu, t, 1);
      B[r] = hz * (tl + t);
      // y-edge between node rows (j0-1+r) and (j0+r), bordered by cell row j0-1+r
      const double wy = qy[r] * (Bp[r] + B[r]);
      fy[r] = wy * (ph0[r + 1] - ph0[r]);
      Bp[r] = B[r];
    }
#pragma unroll
    for (int r = 0; r < RY; ++r) {
      const double a = p[r] + p[r + 1];
      const double A = hz * a;
      // x-edge from this lane's node to the next one, node row j0+r
      const double wx = qx * (Ap[r] + A);
      const double phr = __shfl_down_sync(0xffffffffu, ph0[r + 1], 1);
      const double fx = wx * (phr - ph0[r + 1]);
      const double fxl = __shfl_up_sync(0xffffffffu, fx, 1);
      Ap[r] = A;
      // z-edge from plane k to k+1
      const double ha = hx * a;
      const double hal = __shfl_up_sync(0xffffffffu, ha, 1);
      const double wz = qz * (hal + ha);
      const double fz = wz * (ph1[r + 1] - ph0[r + 1]);
      const double out = (fxl - fx) + (fy[r] - fy[r + 1]) + (fzp[r] - fz);
      fzp[r] = fz;
      const i64 j = j0 + r;
      if (lane >= 1 && lane <= DC_XOUT && node_ok && j < nNy) y[i + nNx * (j + nNy * k)] = out;
    }
#pragma unroll
    for (int r = 0; r < RY + 2; ++r) ph0[r] = ph1[r];
  }
}

static int g_dc_ry = 4;
static int g_dc_kchunk = 32;

}  // namespace dzb

using namespace dzb;

// tuning knobs for experiments (not part of the stable ABI)
extern "C" void dzb_tune_dc(int ry, int kchunk) {
  if (ry == 2 || ry == 4 || ry == 8) g_dc_ry = ry;
  if (kchunk > 0) g_dc_kchunk = kchunk;
}

extern "C" int dzb_tm_dc_nodal(const DzbTensorMesh *mm, const double *sigma, const double *phi, double *y,
                               void *stream) {
  if (!mm || mm->nx <= 0 || mm->ny <= 0 || mm->nz <= 0) {
    set_error("dzb_tm_dc_nodal: invalid mesh");
    return -1;
  }
  TM m(*mm);
  const i64 nNx = m.nx + 1, nNy = m.ny + 1, nNz = m.nz + 1;
  const i64 strips = (nNx + DC_XOUT - 1) / DC_XOUT;
  const int ry = g_dc_ry;
  int kchunk = g_dc_kchunk;
  dim3 grid((unsigned)((strips + DC_WARPS - 1) / DC_WARPS), (unsigned)((nNy + ry - 1) / ry),
            (unsigned)((nNz + kchunk - 1) / kchunk));
  dim3 block(DC_WARPS * 32);
  cudaStream_t s = (cudaStream_t)stream;
  if (ry == 2) k_dc_nodal<2><<<grid, block, 0, s>>>(m, sigma, phi, y, kchunk);
  else if (ry == 4) k_dc_nodal<4><<<grid, block, 0, s>>>(m, sigma, phi, y, kchunk);
  else k_dc_nodal<8><<<grid, block, 0, s>>>(m, sigma, phi, y, kchunk);
  return check_launch("dzb_tm_dc_nodal");
}
