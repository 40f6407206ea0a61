// K9 "ring": y = C^T Mf(mui) C e + Me(sigma) e with a warp-specialised TMA ring (same scheme as K8).
//
// Flux form (derivation in DESIGN.md): with the line integrals  ex~ = hx ex, ey~ = hy ey, ez~ = hz ez,
//   circ_FX(i,j,k) = ez~(i,j+1,k) - ez~(i,j,k) - ey~(i,j,k+1) + ey~(i,j,k)
//   circ_FY(i,j,k) = ex~(i,j,k+1) - ex~(i,j,k) - ez~(i+1,j,k) + ez~(i,j,k)
//   circ_FZ(i,j,k) = ey~(i+1,j,k) - ey~(i,j,k) - ex~(i,j+1,k) + ex~(i,j,k)
//   G_F = (Mf_F / S_F^2) circ_F,   Mf_F = 1/2 sum over the (<=2) cells of V mui
//     w_FX = (hx[i-1] mui[i-1] + hx[i] mui[i]) / (2 hy hz)   (and cyclically)
//   yx = hx [ G_FY(k-1) - G_FY(k) - G_FZ(j-1) + G_FZ(j) ] + Me_x ex,  Me_x = 1/4 sum_{4 cells} V sigma
//   yy = hy [ -G_FX(k-1) + G_FX(k) + G_FZ(i-1) - G_FZ(i) ] + Me_y ey
//   yz = hz [ G_FX(j-1) - G_FX(j) - G_FY(i-1) + G_FY(i) ] + Me_z ez
// Cells outside the mesh have zero width, so boundary faces / edges get exactly the reference's
// one-sided sums (base/base_tensor_mesh.py:792-863 via Av^T (V prop)).
//
// A lane owns node column i and cell column i, RY rows (node rows j0..j0+RY-1 and cell rows of the same
// index), and marches in z.  Stage j of the ring holds ex, ey on node plane j+1 and ez, mui, sigma on cell
// layer j; plane j of ex / ey is carried in registers from the previous step, so a stage lives for one
// step and two stages are always in flight.  x-neighbours travel by warp shuffle (lanes 0 and 31 are halo).
#pragma once
#include "bulk_pipe.cuh"
#include "tm_common.cuh"

namespace dzb {

constexpr int MX_XOUT = 30;
constexpr int MX_MAXCHUNK = 128;

template <int RY, int NWC, int PARX, int NSLOT> struct MxRing {
  static constexpr int RN = RY + 2, RC = RY + 1;   // staged node rows (j0-1..j0+RY) / cell rows (j0-1..j0+RY-1)
  static constexpr int NCOL = NWC * MX_XOUT + 2;
  static constexpr int PX = (NCOL + 3) + ((((NCOL + 3) & 1) != PARX) ? 1 : 0);              // cell-column arrays
  static constexpr int PN = (NCOL + 3) + ((((NCOL + 3) & 1) != ((PARX + 1) & 1)) ? 1 : 0);  // node-column arrays
  static constexpr int pad(int n) { return (n + 4 + 1) & ~1; }
  static constexpr int O_EX = 0;
  static constexpr int O_EY = O_EX + pad(RN * PX);
  static constexpr int O_EZ = O_EY + pad(RC * PN);
  static constexpr int O_MU = O_EZ + pad(RN * PN);
  static constexpr int O_SG = O_MU + pad(RC * PX);
  static constexpr int STAGE = O_SG + pad(RC * PX);
  static constexpr size_t SMEM = (size_t)NSLOT * STAGE * 8 + 2 * NSLOT * 8 + 2 * RC * 8 + 2 * (MX_MAXCHUNK + 2) * 8 + 64;
};

// one staged array: absolute element address of (first staged column, virtual row j0-1, plane 0), strides, extents
struct MxArr {
  long long a0;      // element address
  long long pstride; // elements per plane / layer
  long long end;     // aligned-up end of the allocation (elements)
  int pitch;         // global row pitch (elements)
  int nrows_valid;   // rows 0 .. nrows_valid-1 exist
  int nplanes;       // planes / layers 0 .. nplanes-1 exist
};

template <int RY, int NWC, int PARX, int NSLOT>
__global__ void __launch_bounds__((NWC + 1) * 32, 1)
    k_maxwell_ring(TM m, const double *__restrict__ mui, const double *__restrict__ sigma, const double *__restrict__ e,
                   double *__restrict__ y, int kchunk, int nwc, int k_begin, int k_end) {
  typedef MxRing<RY, NWC, PARX, NSLOT> G;
  constexpr int RN = G::RN, RC = G::RC, S = NSLOT, PX = G::PX, PN = G::PN;
  extern __shared__ __align__(128) unsigned char mx_smem_raw[];
  double *tiles = reinterpret_cast<double *>(mx_smem_raw);
  uint64_t *full = reinterpret_cast<uint64_t *>(tiles + S * G::STAGE);
  uint64_t *empty = full + S;
  double *s_hy = reinterpret_cast<double *>(empty + S);   // cell rows j0-1 .. j0+RY-1 (0 outside)
  double *s_ry = s_hy + RC;
  double *s_hz = s_ry + RC;                               // layers kb-1 .. ke-1 (0 outside)
  double *s_rz = s_hz + (MX_MAXCHUNK + 2);

  const int nx = (int)m.nx, ny = (int)m.ny, nz = (int)m.nz;
  const int nNx = nx + 1, nNy = ny + 1, nNz = nz + 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nthreads = (nwc + 1) * 32;
  const int j0 = blockIdx.y * RY;
  const int Xb = blockIdx.x * (nwc * MX_XOUT);
  const int kb = k_begin + blockIdx.z * kchunk;    // output planes kb .. ke-1 (layers kb .. min(ke,nz)-1)
  const int ke = min(kb + kchunk, k_end);
  const int jfirst = kb - 2;

  for (int t = tid; t < S * G::STAGE; t += nthreads) tiles[t] = 0.0;
  if (tid < RC) {
    const int jc = j0 - 1 + tid;
    const bool ok = (jc >= 0) && (jc < ny);
    s_hy[tid] = ok ? m.hy[jc] : 0.0;
    s_ry[tid] = ok ? m.rhy[jc] : 0.0;
  }
  for (int t = tid; t <= ke - kb; t += nthreads) {
    const int k = kb - 1 + t;
    const bool ok = (k >= 0) && (k < nz);
    s_hz[t] = ok ? m.hz[k] : 0.0;
    s_rz[t] = ok ? m.rhz[k] : 0.0;
  }
  if (tid == 0) {
    for (int q = 0; q < S; ++q) {
      mbar_init(&full[q], 1);
      mbar_init(&empty[q], nwc);
    }
    fence_barrier_init();
  }
  fence_proxy_async_smem();
  __syncthreads();

  // global layout of the stacked edge vector (reference order [Ex;Ey;Ez], x fastest)
  const long long nEx = (long long)nx * nNy * nNz, nEy = (long long)nNx * ny * nNz;
  const int cs = max(Xb - 1, 0);
  const long long eaddr = (long long)((unsigned long long)e >> 3);
  MxArr A[5];
  // 0: ex (cell cols, node rows, node planes)   1: ey (node cols, cell rows, node planes)
  // 2: ez (node cols, node rows, cell layers)   3: mui, 4: sigma (cell cols, cell rows, cell layers)
  A[0] = {eaddr + cs + (long long)nx * (j0 - 1), (long long)nx * nNy, 0, nx, nNy, nNz};
  A[1] = {eaddr + nEx + cs + (long long)nNx * (j0 - 1), (long long)nNx * ny, 0, nNx, ny, nNz};
  A[2] = {eaddr + nEx + nEy + cs + (long long)nNx * (j0 - 1), (long long)nNx * nNy, 0, nNx, nNy, nz};
  A[3] = {(long long)((unsigned long long)mui >> 3) + cs + (long long)nx * (j0 - 1), (long long)nx * ny, 0, nx, ny, nz};
  A[4] = {(long long)((unsigned long long)sigma >> 3) + cs + (long long)nx * (j0 - 1), (long long)nx * ny, 0, nx, ny, nz};
  const int nq = ke - jfirst;

  if (warp == nwc) {
    // ================= producer warp =================
    // Lane l owns row copy l of every stage (17 copies for RY = 2): the per-row address arithmetic runs
    // SIMT across the lanes, lane 0 posts the stage's byte count, then every lane issues its bulk copy.
    {
      const long long nE = nEx + nEy + (long long)nNx * nNy * nz;
      A[0].end = A[1].end = A[2].end = (eaddr + nE + 1) & ~1ll;
      A[3].end = ((long long)((unsigned long long)mui >> 3) + (long long)nx * ny * nz + 1) & ~1ll;
      A[4].end = ((long long)((unsigned long long)sigma >> 3) + (long long)nx * ny * nz + 1) & ~1ll;
      const int ncol = nwc * MX_XOUT + 2;
      // which (array, row) this lane copies
      int a = -1, r = 0;
      {
        int l = lane;
        const int nrow[5] = {RN, RC, RN, RC, RC};
#pragma unroll
        for (int q = 0; q < 5; ++q) {
          if (a < 0 && l < nrow[q]) { a = q; r = l; }
          if (a < 0) l -= nrow[q];
        }
      }
      const bool have = a >= 0;
      const int aa = have ? a : 0;
      const int off = aa == 0 ? G::O_EX : aa == 1 ? G::O_EY : aa == 2 ? G::O_EZ : aa == 3 ? G::O_MU : G::O_SG;
      const int spitch = (aa == 1 || aa == 2) ? PN : PX;
      const int jr = j0 - 1 + r;
      const bool row_ok = have && jr >= 0 && jr < A[aa].nrows_valid;
      const long long rowbase = A[aa].a0 + (long long)r * A[aa].pitch;
      const long long a0 = A[aa].a0, pstr = A[aa].pstride, aend = A[aa].end;
      const int npl = A[aa].nplanes, ploff = aa < 2 ? 1 : 0;
      const int soff = off + 2 + r * spitch;
      for (int q = 0; q < nq; ++q) {
        const int jst = jfirst + q, slot = q % S, n = q / S;
        if (n >= 1) mbar_wait(&empty[slot], (n - 1) & 1);
        const int pl = jst + ploff;
        const bool on = row_ok && pl >= 0 && pl < npl;
        const long long Ad = rowbase + pstr * pl, Aal = Ad & ~1ll;
        long long nel = ((Ad - Aal) + ncol + 1) & ~1ll;
        nel = min(nel, aend - Aal);
        const unsigned bytes = on ? (unsigned)(nel * 8) : 0u;
        const unsigned total = __reduce_add_sync(0xffffffffu, bytes);
        if (lane == 0) mbar_arrive_expect_tx(&full[slot], total);
        __syncwarp();
        if (bytes) {
          const int lead0 = (int)((a0 + pstr * pl) & 1);
          bulk_g2s(tiles + slot * G::STAGE + soff + lead0 - (int)(Ad - Aal), reinterpret_cast<const void *>(Aal << 3),
                   bytes, &full[slot]);
        }
      }
    }
    return;
  }

  // ================= consumer warps =================
  const int i = Xb + warp * MX_XOUT + lane - 1;   // node column and cell column of this lane
  const bool lane_out = (lane >= 1) && (lane <= MX_XOUT);
  const bool ncol_ok = lane_out && (i >= 0) && (i < nNx);
  const bool ccol_ok = lane_out && (i >= 0) && (i < nx);
  const bool cell_ok = (i >= 0) && (i < nx);
  const int icc = min(max(i, 0), nx - 1);
  const double hx = cell_ok ? __ldg(m.hx + icc) : 0.0;
  const double rx = cell_ok ? __ldg(m.rhx + icc) : 0.0;
  const int c2 = i - cs + 2;
  int lead[5], pb[5];
#pragma unroll
  for (int a = 0; a < 5; ++a) {
    lead[a] = (int)(A[a].a0 & 1);
    pb[a] = (int)(A[a].pstride & 1);
  }

  // carried state
  double ex0[RN], ey0[RC];          // ex, ey on plane j (loaded as plane j+1 one step earlier)
  double gfxp[RC], gfyp[RY];        // G_FX, G_FY of layer j-1
  double mzp[RC], ap[RY], bp[RC];   // layer j-1: hz*mui, y-pair / x-pair sums of V sigma
#pragma unroll
  for (int r = 0; r < RN; ++r) ex0[r] = 0.0;
#pragma unroll
  for (int r = 0; r < RC; ++r) ey0[r] = gfxp[r] = mzp[r] = bp[r] = 0.0;
#pragma unroll
  for (int r = 0; r < RY; ++r) gfyp[r] = ap[r] = 0.0;

  // output pointers (plane / layer kb-1 at the first computing step)
  const long long pEx = (long long)nx * nNy, pEy = (long long)nNx * ny, pEz = (long long)nNx * nNy;
  double *yx = y + ((long long)(kb - 1) * pEx + icc + (long long)nx * j0);
  double *yy = y + (nEx + (long long)(kb - 1) * pEy + min(max(i, 0), nNx - 1) + (long long)nNx * j0);
  double *yz = y + (nEx + nEy + (long long)(kb - 1) * pEz + min(max(i, 0), nNx - 1) + (long long)nNx * j0);
  unsigned wx = 0, wy = 0, wz = 0;   // existing output rows per component
#pragma unroll
  for (int r = 0; r < RY; ++r) {
    if (ccol_ok && (j0 + r) < nNy) wx |= 1u << r;
    if (ncol_ok && (j0 + r) < ny) wy |= 1u << r;
    if (ncol_ok && (j0 + r) < nNy) wz |= 1u << r;
  }

  for (int qb = 0; qb < nq; qb += S) {
    const unsigned par = (unsigned)(qb / S) & 1u;
#pragma unroll
    for (int u = 0; u < S; ++u) {
      const int q = qb + u;
      if (q >= nq) break;
      const int j = jfirst + q;          // this step works on plane j / layer j
      constexpr int dummy = 0;
      (void)dummy;
      mbar_wait(&full[u], par);
      const double *T = tiles + u * G::STAGE;
      const double *tEX = T + G::O_EX + c2 + ((lead[0] + pb[0] * (j + 1)) & 1);
      const double *tEY = T + G::O_EY + c2 + ((lead[1] + pb[1] * (j + 1)) & 1);
      const double *tEZ = T + G::O_EZ + c2 + ((lead[2] + pb[2] * j) & 1);
      const double *tMU = T + G::O_MU + c2 + ((lead[3] + pb[3] * j) & 1);
      const double *tSG = T + G::O_SG + c2 + ((lead[4] + pb[4] * j) & 1);
      double ex1[RN], ey1[RC];
#pragma unroll
      for (int r = 0; r < RN; ++r) ex1[r] = tEX[r * PX];
#pragma unroll
      for (int r = 0; r < RC; ++r) ey1[r] = tEY[r * PN];

      if (q >= 1) {
        const int t = j - (kb - 1);
        const double hz = s_hz[t], rz = s_rz[t];
        const double hxz = hx * hz;
        const double cfy = 0.5 * rx * rz;     // w_FY = (my[j-1] + my[j]) * cfy
        const bool out_plane = j >= kb;       // planes kb..ke-1 are written (step kb-1 only builds state)
        const bool out_layer = out_plane && (j < nz);

        double ez_[RN], ezt[RN];
#pragma unroll
        for (int r = 0; r < RN; ++r) {
          ez_[r] = tEZ[r * PN];
          ezt[r] = hz * ez_[r];
        }
        double s[RC], sl[RC], my[RC], gfx[RC], gfz[RC], gfzl[RC];
#pragma unroll
        for (int r = 0; r < RC; ++r) {
          const double hy = s_hy[r], ry = s_ry[r];
          const double mu = tMU[r * PX], sg = tSG[r * PX];
          s[r] = (sg * hy) * hxz;                               // V sigma (0 outside the mesh)
          sl[r] = __shfl_up_sync(0xffffffffu, s[r], 1);
          const double mx = mu * hx;
          const double mxl = __shfl_up_sync(0xffffffffu, mx, 1);
          my[r] = mu * hy;
          const double mz = mu * hz;
          // line integrals on plane j
          const double eyt0 = hy * ey0[r], eyt1 = hy * ey1[r];
          const double eyt0r = __shfl_down_sync(0xffffffffu, eyt0, 1);
          // z-face (plane j, cell row r): needs ex~ on node rows r and r+1 of plane j
          const double circz = (eyt0r - eyt0) - (hx * ex0[r + 1] - hx * ex0[r]);
          gfz[r] = ((mzp[r] + mz) * (0.5 * rx * ry)) * circz;
          gfzl[r] = __shfl_up_sync(0xffffffffu, gfz[r], 1);
          mzp[r] = mz;
          // x-face (layer j, cell row r)
          const double circx = (ezt[r + 1] - ezt[r]) - (eyt1 - eyt0);
          gfx[r] = ((mxl + mx) * (0.5 * ry * rz)) * circx;
        }
        // yy on plane j, cell rows j0 .. j0+RY-1  (index r+1)
#pragma unroll
        for (int r = 0; r < RY; ++r) {
          const double b = sl[r + 1] + s[r + 1];
          const double mey = 0.25 * (bp[r + 1] + b);
          const double v = s_hy[r + 1] * ((gfx[r + 1] - gfxp[r + 1]) + (gfzl[r + 1] - gfz[r + 1])) + mey * ey0[r + 1];
          if (out_plane && ((wy >> r) & 1u)) yy[r * nNx] = v;
        }
#pragma unroll
        for (int r = 0; r < RC; ++r) bp[r] = sl[r] + s[r];
        // yx on plane j and yz on layer j, node rows j0 .. j0+RY-1 (ex row index r+1; cell rows r, r+1)
#pragma unroll
        for (int r = 0; r < RY; ++r) {
          const double a = s[r] + s[r + 1];
          const double al = sl[r] + sl[r + 1];
          // y-face (layer j, node row r): ex~ on planes j, j+1; ez~ at columns i, i+1
          const double eztr = __shfl_down_sync(0xffffffffu, ezt[r + 1], 1);
          const double circy = (hx * ex1[r + 1] - hx * ex0[r + 1]) - (eztr - ezt[r + 1]);
          const double gfy = ((my[r] + my[r + 1]) * cfy) * circy;
          const double gfyl = __shfl_up_sync(0xffffffffu, gfy, 1);
          const double vx = hx * ((gfyp[r] - gfy) + (gfz[r + 1] - gfz[r])) + (0.25 * (ap[r] + a)) * ex0[r + 1];
          const double vz = hz * ((gfx[r] - gfx[r + 1]) + (gfy - gfyl)) + (0.25 * (a + al)) * ez_[r + 1];
          gfyp[r] = gfy;
          ap[r] = a;
          if (out_plane && ((wx >> r) & 1u)) yx[r * nx] = vx;
          if (out_layer && ((wz >> r) & 1u)) yz[r * nNx] = vz;
        }
#pragma unroll
        for (int r = 0; r < RC; ++r) gfxp[r] = gfx[r];
        yx += pEx;
        yy += pEy;
        yz += pEz;
      }
#pragma unroll
      for (int r = 0; r < RN; ++r) ex0[r] = ex1[r];
#pragma unroll
      for (int r = 0; r < RC; ++r) ey0[r] = ey1[r];
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[u]);
    }
  }
}

}  // namespace dzb
