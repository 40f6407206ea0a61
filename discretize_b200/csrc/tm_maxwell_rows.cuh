// K9 "rows": y = C^T Mf(mui) C e + Me(sigma) e, one warp per (31-column strip, y-row), TMA-staged z-march.
//
// Same flux form as the first ring kernel (tm_maxwell_ring.cuh; composition of
// tests/boundaries/test_boundary_maxwell.py:95-99), written with UNSCALED edge values:
//   circ_FX(i,j,k) = hz[k] (ez(i,j+1,k) - ez(i,j,k)) - hy[j] (ey(i,j,k+1) - ey(i,j,k))
//   circ_FY(i,j,k) = hx[i] (ex(i,j,k+1) - ex(i,j,k)) - hz[k] (ez(i+1,j,k) - ez(i,j,k))
//   circ_FZ(i,j,k) = hy[j] (ey(i+1,j,k) - ey(i,j,k)) - hx[i] (ex(i,j+1,k) - ex(i,j,k))
//   G_FX = (hx[i-1] mui(i-1,j,k) + hx[i] mui(i,j,k)) / (2 hy[j] hz[k]) * circ_FX     (= Mf_F / S_F^2 * circ)
//   G_FY = (hy[j-1] mui(i,j-1,k) + hy[j] mui(i,j,k)) / (2 hx[i] hz[k]) * circ_FY
//   G_FZ = (hz[k-1] mui(i,j,k-1) + hz[k] mui(i,j,k)) / (2 hx[i] hy[j]) * circ_FZ
//   yx(i,j,k) = hx[i] [ G_FY(k-1) - G_FY(k) - G_FZ(j-1) + G_FZ(j) ] + Me_x ex,  Me_x = 1/4 sum_{j-1..j, k-1..k} V sigma
//   yy(i,j,k) = hy[j] [ G_FX(k) - G_FX(k-1) + G_FZ(i-1) - G_FZ(i) ] + Me_y ey,  Me_y = 1/4 sum_{i-1..i, k-1..k} V sigma
//   yz(i,j,k) = hz[k] [ G_FX(j-1) - G_FX(j) - G_FY(i-1) + G_FY(i) ] + Me_z ez,  Me_z = 1/4 sum_{i-1..i, j-1..j} V sigma
// Widths outside the mesh are 0 (and their reciprocals 0), which gives the reference's one-sided boundary sums
// (base/base_tensor_mesh.py:792-863, Av^T (V prop)).
//
// What changed against the first ring kernel (ncu at 512^3: 168 registers, 10 warps / SM, 200 warp instructions per
// strip-row-plane, issue- and latency-bound at 0.71-0.77 of HBM):
//   * a thread owns ONE (i, j) column, not two rows + a recomputed halo row: the two face fluxes a row needs from the row
//     below (G_FX, G_FZ at j-1) come from the warp that owns that row through a 512-byte shared-memory mailbox per
//     (ring slot, row, strip) with its own mbarrier; the block's bottom halo row is computed by one light "halo" warp per
//     strip.  No y-redundancy in the arithmetic, ~100 instructions per strip-row-plane, < 100 registers, 2 blocks / SM.
//   * right neighbours (ey, ez at i+1) are read straight from the staged tile, so only lane 0 is an x-halo lane
//     (31 outputs per warp); left-neighbour FLUXES still travel by shuffle.
//   * the producer never reads outside the arrays (plan_row_copy): no "8 readable bytes after the array" clause.
#pragma once
#include "bulk_pipe.cuh"
#include "tm_common.cuh"

namespace dzb {

constexpr int M2_XOUT = 31;       // output columns per strip: lanes 1..31 (lane 0 = left halo)
constexpr int M2_MAXCHUNK = 160;  // max z-planes per block (size of the hz tables)

// compile-time geometry: XS strips x R rows per block, PARX = nx & 1, NSLOT ring slots
template <int XS, int R, int PARX, int NSLOT, int NPROD> struct Mx2 {
  static constexpr int NCONS = XS * R;                 // consumer warps: warp w -> strip w % XS, row w / XS
  static constexpr int NHALO = XS;                     // halo warps: cell row j0-1 of strip h
  static constexpr int NWARPS = NCONS + NHALO + NPROD;  // + producers
  static constexpr int NT = NWARPS * 32;
  static constexpr int RN = R + 2, RC = R + 1;         // staged node rows j0-1 .. j0+R, cell rows j0-1 .. j0+R-1
  static constexpr int NCOL = XS * M2_XOUT + 2;        // staged columns Xb-1 .. Xb+XS*31
  // shared-memory row pitches (doubles), parity-matched to the global pitches nx (cell columns) and nx+1 (node columns)
  static constexpr int PX = (NCOL + 3) + ((((NCOL + 3) & 1) != PARX) ? 1 : 0);
  static constexpr int PN = (NCOL + 3) + ((((NCOL + 3) & 1) != ((PARX + 1) & 1)) ? 1 : 0);
  static constexpr int pad(int n) { return (n + 4 + 1) & ~1; }
  static constexpr int O_EX = 0;                        // ex: cell columns, node rows
  static constexpr int O_EY = O_EX + pad(RN * PX);      // ey: node columns, cell rows
  static constexpr int O_EZ = O_EY + pad(RC * PN);      // ez: node columns, node rows
  static constexpr int O_MU = O_EZ + pad(RN * PN);      // mui, sigma: cell columns, cell rows
  static constexpr int O_SG = O_MU + pad(RC * PX);
  static constexpr int STAGE = O_SG + pad(RC * PX);
  static constexpr int NCOPY = 2 * RN + 3 * RC;         // row copies per stage
  static constexpr int NCPL = (NCOPY + 32 * NPROD - 1) / (32 * NPROD);   // ... per producer lane
  static constexpr int XROWS = R;                       // mailbox rows: 0 = halo row j0-1, rr = row j0+rr-1
  static constexpr int XCH = NSLOT * XROWS * XS * 64;   // doubles: [slot][rr][strip][G_FX | G_FZ][lane]
  static constexpr int NBAR = 2 * NSLOT + NSLOT * XROWS * XS;
  static constexpr size_t SMEM = (size_t)(NSLOT * STAGE + XCH) * 8 + (size_t)NBAR * 8 + 2 * (M2_MAXCHUNK + 2) * 8 + 16;
};

// (array, row) of row copy c in the order ex[RN], ey[RC], ez[RN], mui[RC], sigma[RC]
template <int R> struct M2Copy {
  static constexpr int RN = R + 2, RC = R + 1;
  static constexpr int arr(int c) { return c < RN ? 0 : c < RN + RC ? 1 : c < 2 * RN + RC ? 2 : c < 2 * RN + 2 * RC ? 3 : 4; }
  static constexpr int row(int c) {
    return c < RN ? c : c < RN + RC ? c - RN : c < 2 * RN + RC ? c - RN - RC : c < 2 * RN + 2 * RC ? c - 2 * RN - RC : c - 2 * RN - 2 * RC;
  }
};

template <int XS, int R, int PARX, int NSLOT, int NPROD, int MINB>
__global__ void __launch_bounds__((XS *R + XS + NPROD) * 32, MINB)
    k_maxwell_rows(TM m, const double *__restrict__ mui, const double *__restrict__ sigma, const double *__restrict__ e,
                   double *__restrict__ y, int kchunk, int k_begin, int k_end, int dbg) {
  typedef Mx2<XS, R, PARX, NSLOT, NPROD> G;
  constexpr int S = NSLOT, PX = G::PX, PN = G::PN, RN = G::RN, RC = G::RC;
  static_assert(S % 2 == 0, "the lead-bit bookkeeping assumes an even ring depth");
  extern __shared__ __align__(128) unsigned char m2_smem_raw[];
  double *tiles = reinterpret_cast<double *>(m2_smem_raw);          // [S][STAGE]
  double *xch = tiles + S * G::STAGE;                               // mailboxes
  uint64_t *full = reinterpret_cast<uint64_t *>(xch + G::XCH);      // [S]
  uint64_t *empty = full + S;                                       // [S]
  uint64_t *xready = empty + S;                                     // [S][XROWS][XS]
  double *s_hz = reinterpret_cast<double *>(xready + S * G::XROWS * XS);   // layers kb-1 .. ke-1 (0 outside)
  double *s_rzh = s_hz + (M2_MAXCHUNK + 2);                         // 0.5 / hz
  int *s_f = reinterpret_cast<int *>(s_rzh + (M2_MAXCHUNK + 2));    // regular stage range agreed by the producer warps

  const int nx = (int)m.nx, ny = (int)m.ny, nz = (int)m.nz;
  const int nNx = nx + 1, nNy = ny + 1, nNz = nz + 1;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int j0 = blockIdx.y * R;
  const int Xb = blockIdx.x * (XS * M2_XOUT);
  const int kb = k_begin + blockIdx.z * kchunk;     // output planes kb .. ke-1 (layers kb .. min(ke, nz)-1)
  const int ke = min(kb + kchunk, k_end);
  const int jfirst = kb - 2;                        // stage q holds ex, ey on plane jfirst+q+1 and ez, mui, sigma on layer jfirst+q
  const int nq = ke - jfirst;

  for (int t = tid; t < S * G::STAGE + G::XCH; t += G::NT) tiles[t] = 0.0;
  for (int t = tid; t <= ke - kb; t += G::NT) {
    const int k = kb - 1 + t;
    const bool ok = (k >= 0) && (k < nz);
    s_hz[t] = ok ? m.hz[k] : 0.0;
    s_rzh[t] = ok ? 0.5 * m.rhz[k] : 0.0;
  }
  if (tid == 0) {
    for (int q = 0; q < S; ++q) {
      mbar_init(&full[q], NPROD);
      mbar_init(&empty[q], G::NCONS + G::NHALO);
    }
    for (int q = 0; q < S * G::XROWS * XS; ++q) mbar_init(&xready[q], 1);
    s_f[0] = 0;
    s_f[1] = 0x7fffffff;
    fence_barrier_init();
  }
  fence_proxy_async_smem();
  __syncthreads();

  // global layout of the stacked edge vector (reference order [Ex; Ey; Ez], x fastest)
  const long long nEx = (long long)nx * nNy * nNz, nEy = (long long)nNx * ny * nNz, nEz = (long long)nNx * nNy * nz;
  const long long nCc = (long long)nx * ny * nz;
  const int cs = max(Xb - 1, 0);                    // first staged column
  const long long eaddr = (long long)((unsigned long long)e >> 3);
  const long long maddr = (long long)((unsigned long long)mui >> 3), saddr = (long long)((unsigned long long)sigma >> 3);
  // per staged array: element address of (column cs, virtual row j0-1, plane 0) and the plane stride
  const long long a0[5] = {eaddr + cs + (long long)nx * (j0 - 1), eaddr + nEx + cs + (long long)nNx * (j0 - 1),
                           eaddr + nEx + nEy + cs + (long long)nNx * (j0 - 1), maddr + cs + (long long)nx * (j0 - 1),
                           saddr + cs + (long long)nx * (j0 - 1)};
  const long long pstr[5] = {(long long)nx * nNy, (long long)nNx * ny, (long long)nNx * nNy, (long long)nx * ny,
                             (long long)nx * ny};

  if (warp >= G::NCONS + G::NHALO) {
    // ================= producer warps =================
    const int prod = warp - (G::NCONS + G::NHALO);
    // Lane l owns row copies l, l+32 of every stage: the address arithmetic runs SIMT across the lanes, lane 0 posts the
    // stage's byte count, then every lane issues its bulk copies.
    // the three edge components are one allocation: only its two ends are hard limits
    const long long abeg[5] = {eaddr, eaddr, eaddr, maddr, saddr};
    const long long eend = eaddr + nEx + nEy + nEz;
    const long long aend[5] = {eend, eend, eend, maddr + nCc, saddr + nCc};
    long long c_rowbase[G::NCPL], c_pstr[G::NCPL], c_beg[G::NCPL], c_end[G::NCPL], c_a0[G::NCPL];
    int c_soff[G::NCPL], c_npl[G::NCPL], c_ploff[G::NCPL];
    bool c_ok[G::NCPL], c_model[G::NCPL];
#pragma unroll
    for (int t = 0; t < G::NCPL; ++t) {
      int l = (lane + 32 * t) * NPROD + prod, a = -1, r = 0;
      const int nrow[5] = {RN, RC, RN, RC, RC};
#pragma unroll
      for (int q = 0; q < 5; ++q) {
        if (a < 0 && l < nrow[q]) { a = q; r = l; }
        if (a < 0) l -= nrow[q];
      }
      const bool have = a >= 0;
      const int aa = have ? a : 0;
      const int off = aa == 0 ? G::O_EX : aa == 1 ? G::O_EY : aa == 2 ? G::O_EZ : aa == 3 ? G::O_MU : G::O_SG;
      const int spitch = (aa == 1 || aa == 2) ? PN : PX;
      const int gpitch = (aa == 1 || aa == 2) ? nNx : nx;
      const int nrows_valid = (aa == 0 || aa == 2) ? nNy : ny;
      const int jr = j0 - 1 + r;
      c_ok[t] = have && jr >= 0 && jr < nrows_valid;
      c_rowbase[t] = a0[aa] + (long long)r * gpitch;
      c_a0[t] = a0[aa];
      c_pstr[t] = pstr[aa];
      c_beg[t] = abeg[aa];
      c_end[t] = aend[aa];
      c_npl[t] = aa < 2 ? nNz : nz;
      c_ploff[t] = aa < 2 ? 1 : 0;
      c_model[t] = aa >= 2;          // ez / mui / sigma of the first stage (layer kb-2) are never read
      c_soff[t] = off + 2 + r * spitch;
    }
    // General stage: any row / plane may be missing or touch an end of its array.
    auto stage_general = [&](int q) {
      const int jst = jfirst + q, slot = q % S, n = q / S;
      if (n >= 1) mbar_wait(&empty[slot], (n - 1) & 1);
      double *T = tiles + slot * G::STAGE;
      RowCopy rc[G::NCPL];
      long long lo[G::NCPL];
      int dst0[G::NCPL];
      unsigned bytes = 0;
      bool ragged = false;
#pragma unroll
      for (int t = 0; t < G::NCPL; ++t) {
        const int pl = jst + c_ploff[t];
        const bool on = c_ok[t] && pl >= 0 && pl < c_npl[t] && !(q == 0 && c_model[t]);
        lo[t] = c_rowbase[t] + c_pstr[t] * pl;
        rc[t] = plan_row_copy(lo[t], lo[t] + G::NCOL, c_beg[t], c_end[t]);
        if (!on) { rc[t].bytes = 0; rc[t].head = rc[t].tail = false; }
        // element lo lands at dst0: one lead bit per (array, plane), the same for all rows of the stage
        dst0[t] = c_soff[t] + (int)((c_a0[t] + c_pstr[t] * pl) & 1);
        bytes += rc[t].bytes;
        ragged |= rc[t].head | rc[t].tail;
      }
      const unsigned total = __reduce_add_sync(0xffffffffu, bytes);
      if (__any_sync(0xffffffffu, ragged)) {   // only the two rows that hold the first / last element of an odd-aligned array
#pragma unroll
        for (int t = 0; t < G::NCPL; ++t) {
          if (rc[t].head) T[dst0[t]] = *reinterpret_cast<const double *>(lo[t] << 3);
          if (rc[t].tail) T[dst0[t] + (int)(rc[t].hi - 1 - lo[t])] = *reinterpret_cast<const double *>((rc[t].hi - 1) << 3);
        }
        fence_proxy_async_smem();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive_expect_tx(&full[slot], total);
      __syncwarp();
#pragma unroll
      for (int t = 0; t < G::NCPL; ++t)
        if (rc[t].bytes)
          bulk_g2s(T + dst0[t] + (int)(rc[t].tlo - lo[t]), reinterpret_cast<const void *>(rc[t].tlo << 3), rc[t].bytes,
                   &full[slot]);
    };

    // Regular stages [f0, f1): every existing row of every array is copied whole and no copy comes near an end of its
    // array.  There the producers are on the critical path of the ring, so the per-stage work is cut to: wait for the
    // slot, post the (precomputed) byte count, one 64-bit add + one bulk copy per row.  A bulk copy is one instruction per
    // WARP fed from uniform registers (SASS UBLKCP [UR], [UR], UR): the lanes' copies are issued by a serial loop that
    // extracts one lane's operands at a time, which is why the rows of a stage are dealt to NPROD producer warps.
    int f0 = 0, f1 = nq;
#pragma unroll
    for (int t = 0; t < G::NCPL; ++t) {
      if (c_ok[t]) {
        int qlo = max(c_model[t] ? 1 : 0, -jfirst - c_ploff[t]);
        int qhi = min(nq, c_npl[t] - jfirst - c_ploff[t]);
        while (qlo < qhi && c_rowbase[t] + c_pstr[t] * (jfirst + qlo + c_ploff[t]) - 1 < c_beg[t]) ++qlo;
        while (qhi > qlo && c_rowbase[t] + c_pstr[t] * (jfirst + qhi - 1 + c_ploff[t]) + G::NCOL + 1 > c_end[t]) --qhi;
        f0 = max(f0, qlo);
        f1 = min(f1, qhi);
      }
    }
    // all producer warps must agree on the range: combine through the mailbox area's spare barrier-free scratch
    f0 = __reduce_max_sync(0xffffffffu, f0);
    f1 = __reduce_min_sync(0xffffffffu, f1);
    if (NPROD > 1) {
      if (lane == 0) {
        atomicMax(&s_f[0], f0);
        atomicMin(&s_f[1], f1);
      }
      asm volatile("bar.sync 1, %0;" ::"r"(NPROD * 32) : "memory");
      f0 = s_f[0];
      f1 = s_f[1];
    }
    constexpr unsigned BY0 = ((G::NCOL + 1) & ~1) * 8, BY1 = ((G::NCOL + 2) & ~1) * 8;   // row start even / odd
    long long src[G::NCPL];
    int lead_q[G::NCPL], lead_step[G::NCPL];
    unsigned tot[2] = {0u, 0u};
#pragma unroll
    for (int t = 0; t < G::NCPL; ++t) {
      const int pl = jfirst + f0 + c_ploff[t];
      src[t] = c_rowbase[t] + c_pstr[t] * pl;
      lead_q[t] = (int)((c_a0[t] + c_pstr[t] * pl) & 1);
      lead_step[t] = (int)(c_pstr[t] & 1);
      if (c_ok[t]) {
        tot[0] += (src[t] & 1) ? BY1 : BY0;
        tot[1] += ((src[t] + c_pstr[t]) & 1) ? BY1 : BY0;
      }
    }
    tot[0] = __reduce_add_sync(0xffffffffu, tot[0]);
    tot[1] = __reduce_add_sync(0xffffffffu, tot[1]);

    int q = 0;
    for (; q < min(f0, nq); ++q) stage_general(q);
    for (; q < f1; ++q) {
      const int slot = q % S, n = q / S;
      if (n >= 1) mbar_wait(&empty[slot], (n - 1) & 1);
      if (lane == 0) mbar_arrive_expect_tx(&full[slot], dbg == 2 ? 0u : tot[(q - f0) & 1]);
      __syncwarp();
      double *T = tiles + slot * G::STAGE;
#pragma unroll
      for (int t = 0; t < G::NCPL; ++t) {
        if (c_ok[t] && dbg != 2) {
          const int d = (int)(src[t] & 1);
          bulk_g2s(T + c_soff[t] + lead_q[t] - d, reinterpret_cast<const void *>((src[t] - d) << 3), d ? BY1 : BY0,
                   &full[slot]);
        }
        src[t] += c_pstr[t];
        lead_q[t] ^= lead_step[t];
      }
    }
    for (; q < nq; ++q) stage_general(q);
    return;
  }

  // ================= consumer / halo warps =================
  // `warp` comes out of a shuffle so that the compiler knows every role decision below is warp-uniform.
  const bool halo = warp >= G::NCONS;
  const int s = halo ? warp - G::NCONS : warp % XS;   // strip
  const int r = halo ? -1 : warp / XS;                // row inside the block (-1: the halo row below it)
  const int i = Xb + s * M2_XOUT + lane - 1;          // node column and cell column of this lane
  const int j = j0 + r;                               // node row and cell row
  const bool ci = (i >= 0) && (i < nx), cil = (i >= 1) && (i <= nx);
  const bool cj = (j >= 0) && (j < ny), cjd = (j >= 1) && (j <= ny);
  const double hx = ci ? __ldg(m.hx + i) : 0.0, rx = ci ? __ldg(m.rhx + i) : 0.0;
  const double hxl = cil ? __ldg(m.hx + i - 1) : 0.0;
  const double hy = cj ? __ldg(m.hy + j) : 0.0, ry = cj ? __ldg(m.rhy + j) : 0.0;
  const double hyd = cjd ? __ldg(m.hy + j - 1) : 0.0;
  const double cfz = 0.5 * rx * ry;

  // shared-memory addressing: thread part (column) + warp-uniform part (row, array, slot, lead bit of the stage)
  const double *tcol = tiles + (i - cs + 2);
  const int rowX = (r + 1) * PX, rowN = (r + 1) * PN;
  const int l0 = (int)(a0[0] & 1), l1 = (int)(a0[1] & 1), l2 = (int)(a0[2] & 1), l3 = (int)(a0[3] & 1), l4 = (int)(a0[4] & 1);
  const int p0 = (int)(pstr[0] & 1), p1 = (int)(pstr[1] & 1), p2 = (int)(pstr[2] & 1), p3 = (int)(pstr[3] & 1);
  // (mui and sigma share pitch, plane stride and therefore the stride parity p3)

  const long long pEx = (long long)nx * nNy, pEy = (long long)nNx * ny, pEz = (long long)nNx * nNy;

  if (halo) {
    // ---- halo warp: G_FX and G_FZ of cell row j0-1 for the block's bottom row of consumers ----
    double ex0 = 0.0, ex0u = 0.0, ey0 = 0.0, ey0r = 0.0, mzp = 0.0;
    double *xw = xch + s * 64 + lane;
    uint64_t *xw_bar = xready + s;
    for (int qb = 0; qb < nq; qb += S) {
      const unsigned par = (unsigned)(qb / S) & 1u;
#pragma unroll
      for (int u = 0; u < S; ++u) {
        const int q = qb + u;
        if (q >= nq) break;
        const int ku = jfirst + u;   // S is even: the plane parity of slot u is the same in every ring cycle
        mbar_wait(&full[u], par);
        const double *T = tcol + u * G::STAGE;
        const double *tEX = T + (G::O_EX + rowX + ((l0 + p0 * (ku + 1)) & 1));
        const double *tEY = T + (G::O_EY + rowN + ((l1 + p1 * (ku + 1)) & 1));
        const double ex1 = tEX[0], ex1u = tEX[PX];
        const double ey1 = tEY[0], ey1r = tEY[1];
        if (q >= 1) {
          const double *tEZ = T + (G::O_EZ + rowN + ((l2 + p2 * ku) & 1));
          const double *tMU = T + (G::O_MU + rowX + ((l3 + p3 * ku) & 1));
          const int t = q - 1;
          const double hz = s_hz[t], rzh = s_rzh[t];
          const double ez0 = tEZ[0], ezu = tEZ[PN];
          const double mu0 = tMU[0], mul = tMU[-1];
          const double circx = hz * (ezu - ez0) - hy * (ey1 - ey0);
          const double circz = hy * (ey0r - ey0) - hx * (ex0u - ex0);
          const double mz = hz * mu0;
          const double gfz = ((mzp + mz) * cfz) * circz;
          const double gfx = ((hxl * mul + hx * mu0) * (ry * rzh)) * circx;
          mzp = mz;
          double *X = xw + u * (G::XROWS * XS * 64);
          X[0] = gfx;
          X[32] = gfz;
          __syncwarp();
          if (lane == 0) mbar_arrive(xw_bar + u * (G::XROWS * XS));
        }
        ex0 = ex1;
        ex0u = ex1u;
        ey0 = ey1;
        ey0r = ey1r;
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[u]);
      }
    }
    return;
  }

  // ---- consumer warp: row j of strip s ----
  const double vxy4 = 0.25 * hx * hy, vxyd4 = 0.25 * hx * hyd;   // quarter of the x-y cell areas (rows j, j-1)
  // carried state (plane k / layer k-1)
  double ex0 = 0.0, ex0u = 0.0, ey0 = 0.0, ey0r = 0.0;   // ex(i,j), ex(i,j+1), ey(i,j), ey(i+1,j) on plane k
  double gfxp = 0.0, gfyp = 0.0, mzp = 0.0, ap = 0.0, bp = 0.0;

  const int ic = min(max(i, 0), nx - 1), in_ = min(max(i, 0), nNx - 1), jc = min(j, ny - 1), jn = min(j, nNy - 1);
  double *yx = y + ((long long)(kb - 1) * pEx + ic + (long long)nx * jn);
  double *yy = y + (nEx + (long long)(kb - 1) * pEy + in_ + (long long)nNx * jc);
  double *yz = y + (nEx + nEy + (long long)(kb - 1) * pEz + in_ + (long long)nNx * jn);
  const bool lane_out = lane >= 1;
  const bool wx = lane_out && ci && (j < nNy);
  const bool wy = lane_out && (i >= 0) && (i < nNx) && cj;
  const bool wz = lane_out && (i >= 0) && (i < nNx) && (j < nNy);

  double *xw = xch + ((r + 1) * XS + s) * 64 + lane;     // mailbox this warp fills (rows 0 .. R-2)
  const double *xr = xch + (r * XS + s) * 64 + lane;      // mailbox of the row below (or of the halo warp)
  uint64_t *xw_bar = xready + (r + 1) * XS + s;
  uint64_t *xr_bar = xready + r * XS + s;
  const bool publish = r < R - 1;

  for (int qb = 0; qb < nq; qb += S) {
    const unsigned par = (unsigned)(qb / S) & 1u;
#pragma unroll
    for (int u = 0; u < S; ++u) {
      const int q = qb + u;
      if (q >= nq) break;
      const int ku = jfirst + u;
      mbar_wait(&full[u], par);
      const double *T = tcol + u * G::STAGE;
      const double *tEX = T + (G::O_EX + rowX + ((l0 + p0 * (ku + 1)) & 1));
      const double *tEY = T + (G::O_EY + rowN + ((l1 + p1 * (ku + 1)) & 1));
      const double ex1 = tEX[0], ex1u = tEX[PX];
      const double ey1 = tEY[0], ey1r = tEY[1];
      if (q >= 1 && dbg != 1) {
        const int k = jfirst + q;          // this step works on plane k / layer k
        const double *tEZ = T + (G::O_EZ + rowN + ((l2 + p2 * ku) & 1));
        const double *tMU = T + (G::O_MU + rowX + ((l3 + p3 * ku) & 1));
        const double *tSG = T + (G::O_SG + rowX + ((l4 + p3 * ku) & 1));
        const int t = q - 1;
        const double hz = s_hz[t], rzh = s_rzh[t];
        const double ez0 = tEZ[0], ezu = tEZ[PN];
        const double mu0 = tMU[0], mul = tMU[-1];
        const double circx = hz * (ezu - ez0) - hy * (ey1 - ey0);
        const double circz = hy * (ey0r - ey0) - hx * (ex0u - ex0);
        const double mz = hz * mu0;
        const double gfz = ((mzp + mz) * cfz) * circz;
        const double gfx = ((hxl * mul + hx * mu0) * (ry * rzh)) * circx;
        mzp = mz;
        if (publish) {
          double *X = xw + u * (G::XROWS * XS * 64);
          X[0] = gfx;
          X[32] = gfz;
          __syncwarp();
          if (lane == 0) mbar_arrive(xw_bar + u * (G::XROWS * XS));
        }
        const unsigned xpar = (u == 0) ? (par ^ 1u) : par;   // slot 0 carries no fluxes in the first ring cycle
        const bool xok = mbar_test_wait(xr_bar + u * (G::XROWS * XS), xpar);   // probe now, consume after the y-face work
        const double ezr = tEZ[1], mud = tMU[-PX];
        const double sg0 = tSG[0], sgd = tSG[-PX];
        const double circy = hx * (ex1 - ex0) - hz * (ezr - ez0);
        const double gfy = ((hyd * mud + hy * mu0) * (rx * rzh)) * circy;
        const double gfzl = __shfl_up_sync(0xffffffffu, gfz, 1);
        const double gfyl = __shfl_up_sync(0xffffffffu, gfy, 1);
        const double sp = sg0 * vxy4, sdp = sgd * vxyd4;
        const double a = hz * (sp + sdp);
        const double slp = __shfl_up_sync(0xffffffffu, sp, 1);
        const double b = hz * (sp + slp);
        const double al = __shfl_up_sync(0xffffffffu, a, 1);
        if (!xok) mbar_wait(xr_bar + u * (G::XROWS * XS), xpar);
        const double *X = xr + u * (G::XROWS * XS * 64);
        const double gfxd = X[0], gfzd = X[32];
        const double vx = hx * ((gfyp - gfy) + (gfz - gfzd)) + (ap + a) * ex0;
        const double vy = hy * ((gfx - gfxp) + (gfzl - gfz)) + (bp + b) * ey0;
        const double vz = hz * ((gfxd - gfx) + (gfy - gfyl)) + (a + al) * ez0;
        gfyp = gfy;
        gfxp = gfx;
        ap = a;
        bp = b;
        if (k >= kb) {
          if (wx) *yx = vx;
          if (wy) *yy = vy;
          if (wz && k < nz) *yz = vz;
        }
        yx += pEx;
        yy += pEy;
        yz += pEz;
      }
      ex0 = ex1;
      ex0u = ex1u;
      ey0 = ey1;
      ey0r = ey1r;
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[u]);
    }
  }
}

}  // namespace dzb
