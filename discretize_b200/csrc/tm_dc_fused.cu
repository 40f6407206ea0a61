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
//     registers and prefetching the next one, so every phi and sigma value is loaded
//     once per strip and every output written once: algorithmic bytes = 8 (2 nN + nC),
//   * loads are 256-byte coalesced row segments; out-of-range rows / columns / planes
//     are handled by CLAMPING the address (always a valid, finite entry) and zeroing
//     the cell measure (hx, hy or hz = 0), so that the inner loop carries no masks:
//     a cell outside the mesh has V sigma = 0 and an edge without cells has weight 0.
#include "bulk_pipe.cuh"
#include "tm_common.cuh"

namespace dzb {

constexpr int DC_WARPS = 4;      // warps per block, consecutive x strips
constexpr int DC_XOUT = 30;      // outputs per warp strip (lanes 1..30)

template <int RY, int MINB>
__global__ void __launch_bounds__(DC_WARPS * 32, MINB)
    k_dc_nodal(TM m, const double *__restrict__ sigma, const double *__restrict__ phi, double *__restrict__ y,
               int kchunk, int k_begin, int k_end) {
  // per-block row constants (all warps of a block share j0): cell rows j0-1 .. j0+RY-1
  __shared__ double s_hy[RY + 1], s_qy[RY + 1];
  const int j0 = blockIdx.y * RY;  // first output row
  if (threadIdx.x < RY + 1) {
    const int jc = j0 - 1 + (int)threadIdx.x;
    const bool ok = (jc >= 0) && (jc < m.ny);
    const double r = ok ? m.rhy[jc] : 0.0;
    s_hy[threadIdx.x] = ok ? m.hy[jc] : 0.0;
    s_qy[threadIdx.x] = 0.25 * r * r;
  }
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nx = (int)m.nx, ny = (int)m.ny, nz = (int)m.nz;
  const int nNx = nx + 1, nNy = ny + 1, nNz = nz + 1;
  const int X0 = (blockIdx.x * DC_WARPS + warp) * DC_XOUT;
  if (X0 >= nNx) return;
  const int i = X0 + lane - 1;          // node column of this lane (lane 0 / 31 are halo)
  const int kb = k_begin + blockIdx.z * kchunk;
  const int ke = min(kb + kchunk, k_end);

  const bool node_ok = (i >= 0) && (i < nNx);
  const bool cell_ok = (i >= 0) && (i < nx) && (lane < 31);  // lane 31's cells are never needed
  const int ic = min(max(i, 0), nNx - 1), icc = min(max(i, 0), nx - 1);
  const double hx = cell_ok ? __ldg(m.hx + icc) : 0.0;
  const double rx = cell_ok ? __ldg(m.rhx + icc) : 0.0;
  const double qx = 0.25 * rx * rx;

  // clamped in-plane offsets (elements) of the rows this lane reads
  int offp[RY + 2], offc[RY + 1];
#pragma unroll
  for (int r = 0; r < RY + 2; ++r) offp[r] = ic + nNx * min(max(j0 - 1 + r, 0), nNy - 1);
#pragma unroll
  for (int r = 0; r < RY + 1; ++r) offc[r] = icc + nx * min(max(j0 - 1 + r, 0), ny - 1);

  const i64 pstride = (i64)nNx * nNy;     // node plane
  const i64 cstride = (i64)nx * ny;       // cell layer

  double ph0[RY + 2], ph1[RY + 2], ph2[RY + 2];  // phi planes, roles rotate: k, k+1, (prefetched) k+2
  double sgA[RY + 1], sgB[RY + 1], sgC[RY + 1];  // sigma layers, roles rotate: k, (prefetched) k+1
  double Ap[RY], Bp[RY + 1];                      // layer k-1: y-pair / x-pair sums of V sigma
  double fzp[RY];                                 // z-flux between planes k-1 and k

#define DC_LOAD_PLANE(dst, base)                                    \
  _Pragma("unroll") for (int r = 0; r < RY + 2; ++r) dst[r] = __ldg((base) + offp[r]);
#define DC_LOAD_LAYER(dst, base)                                    \
  _Pragma("unroll") for (int r = 0; r < RY + 1; ++r) dst[r] = __ldg((base) + offc[r]);

  // ---- preamble: plane kb-1 / layer kb-1 (clamped to 0 with zero thickness at the bottom) --------
  {
    const bool have = kb >= 1;
    const int km = have ? kb - 1 : 0;
    DC_LOAD_PLANE(ph1, phi + km * pstride);  // ph1 temporarily holds plane kb-1
    DC_LOAD_PLANE(ph0, phi + kb * pstride);
    DC_LOAD_LAYER(sgA, sigma + km * cstride);
    const double hz = have ? __ldg(m.hz + km) : 0.0;
    const double rz = have ? __ldg(m.rhz + km) : 0.0;
    const double qz = 0.25 * rz * rz;
    const double hxz = hx * hz;
    double s[RY + 1], sl[RY + 1];
#pragma unroll
    for (int r = 0; r < RY + 1; ++r) {
      s[r] = (sgA[r] * s_hy[r]) * hxz;
      sl[r] = __shfl_up_sync(0xffffffffu, s[r], 1);
      Bp[r] = sl[r] + s[r];
    }
#pragma unroll
    for (int r = 0; r < RY; ++r) {
      const double a = s[r] + s[r + 1];
      const double al = sl[r] + sl[r + 1];
      Ap[r] = a;
      fzp[r] = (qz * (a + al)) * (ph0[r + 1] - ph1[r + 1]);
    }
  }
  // plane kb+1 and layer kb (clamped), then running pointers for the prefetches
  const double *pnext = phi + (i64)min(kb + 1, nNz - 1) * pstride;
  const double *cnext = sigma + (i64)min(kb, nz - 1) * cstride;
  DC_LOAD_PLANE(ph1, pnext);
  DC_LOAD_LAYER(sgA, cnext);
  double *py = y + ((i64)kb * pstride + ic + (i64)nNx * j0);
  const bool writer = (lane >= 1) && (lane <= DC_XOUT) && node_ok;

  // ---- march ------------------------------------------------------------------------------
  // One z-step.  P0/P1 = phi on planes k, k+1; P2 receives the prefetch of plane k+2;
  // S0 = sigma on layer k; S1 receives the prefetch of layer k+1.  The caller rotates the
  // roles (period 3) so that no register-to-register copies are needed.
  auto step = [&](const double(&P0)[RY + 2], const double(&P1)[RY + 2], double(&P2)[RY + 2],
                  const double(&S0)[RY + 1], double(&S1)[RY + 1], int k) __attribute__((always_inline)) {
    if (k + 2 < nNz) pnext += pstride;   // clamped at the top plane / layer
    if (k + 1 < nz) cnext += cstride;
    DC_LOAD_PLANE(P2, pnext);
    DC_LOAD_LAYER(S1, cnext);

    const bool havec = k < nz;        // layer k exists
    const int kc = havec ? k : nz - 1;
    const double hz = havec ? __ldg(m.hz + kc) : 0.0;
    const double rz = havec ? __ldg(m.rhz + kc) : 0.0;
    const double qz = 0.25 * rz * rz;
    const double hxz = hx * hz;

    double s[RY + 1], sl[RY + 1], fy[RY + 1];
#pragma unroll
    for (int r = 0; r < RY + 1; ++r) {
      s[r] = (S0[r] * s_hy[r]) * hxz;                 // V sigma of cell (i, j0-1+r, k); 0 outside the mesh
      sl[r] = __shfl_up_sync(0xffffffffu, s[r], 1);   // same for cell column i-1
      const double b = sl[r] + s[r];
      // y-edge between node rows (j0-1+r) and (j0+r), bordered by cell row j0-1+r
      fy[r] = (s_qy[r] * (Bp[r] + b)) * (P0[r + 1] - P0[r]);
      Bp[r] = b;
    }
#pragma unroll
    for (int r = 0; r < RY; ++r) {
      const double a = s[r] + s[r + 1];
      const double al = sl[r] + sl[r + 1];
      // x-edge from this lane's node to the next one, node row j0+r
      const double phr = __shfl_down_sync(0xffffffffu, P0[r + 1], 1);
      const double fx = (qx * (Ap[r] + a)) * (phr - P0[r + 1]);
      const double fxl = __shfl_up_sync(0xffffffffu, fx, 1);
      Ap[r] = a;
      // z-edge from plane k to k+1
      const double fz = (qz * (a + al)) * (P1[r + 1] - P0[r + 1]);
      const double out = (fxl - fx) + (fy[r] - fy[r + 1]) + (fzp[r] - fz);
      fzp[r] = fz;
      if (writer && (j0 + r) < nNy) py[r * nNx] = out;
    }
    py += pstride;
  };

  int k = kb;
  while (true) {
    step(ph0, ph1, ph2, sgA, sgB, k);
    if (++k >= ke) break;
    step(ph1, ph2, ph0, sgB, sgC, k);
    if (++k >= ke) break;
    step(ph2, ph0, ph1, sgC, sgA, k);
    if (++k >= ke) break;
  }
#undef DC_LOAD_PLANE
#undef DC_LOAD_LAYER
}

// ---------------------------------------------------------------------------------------------
// Variant "ring" (default): same math, but the phi planes / sigma layers are staged in shared
// memory by the TMA engine (cp.async.bulk -> SASS UBLKCP; one instruction per row segment,
// issued by a dedicated producer warp) in a DC_SLOTS-deep mbarrier ring, warp-specialised:
//
//   producer warp : for stage j = kb-2 .. ke-1: wait for the slot to be released, then issue one
//                   bulk copy per in-range row of node plane j+1 (RY+2 rows) and cell layer j
//                   (RY+1 rows); completion is counted in bytes on full[slot].
//   consumer warps: z-step j waits on full[slot(j)], reads P1 = plane j+1 and sigma layer j from
//                   stage j and P0 = plane j from stage j-1 with LDS, computes the fluxes (x
//                   neighbours by warp shuffle), writes plane j, and releases stage j-1.
//
// A block spans up to DC_MAXW strips of 30 node columns (the whole row for nx <= 538), so the
// consumers carry no global address arithmetic and no register prefetch buffers, and two stages
// (~90 KB) of loads are in flight per SM independent of occupancy.
//
// Bulk copies need 16-byte aligned source, destination and size, but rows of the reference's
// layout start at arbitrary 8-byte offsets ((n+1)*8-byte pitch).  Each row copy therefore starts
// at the aligned-down address; with a shared-memory row pitch of the same parity as the global
// pitch, every row of a stage lands at  base + lead + r*pitch + c  with ONE lead bit per stage.
// Rows / planes outside the mesh are not copied at all: the tiles are zero-initialised, hold only
// finite values, and everything outside the mesh is multiplied by a zero cell measure.
// The copies never leave the arrays: where an array starts / ends on an odd element address, the one element a 16-byte
// aligned copy cannot reach is staged by a plain load + shared store of the producer lane (bulk_pipe.cuh: plan_row_copy).
// ---------------------------------------------------------------------------------------------
constexpr int DC_SLOTS = 4;       // ring depth: stage j-1 and j in use, j+1 and j+2 in flight
constexpr int DC_MAXCHUNK = 128;  // max z-planes per block (size of the hz table)

// compile-time geometry of a block with NWC consumer warps; PARX = nx & 1
template <int RY, int NWC, int PARX> struct DcRing {
  static constexpr int RP = RY + 2, RC = RY + 1;
  static constexpr int NCP = NWC * DC_XOUT + 2, NCC = NWC * DC_XOUT + 1;  // staged node / cell columns
  // shared-memory row pitches (doubles), parity-matched to nNx = nx+1 and nx
  static constexpr int PP = (NCP + 3) + ((((NCP + 3) & 1) != ((PARX + 1) & 1)) ? 1 : 0);
  static constexpr int CP = (NCC + 3) + ((((NCC + 3) & 1) != PARX) ? 1 : 0);
  static constexpr int REGP = (RP * PP + 4 + 1) & ~1, REGC = (RC * CP + 4 + 1) & ~1;
  static constexpr int STAGE = REGP + REGC;
  static constexpr size_t SMEM = (size_t)DC_SLOTS * STAGE * 8 + 2 * DC_SLOTS * 8 + 2 * RC * 8 + 2 * (DC_MAXCHUNK + 2) * 8;
};

// Peer-read halos (z-slabs, one process per GPU): the node planes lo_plane / hi_plane of the local vector are not read
// from it but from a neighbour's vector mapped into this process (NVLink): *_addr is the ELEMENT address (address / 8) of
// that plane in the neighbour's array, [*_beg, *_end) the extent of that array.  plane < 0: no such neighbour.
//
// Ordering against the neighbours' writes, inside the kernel (sync != NULL; else the caller puts a barrier in front):
// sync[0] / sync[1] are set by the lower / upper neighbour, sync[2] counts this rank's completed launches, sync[3] the
// blocks of the running launch that have finished, sync[4] != 0 reports a wait that timed out.  A launch belongs to
// epoch e = sync[2] + 1.  Its first block tells both neighbours "my vector of epoch e is complete" (true by stream
// order: the launch follows whatever produced phi) by writing e to their sync words; a producer warp that is about to
// copy a halo plane waits until the neighbour it comes from has announced epoch e.  Nobody waits for anything that
// depends on another rank's progress through its own launch, so the ranks cannot block each other; every rank must apply
// the operator the same number of times.
struct DcHalo {
  int lo_plane, hi_plane;
  long long lo_addr, lo_beg, lo_end, hi_addr, hi_beg, hi_end;
  unsigned *sync, *lo_signal, *hi_signal;
};

template <int RY, int NWC, int PARX, bool PEER, bool DOT>
__global__ void __launch_bounds__((NWC + 1) * 32, 1)
    k_dc_nodal_ring(TM m, const double *__restrict__ sigma, const double *__restrict__ phi_all, double *__restrict__ y_all,
                    int kchunk, int nwc, int k_begin, int k_end, int nrhs, long long ldx, DcHalo hal,
                    double *__restrict__ dotp) {
  typedef DcRing<RY, NWC, PARX> G;
  // multi-RHS launch: blockIdx.x = x-block * nrhs + rhs -- the right-hand side is the FASTEST index of the launch order, so
  // that the nrhs blocks that stage the same sigma tile run side by side and all but the first find it in L2 (sigma is
  // 1/3 of the traffic of one apply)
  const int rhs = (int)(blockIdx.x % (unsigned)nrhs), xblock = (int)(blockIdx.x / (unsigned)nrhs), zchunk = (int)blockIdx.z;
  const double *__restrict__ phi = phi_all + (long long)rhs * ldx;
  double *__restrict__ y = y_all + (long long)rhs * ldx;
  constexpr int RP = G::RP, RC = G::RC, S = DC_SLOTS, PP = G::PP, CP = G::CP;
  static_assert(S == 4, "the consumer loop is unrolled over a ring of four slots");
  extern __shared__ __align__(128) unsigned char dc_smem_raw[];
  const int nx = (int)m.nx, ny = (int)m.ny, nz = (int)m.nz;
  const int nNx = nx + 1, nNy = ny + 1, nNz = nz + 1;
  double *tiles = reinterpret_cast<double *>(dc_smem_raw);               // [S][STAGE]
  uint64_t *full = reinterpret_cast<uint64_t *>(tiles + S * G::STAGE);
  uint64_t *empty = full + S;
  double *s_hy = reinterpret_cast<double *>(empty + S);
  double *s_qy = s_hy + RC;
  double *s_hz = s_qy + RC;                 // [kchunk+1]: layers kb-1 .. ke-1 (0 outside the mesh)
  double *s_qz = s_hz + (DC_MAXCHUNK + 2);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nthreads = (nwc + 1) * 32;
  const int j0 = blockIdx.y * RY;
  const int Xb = xblock * (nwc * DC_XOUT);
  const int kb = k_begin + zchunk * kchunk;
  const int ke = min(kb + kchunk, k_end);
  const int jfirst = kb - 2;  // first stage of this block

  for (int t = tid; t < S * G::STAGE; t += nthreads) tiles[t] = 0.0;  // tiles only ever hold finite values
  if (tid < RC) {
    const int jc = j0 - 1 + tid;
    const bool ok = (jc >= 0) && (jc < ny);
    const double r = ok ? m.rhy[jc] : 0.0;
    s_hy[tid] = ok ? m.hy[jc] : 0.0;
    s_qy[tid] = 0.25 * r * r;
  }
  for (int t = tid; t <= ke - kb; t += nthreads) {
    const int k = kb - 1 + t;
    const bool ok = (k >= 0) && (k < nz);
    const double r = ok ? m.rhz[k] : 0.0;
    s_hz[t] = ok ? m.hz[k] : 0.0;
    s_qz[t] = 0.25 * r * r;
  }
  if (tid == 0) {
    for (int q = 0; q < S; ++q) {
      mbar_init(&full[q], 1);
      mbar_init(&empty[q], nwc);
    }
    fence_barrier_init();
  }
  unsigned epoch = 0;
  if (PEER && hal.sync) {
    epoch = *reinterpret_cast<volatile unsigned *>(hal.sync + 2) + 1;   // changes only after ALL blocks have passed here
    if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
      __threadfence_system();
      if (hal.lo_signal) signal_release_sys(hal.lo_signal, epoch);
      if (hal.hi_signal) signal_release_sys(hal.hi_signal, epoch);
    }
  }
  fence_proxy_async_smem();
  __syncthreads();
  if (PEER && hal.sync && tid == 0) {   // the last block to get here closes the epoch
    const unsigned total = gridDim.x * gridDim.y * gridDim.z;
    if (atomicAdd(hal.sync + 3, 1u) == total - 1) {
      hal.sync[3] = 0;
      __threadfence();
      *reinterpret_cast<volatile unsigned *>(hal.sync + 2) = epoch;
    }
  }

  const long long pstride = (long long)nNx * nNy, cstride = (long long)nx * ny;
  const int cs = max(Xb - 1, 0);  // first staged column (nodes and cells)
  // absolute element address (address / 8) of (column cs, virtual row j0-1, plane 0)
  const long long pa0 = (long long)((unsigned long long)phi >> 3) + cs + (long long)nNx * (j0 - 1);
  const long long ca0 = (long long)((unsigned long long)sigma >> 3) + cs + (long long)nx * (j0 - 1);
  const int nq = ke - jfirst;  // stages / steps are numbered q = j - jfirst

  if (warp == nwc) {
    // ================= producer warp =================
    // Lane l owns row copy l of every stage (RP phi rows, then RC sigma rows): the address arithmetic runs
    // SIMT across the lanes, lane 0 posts the stage's byte count, then every lane issues its bulk copy.
    {
      const bool isp = lane < RP;
      const bool have = lane < RP + RC;
      const int r = isp ? lane : lane - RP;
      const int jr = j0 - 1 + r;
      const bool row_ok = have && jr >= 0 && jr < (isp ? nNy : ny);
      const long long a0 = isp ? pa0 : ca0, pstr = isp ? pstride : cstride;
      const long long rowbase = a0 + (long long)r * (isp ? nNx : nx);
      // exact extent of the array this lane copies from: nothing outside [abeg, aend) is ever read (plan_row_copy)
      const long long abeg = isp ? (long long)((unsigned long long)phi >> 3) : (long long)((unsigned long long)sigma >> 3);
      const long long aend = abeg + (isp ? pstride * nNz : cstride * nz);
      const int ncol = nwc * DC_XOUT + (isp ? 2 : 1);
      const int npl = isp ? nNz : nz, ploff = isp ? 1 : 0;
      const int soff = (isp ? 0 : G::REGP) + 2 + r * (isp ? PP : CP);
      for (int q = 0; q < nq; ++q) {
        const int jst = jfirst + q, slot = q % S, n = q / S;
        if (n >= 1) mbar_wait(&empty[slot], (n - 1) & 1);
        const int pl = jst + ploff;
        const bool on = row_ok && pl >= 0 && pl < npl;
        long long lo = rowbase + pstr * pl, lead_a = a0 + pstr * pl, ab = abeg, ae = aend;
        if (PEER && isp && (pl == hal.lo_plane || pl == hal.hi_plane)) {   // this plane lives in a neighbour's memory
          const bool low = pl == hal.lo_plane;
          const long long pb = low ? hal.lo_addr : hal.hi_addr;
          lead_a = pb + cs + (long long)nNx * (j0 - 1);
          lo = lead_a + (long long)r * nNx;
          ab = low ? hal.lo_beg : hal.hi_beg;
          ae = low ? hal.lo_end : hal.hi_end;
          if (hal.sync && on) {   // the neighbour's vector of this epoch must be complete
            wait_signal_sys(hal.sync + (low ? 0 : 1), epoch, hal.sync + 4);
            fence_proxy_async_global();
          }
        }
        RowCopy rc = plan_row_copy(lo, lo + ncol, ab, ae);
        if (!on) {
          rc.bytes = 0;
          rc.head = rc.tail = false;
        }
        double *T = tiles + slot * G::STAGE;
        const int dst0 = soff + (int)(lead_a & 1);   // element lo lands here: one lead bit per stage
        const unsigned total = __reduce_add_sync(0xffffffffu, rc.bytes);
        if (__any_sync(0xffffffffu, rc.head | rc.tail)) {   // only the rows that hold the first / last element of an odd-aligned array
          if (rc.head) T[dst0] = *reinterpret_cast<const double *>(lo << 3);
          if (rc.tail) T[dst0 + (int)(rc.hi - 1 - lo)] = *reinterpret_cast<const double *>((rc.hi - 1) << 3);
          fence_proxy_async_smem();
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_expect_tx(&full[slot], total);
        __syncwarp();
        if (rc.bytes)
          bulk_g2s(T + dst0 + (int)(rc.tlo - lo), reinterpret_cast<const void *>(rc.tlo << 3), rc.bytes, &full[slot]);
      }
    }
    return;
  }

  // ================= consumer warps =================
  const int i = Xb + warp * DC_XOUT + lane - 1;  // node column of this lane (lane 0 / 31 are halo)
  const bool node_ok = (i >= 0) && (i < nNx);
  const bool cell_ok = (i >= 0) && (i < nx) && (lane < 31);
  const int icc = min(max(i, 0), nx - 1);
  const double hx = cell_ok ? __ldg(m.hx + icc) : 0.0;
  const double rx = cell_ok ? __ldg(m.rhx + icc) : 0.0;
  const double qx = 0.25 * rx * rx;
  const int c2 = i - cs + 2;  // shared-memory column of this lane for lead 0 (>= 1)
  const bool writer = (lane >= 1) && (lane <= DC_XOUT) && node_ok;
  const int pbP = (int)(pstride & 1), pbC = (int)(cstride & 1);
  const int l0p = (int)(pa0 & 1), l0c = (int)(ca0 & 1);
  unsigned wmask = 0;  // output rows that exist
#pragma unroll
  for (int r = 0; r < RY; ++r)
    if (writer && (j0 + r) < nNy) wmask |= 1u << r;

  double Ap[RY], Bp[RY + 1], fzp[RY];
  double dacc = 0.0;   // dotp: this warp's part of phi . (A phi), the p.Ap of a conjugate-gradient iteration
#pragma unroll
  for (int r = 0; r < RY; ++r) Ap[r] = fzp[r] = 0.0;
#pragma unroll
  for (int r = 0; r < RY + 1; ++r) Bp[r] = 0.0;

  double *py = y + ((long long)(kb - 1) * pstride + min(max(i, 0), nNx - 1) + (long long)nNx * j0);

  // steps q = 1 .. nq-1 (z-plane j = jfirst + q); slots are static inside the 4x unrolled body
  for (int qb = 0; qb < nq; qb += S) {
    const unsigned par = (unsigned)(qb / S) & 1u;
#pragma unroll
    for (int u = 0; u < S; ++u) {
      const int q = qb + u;
      if (q == 0) continue;
      if (q >= nq) break;
      const int j = jfirst + q;
      constexpr int dummy = 0;
      (void)dummy;
      const int slot = u, slot_prev = (u + S - 1) % S;
      if (q == 1) mbar_wait(&full[0], 0);  // stage 0 only carries plane kb-1 (read below as P0)
      mbar_wait(&full[slot], par);
      // lead bits of the planes j (stage j-1), j+1 (stage j) and of layer j (stage j)
      int lp0 = (l0p + pbP * j) & 1, lp1 = (l0p + pbP * (j + 1)) & 1;
      const int lc = (l0c + pbC * j) & 1;
      if (PEER) {   // a plane fetched from a neighbour carries the lead bit of its address there
        const long long rel = cs + (long long)nNx * (j0 - 1);
        if (j == hal.lo_plane) lp0 = (int)((hal.lo_addr + rel) & 1);
        if (j == hal.hi_plane) lp0 = (int)((hal.hi_addr + rel) & 1);
        if (j + 1 == hal.lo_plane) lp1 = (int)((hal.lo_addr + rel) & 1);
        if (j + 1 == hal.hi_plane) lp1 = (int)((hal.hi_addr + rel) & 1);
      }
      const double *tP0 = tiles + slot_prev * G::STAGE + (c2 + lp0);
      const double *tP1 = tiles + slot * G::STAGE + (c2 + lp1);
      const double *tS = tiles + slot * G::STAGE + G::REGP + (c2 + lc);
      const double hz = s_hz[j - (kb - 1)], qz = s_qz[j - (kb - 1)];
      const double hxz = hx * hz;

      double P0[RP], s[RC], sl[RC], fy[RC];
#pragma unroll
      for (int r = 0; r < RP; ++r) P0[r] = tP0[r * PP];
#pragma unroll
      for (int r = 0; r < RC; ++r) {
        s[r] = (tS[r * CP] * s_hy[r]) * hxz;            // V sigma of cell (i, j0-1+r, j); 0 outside the mesh
        sl[r] = __shfl_up_sync(0xffffffffu, s[r], 1);   // same for cell column i-1
        const double b = sl[r] + s[r];
        fy[r] = (s_qy[r] * (Bp[r] + b)) * (P0[r + 1] - P0[r]);
        Bp[r] = b;
      }
      const unsigned wm = (j >= kb) ? wmask : 0u;
#pragma unroll
      for (int r = 0; r < RY; ++r) {
        const double a = s[r] + s[r + 1];
        const double al = sl[r] + sl[r + 1];
        const double phr = tP0[(r + 1) * PP + 1];
        const double ph1 = tP1[(r + 1) * PP];
        const double fx = (qx * (Ap[r] + a)) * (phr - P0[r + 1]);
        const double fxl = __shfl_up_sync(0xffffffffu, fx, 1);
        Ap[r] = a;
        const double fz = (qz * (a + al)) * (ph1 - P0[r + 1]);
        const double out = (fxl - fx) + (fy[r] - fy[r + 1]) + (fzp[r] - fz);
        fzp[r] = fz;
        if ((wm >> r) & 1u) {
          py[r * nNx] = out;
          if (DOT) dacc += out * P0[r + 1];
        }
      }
      py += pstride;
      // the plane of stage j-1 is dead now: hand its slot back to the producer
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[slot_prev]);
    }
  }
  if (DOT) {   // one partial per (block, consumer warp), summed in index order by k_sum_partials: deterministic
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dacc += __shfl_xor_sync(0xffffffffu, dacc, o);
    if (lane == 0) {
      const long long b = ((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
      dotp[b * nwc + warp] = dacc;
    }
  }
}

// out[0] = sum of partials[0 .. n) in index order (one block)
__global__ void __launch_bounds__(256) k_sum_partials(long long n, const double *__restrict__ partials, double *out) {
  __shared__ double sm[256];
  double v = 0.0;
  for (long long t = threadIdx.x; t < n; t += 256) v += partials[t];
  sm[threadIdx.x] = v;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sm[0];
}

static int g_dc_ry = 4;
static int g_dc_kchunk = 64;
static int g_dc_variant = 1;  // 0 = register marching, 1 = bulk-copy staged ring

}  // namespace dzb

using namespace dzb;

// tuning knobs for experiments (not part of the stable ABI)
extern "C" void dzb_tune_dc(int ry, int kchunk) {
  if (ry == 2 || ry == 4 || ry == 8) g_dc_ry = ry;
  if (kchunk > 0) g_dc_kchunk = kchunk;
}
extern "C" void dzb_tune_dc_variant(int variant) { g_dc_variant = variant; }

template <int RY, int NWC, int PARX, bool PEER = false>
static int launch_dc_ring_t(const TM &m, const double *sigma, const double *phi, double *y, int kchunk, int xblocks,
                            int nwc, int k_begin, int k_end, cudaStream_t s, int nrhs = 1, long long ldx = 0,
                            const DcHalo *hal = nullptr, double *dotp = nullptr, long long dot_cap = 0,
                            long long *n_partials = nullptr) {
  typedef DcRing<RY, NWC, PARX> G;
  // the attribute is per device (a process may drive several GPUs); setting it is cheap, so no process-wide flag
  if (cudaFuncSetAttribute(k_dc_nodal_ring<RY, NWC, PARX, PEER, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)G::SMEM) != cudaSuccess)
    return -1;
  if (!PEER && cudaFuncSetAttribute(k_dc_nodal_ring<RY, NWC, PARX, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)G::SMEM) != cudaSuccess)
    return -1;
  const DcHalo none{-1, -1, 0, 0, 0, 0, 0, 0, nullptr, nullptr, nullptr};
  const int nNy = (int)m.ny + 1;
  // z-chunk: one block per SM is resident (192 KB of stages), so a launch runs in waves of `sms` blocks and a block
  // costs (planes + 2 halo stages + ~3 stages of start-up) stage times; pick the split that minimises waves x cost.
  // (A 64-plane slab of 512^3 on 8 GPUs: 129 blocks of 65 planes in ONE wave instead of 645 blocks of 13 planes in
  // five, each of them staging 15 planes.)
  const int nxy = xblocks * ((nNy + RY - 1) / RY) * nrhs;
  {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int planes = k_end - k_begin, maxc = kchunk > DC_MAXCHUNK ? DC_MAXCHUNK : kchunk + kchunk / 4;
    long long best = -1;
    int best_kc = planes < maxc ? planes : maxc;
    for (int zb = (planes + maxc - 1) / maxc; zb <= planes && zb <= 256; ++zb) {
      const int kc = (planes + zb - 1) / zb;
      const long long blocks = (long long)nxy * ((planes + kc - 1) / kc);
      const long long cost = ((blocks + sms - 1) / sms) * (kc + 5);
      if (best < 0 || cost < best) {
        best = cost;
        best_kc = kc;
      }
    }
    kchunk = best_kc;
  }
  dim3 grid((unsigned)xblocks * nrhs, (unsigned)((nNy + RY - 1) / RY), (unsigned)((k_end - k_begin + kchunk - 1) / kchunk));
  if (dotp) {
    const long long need = (long long)grid.x * grid.y * grid.z * nwc;
    if (need > dot_cap) return -3;          // workspace too small: the caller takes a separate dot kernel
    if (n_partials) *n_partials = need;
  }
  if (dotp && !PEER)   // the variant that also accumulates phi . y (kept apart: it costs the plain kernel 64 bytes of spills)
    k_dc_nodal_ring<RY, NWC, PARX, false, true><<<grid, dim3((nwc + 1) * 32), G::SMEM, s>>>(m, sigma, phi, y, kchunk, nwc,
                                                                                          k_begin, k_end, nrhs, ldx, none, dotp);
  else
    k_dc_nodal_ring<RY, NWC, PARX, PEER, false><<<grid, dim3((nwc + 1) * 32), G::SMEM, s>>>(m, sigma, phi, y, kchunk, nwc,
                                                                                          k_begin, k_end, nrhs, ldx,
                                                                                          hal ? *hal : none, nullptr);
  return 0;
}

constexpr int DC_NWC_BIG = 18, DC_NWC_SMALL = 6;

static int launch_dc_ring(const TM &m, const double *sigma, const double *phi, double *y, int kchunk, int k_begin,
                          int k_end, cudaStream_t s, int nrhs = 1, long long ldx = 0, const DcHalo *hal = nullptr,
                          double *dotp = nullptr, long long dot_cap = 0, long long *n_partials = nullptr) {
  const int nNx = (int)m.nx + 1;
  const int strips = (nNx + DC_XOUT - 1) / DC_XOUT;
  const bool small = strips <= DC_NWC_SMALL;
  const int maxw = small ? DC_NWC_SMALL : DC_NWC_BIG;
  const int xblocks = (strips + maxw - 1) / maxw;
  const int nwc = (strips + xblocks - 1) / xblocks;  // balanced consumer warps per block
  kchunk = kchunk > DC_MAXCHUNK ? DC_MAXCHUNK : kchunk;
  const int parx = (int)(m.nx & 1);
  if (hal) {
    if (small)
      return parx ? launch_dc_ring_t<4, DC_NWC_SMALL, 1, true>(m, sigma, phi, y, kchunk, xblocks, nwc, k_begin, k_end, s, 1, 0, hal)
                  : launch_dc_ring_t<4, DC_NWC_SMALL, 0, true>(m, sigma, phi, y, kchunk, xblocks, nwc, k_begin, k_end, s, 1, 0, hal);
    return parx ? launch_dc_ring_t<4, DC_NWC_BIG, 1, true>(m, sigma, phi, y, kchunk, xblocks, nwc, k_begin, k_end, s, 1, 0, hal)
                : launch_dc_ring_t<4, DC_NWC_BIG, 0, true>(m, sigma, phi, y, kchunk, xblocks, nwc, k_begin, k_end, s, 1, 0, hal);
  }
  if (small)
    return parx ? launch_dc_ring_t<4, DC_NWC_SMALL, 1>(m, sigma, phi, y, kchunk, xblocks, nwc, k_begin, k_end, s, nrhs, ldx, nullptr, dotp, dot_cap, n_partials)
                : launch_dc_ring_t<4, DC_NWC_SMALL, 0>(m, sigma, phi, y, kchunk, xblocks, nwc, k_begin, k_end, s, nrhs, ldx, nullptr, dotp, dot_cap, n_partials);
  return parx ? launch_dc_ring_t<4, DC_NWC_BIG, 1>(m, sigma, phi, y, kchunk, xblocks, nwc, k_begin, k_end, s, nrhs, ldx, nullptr, dotp, dot_cap, n_partials)
              : launch_dc_ring_t<4, DC_NWC_BIG, 0>(m, sigma, phi, y, kchunk, xblocks, nwc, k_begin, k_end, s, nrhs, ldx, nullptr, dotp, dot_cap, n_partials);
}

extern "C" int dzb_tm_dc_nodal_range(const DzbTensorMesh *mm, const double *sigma, const double *phi, double *y,
                                     int64_t k_begin, int64_t k_end, void *stream) {
  if (!mm || mm->nx <= 0 || mm->ny <= 0 || mm->nz <= 0) {
    set_error("dzb_tm_dc_nodal: invalid mesh");
    return -1;
  }
  TM m(*mm);
  const i64 nNx = m.nx + 1, nNy = m.ny + 1, nNz = m.nz + 1;
  if (nNx * nNy >= (i64)1 << 31 || nNz >= (i64)1 << 30) {
    set_error("dzb_tm_dc_nodal: a node plane must have fewer than 2^31 entries");
    return -1;
  }
  if (k_begin < 0 || k_end > nNz || k_begin > k_end) {
    set_error("dzb_tm_dc_nodal_range: plane range outside [0, nz+1]");
    return -1;
  }
  if (k_begin == k_end) return 0;
  const int kb0 = (int)k_begin, ke0 = (int)k_end;
  const i64 strips = (nNx + DC_XOUT - 1) / DC_XOUT;
  const int ry = g_dc_ry;
  int kchunk = g_dc_kchunk;
  cudaStream_t s = (cudaStream_t)stream;
  if (g_dc_variant == 1 && m.ny >= 2 && m.nN() < ((i64)1 << 31)) {
    if (launch_dc_ring(m, sigma, phi, y, kchunk, kb0, ke0, s) == 0) return check_launch("dzb_tm_dc_nodal");
    cudaGetLastError();  // fall through to the register variant
  }
  dim3 grid((unsigned)((strips + DC_WARPS - 1) / DC_WARPS), (unsigned)((nNy + ry - 1) / ry),
            (unsigned)((ke0 - kb0 + kchunk - 1) / kchunk));
  dim3 block(DC_WARPS * 32);
  if (ry == 2) k_dc_nodal<2, 6><<<grid, block, 0, s>>>(m, sigma, phi, y, kchunk, kb0, ke0);
  else if (ry == 4) k_dc_nodal<4, 4><<<grid, block, 0, s>>>(m, sigma, phi, y, kchunk, kb0, ke0);
  else k_dc_nodal<8, 2><<<grid, block, 0, s>>>(m, sigma, phi, y, kchunk, kb0, ke0);
  return check_launch("dzb_tm_dc_nodal");
}

// A z-slab's apply with PEER-READ halos: the node planes lo_plane / hi_plane of the local vector are fetched by the
// producer warp from the neighbours' vectors phi_lo / phi_hi (plane lo_peer_plane / hi_peer_plane of arrays with
// nz_lo + 1 / nz_hi + 1 node planes of the same nx, ny): one launch per apply, no send / recv.
extern "C" int dzb_tm_dc_nodal_range_peer(const DzbTensorMesh *mm, const double *sigma, const double *phi,
                                          const double *phi_lo, int64_t nz_lo, int64_t lo_plane, int64_t lo_peer_plane,
                                          const double *phi_hi, int64_t nz_hi, int64_t hi_plane, int64_t hi_peer_plane,
                                          uint32_t *sync, uint32_t *lo_signal, uint32_t *hi_signal, double *y,
                                          int64_t k_begin, int64_t k_end, void *stream) {
  if (!mm || mm->nx <= 0 || mm->ny <= 0 || mm->nz <= 0) {
    set_error("dzb_tm_dc_nodal_range_peer: invalid mesh");
    return -1;
  }
  TM m(*mm);
  const i64 nNx = m.nx + 1, nNy = m.ny + 1, nNz = m.nz + 1, pstride = nNx * nNy;
  if (k_begin < 0 || k_end > nNz || k_begin > k_end) {
    set_error("dzb_tm_dc_nodal_range_peer: plane range outside [0, nz+1]");
    return -1;
  }
  if ((phi_lo && (lo_plane < 0 || lo_plane >= nNz || lo_peer_plane < 0 || lo_peer_plane > nz_lo)) ||
      (phi_hi && (hi_plane < 0 || hi_plane >= nNz || hi_peer_plane < 0 || hi_peer_plane > nz_hi))) {
    set_error("dzb_tm_dc_nodal_range_peer: halo plane outside the local / the neighbour's array");
    return -1;
  }
  if (k_begin == k_end) return 0;
  if (!(m.ny >= 2 && m.nN() < ((i64)1 << 31) && pstride < ((i64)1 << 31))) return -2;   // ring kernel only
  DcHalo hal{-1, -1, 0, 0, 0, 0, 0, 0, sync, sync ? lo_signal : nullptr, sync ? hi_signal : nullptr};
  if (phi_lo) {
    hal.lo_plane = (int)lo_plane;
    hal.lo_beg = (long long)((unsigned long long)phi_lo >> 3);
    hal.lo_end = hal.lo_beg + pstride * (nz_lo + 1);
    hal.lo_addr = hal.lo_beg + pstride * lo_peer_plane;
  }
  if (phi_hi) {
    hal.hi_plane = (int)hi_plane;
    hal.hi_beg = (long long)((unsigned long long)phi_hi >> 3);
    hal.hi_end = hal.hi_beg + pstride * (nz_hi + 1);
    hal.hi_addr = hal.hi_beg + pstride * hi_peer_plane;
  }
  if (launch_dc_ring(m, sigma, phi, y, g_dc_kchunk, (int)k_begin, (int)k_end, (cudaStream_t)stream, 1, 0, &hal) != 0) {
    cudaGetLastError();
    return -2;
  }
  return check_launch("dzb_tm_dc_nodal_range_peer");
}

// y[r] = A phi[r] for nrhs right-hand sides stored one after the other (right-hand side r at phi + r * ldx): ONE launch;
// the blocks of the nrhs right-hand sides that stage the same sigma tile run side by side and share it through L2.
extern "C" int dzb_tm_dc_nodal_multi(const DzbTensorMesh *mm, const double *sigma, const double *phi, double *y, int nrhs,
                                     int64_t ldx, void *stream) {
  if (!mm || mm->nx <= 0 || mm->ny <= 0 || mm->nz <= 0) {
    set_error("dzb_tm_dc_nodal_multi: invalid mesh");
    return -1;
  }
  TM m(*mm);
  if (nrhs < 1 || ldx < m.nN()) {
    set_error("dzb_tm_dc_nodal_multi: nrhs >= 1 and ldx >= number of nodes required");
    return -1;
  }
  const i64 nNx = m.nx + 1, nNy = m.ny + 1, nNz = m.nz + 1;
  const bool ring = g_dc_variant == 1 && m.ny >= 2 && m.nN() < ((i64)1 << 31) && nNx * nNy < ((i64)1 << 31) &&
                    nrhs <= 4096;
  if (ring && launch_dc_ring(m, sigma, phi, y, g_dc_kchunk, 0, (int)nNz, (cudaStream_t)stream, nrhs, ldx) == 0)
    return check_launch("dzb_tm_dc_nodal_multi");
  cudaGetLastError();
  for (int r = 0; r < nrhs; ++r) {
    const int rc = dzb_tm_dc_nodal_range(mm, sigma, phi + (i64)r * ldx, y + (i64)r * ldx, 0, nNz, stream);
    if (rc) return rc;
  }
  return 0;
}

// y = A phi and, in the same launch, dot[0] = phi . y (the p.Ap of a conjugate-gradient iteration: the consumers hold both
// values of every node they write, so the separate 16 B / node dot kernel goes away).  Partials per (block, consumer warp)
// go to `workspace` (capacity in doubles) and are summed in index order by a one-block kernel: deterministic.  Returns -2
// where the ring kernel does not apply or the workspace is too small (use dzb_tm_dc_nodal + a dot kernel there).
extern "C" int dzb_tm_dc_nodal_dot(const DzbTensorMesh *mm, const double *sigma, const double *phi, double *y,
                                   double *workspace, int64_t workspace_doubles, double *dot, void *stream) {
  if (!mm || mm->nx <= 0 || mm->ny <= 0 || mm->nz <= 0) {
    set_error("dzb_tm_dc_nodal_dot: invalid mesh");
    return -1;
  }
  TM m(*mm);
  const i64 nNx = m.nx + 1, nNy = m.ny + 1, nNz = m.nz + 1;
  if (!(g_dc_variant == 1 && m.ny >= 2 && m.nN() < ((i64)1 << 31) && nNx * nNy < ((i64)1 << 31)) || !workspace) return -2;
  long long np = 0;
  if (launch_dc_ring(m, sigma, phi, y, g_dc_kchunk, 0, (int)nNz, (cudaStream_t)stream, 1, 0, nullptr, workspace,
                     (long long)workspace_doubles, &np) != 0) {
    cudaGetLastError();
    return -2;
  }
  k_sum_partials<<<1, 256, 0, (cudaStream_t)stream>>>(np, workspace, dot);
  return check_launch("dzb_tm_dc_nodal_dot");
}

extern "C" int dzb_tm_dc_nodal(const DzbTensorMesh *mm, const double *sigma, const double *phi, double *y,
                               void *stream) {
  if (!mm) {
    set_error("dzb_tm_dc_nodal: invalid mesh");
    return -1;
  }
  return dzb_tm_dc_nodal_range(mm, sigma, phi, y, 0, mm->nz + 1, stream);
}
