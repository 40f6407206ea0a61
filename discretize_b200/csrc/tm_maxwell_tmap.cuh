// K9 "tmap": y = C^T Mf(mui) C e + Me(sigma) e -- the row-per-warp ring of tm_maxwell_rows.cuh fed by tensor-map TMA.
//
// Why: with row-wise bulk copies (UBLKCP) a stage of the ring costs 17-47 copy instructions, each issued from uniform
// registers by ONE warp at ~65 cycles apiece; measured at 1024^3 the copy pipeline alone (no arithmetic) then needs
// 14.2 ms with one producer warp and the arithmetic alone 8.3-9.7 ms, and the two do not overlap (15.4 ms together).
// A tensor map moves a whole (rows x columns) box with one instruction (SASS UTMALDG), needs no 16-byte aligned row
// starts (only a 16-byte aligned base address and row pitch) and zero-fills everything outside the array.
//
// The reference's layouts have row pitches of nx (ex, mui, sigma) and nx+1 (ey, ez) doubles, so one of the two groups has
// an odd pitch = a byte pitch that is not a multiple of 16.  Those arrays are described as matrices of SUPER-ROWS (two
// consecutive rows, pitch 2(nx+1) doubles = a multiple of 16 bytes): the rows J0 .. J0+H-1 a stage needs are fetched as
// two boxes, the even rows (x in [c, c+W)) and the odd rows (x in [P+c, P+c+W)); the producer steers them so that the
// staged rows rr = 0, 2, 4, .. land in sub-tile 0 and rr = 1, 3, .. in sub-tile 1, each at box row rr>>1, whatever the
// parity of J0 (derivation in DESIGN.md).  Every array is one 2-D matrix of
// (rows per plane x planes) rows, because planes follow each other without padding.
//
// Two maps per odd-pitch array keep every read inside the array: one over its even rows only (width = one row, row
// stride = two rows) and one over the super-rows, from which the boxes take the odd rows.  Used when the five base
// addresses are 16-byte aligned (e.g. nx and ny even: all BASELINE configs and their z-slabs); otherwise the launcher
// falls back to the bulk-copy ring.
// The arithmetic, the mailboxes and the halo warps are those of tm_maxwell_rows.cuh (see there for the formulas).
#pragma once
#include <type_traits>
#include <cuda.h>
#include "bulk_pipe.cuh"
#include "tm_common.cuh"

namespace dzb {

__device__ __forceinline__ void tmap_load_2d(void *dst_smem, const CUtensorMap *map, int x, int y, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst_smem)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tmap_load_2d_hint(void *dst_smem, const CUtensorMap *map, int x, int y, uint64_t *bar,
                                                  uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], "
      "[%4], %5;" ::"r"(smem_u32(dst_smem)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tmap_prefetch(const CUtensorMap *map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

constexpr int M3_XOUT = 31;       // output columns per strip: lanes 1..31 (lane 0 = left halo)
constexpr int M3_MAXCHUNK = 160;  // max z-planes per block (size of the hz tables)

// compile-time geometry: XS strips x R rows per block, PARX = nx & 1, NSLOT ring slots
template <int XS, int R, int PARX, int NSLOT> struct Mx3 {
  // NPW producer warps share the copies of a stage (one elected lane each).  Measured at 1024^3 under the power cap: two
  // producers 12.16 ms, one 12.08 ms -- the ~225 instructions one lane spends per stage are not what limits the ring, so
  // one producer warp it is (what did cost 12-20 % in the peer-read variant was selecting the tensor map at run time in
  // EVERY stage; see the producer loop)
  static constexpr int NPW = 1;
  static constexpr int NCONS = XS * R, NHALO = XS, NWARPS = NCONS + NHALO + NPW, NT = NWARPS * 32;
  // producer 0 copies the arrays of mask0, producer 1 the others (balanced in boxes: odd-pitch arrays take two)
  static constexpr bool mine(int a, int pw) { return NPW == 1 || (PARX ? (a <= 2) : (a == 1 || a == 2)) == (pw == 0); }
  static constexpr int RN = R + 2, RC = R + 1;          // staged node rows j0-1 .. j0+R, cell rows j0-1 .. j0+R-1
  // box width: the box starts at the even column <= Xb-1 (the TMA unit wants 16-byte aligned inner coordinates: an odd
  // start raises an illegal-instruction fault, tools/tma_probe) and must reach column Xb+XS*31
  static constexpr int W = (XS * M3_XOUT + 3 + 1) & ~1;
  static constexpr int pad16(int n) { return (n + 15) & ~15; }          // boxes start on 128-byte boundaries
  // array a: 0 ex (pitch nx, RN rows) 1 ey (nx+1, RC) 2 ez (nx+1, RN) 3 mui (nx, RC) 4 sigma (nx, RC)
  static constexpr bool odd(int a) { return ((a == 1 || a == 2) ? (PARX + 1) & 1 : PARX) != 0; }
  static constexpr int rows(int a) { return (a == 0 || a == 2) ? RN : RC; }
  static constexpr int hbox(int a) { return odd(a) ? (rows(a) + 1) / 2 : rows(a); }   // box height
  static constexpr int sub(int a) { return pad16(W * hbox(a)); }                        // one box in smem (doubles)
  static constexpr int size(int a) { return odd(a) ? 2 * sub(a) : sub(a); }
  // (no recursion: these are also called with unrolled-loop indices in device code)
  static constexpr int off(int a) {
    return (a > 0 ? size(0) : 0) + (a > 1 ? size(1) : 0) + (a > 2 ? size(2) : 0) + (a > 3 ? size(3) : 0);
  }
  static constexpr int STAGE = size(0) + size(1) + size(2) + size(3) + size(4);
  static constexpr unsigned bytes(int a) { return (unsigned)(W * hbox(a) * 8 * (odd(a) ? 2 : 1)); }
  static constexpr int XROWS = R;                       // mailbox rows: 0 = halo row j0-1, rr = row j0+rr-1
  static constexpr int XCH = NSLOT * XROWS * XS * 64;   // doubles: [slot][rr][strip][G_FX | G_FZ][lane]
  static constexpr int NBAR = 2 * NSLOT + NSLOT * XROWS * XS;
  static constexpr size_t SMEM = (size_t)(NSLOT * STAGE + XCH) * 8 + (size_t)NBAR * 8 + 2 * (M3_MAXCHUNK + 2) * 8 + 128;
};

struct M3Maps {
  CUtensorMap m[5];   // whole array (even pitch) or its even rows (odd pitch)
  CUtensorMap o[5];   // odd pitch: super-rows, used for the odd rows
  // z-slabs with peer-read halos: the same two maps of ex, ey, ez in the memory of the lower (0) / upper (1) neighbour
  CUtensorMap pm[2][3];
  CUtensorMap po[2][3];
};

// Which local planes of a z-slab are halo planes and where they live in the neighbour's local edge vector.  The ring
// fetches them straight from the neighbour's memory over NVLink (tensor maps on the peer allocation), so a distributed
// apply is ONE launch per rank: no send / recv, no separate launch for the planes next to a halo.
struct M3Halo {
  int lo_plane, lo_peer_plane;   // ex, ey on local node plane lo_plane = plane lo_peer_plane of the lower neighbour (-1: none)
  int lo_layer, lo_peer_layer;   // ez in local cell layer lo_layer = layer lo_peer_layer of the lower neighbour (-1: none)
  int hi_plane, hi_peer_plane;   // ex, ey on local node plane hi_plane = plane hi_peer_plane of the upper neighbour (-1: none)
  // ordering against the neighbours' writes inside the kernel (NULL: the caller put a barrier in front of the launch):
  // sync[0] / sync[1] are written by the lower / upper neighbour, sync[2] = launches completed, sync[3] = blocks of this
  // launch that have started, sync[4] != 0: a wait gave up.  Same protocol as K8 (tm_dc_fused.cu, DcHalo): a launch
  // announces "e of epoch sync[2] + 1 is complete" to both neighbours from its first block, and the producer lane waits for
  // the neighbour's announcement of that epoch before its first copy out of that neighbour's memory.
  unsigned *sync, *lo_signal, *hi_signal;
};

template <int XS, int R, int PARX, int NSLOT, int MINB, bool PEER>
__global__ void __launch_bounds__((XS *R + XS + Mx3<XS, R, PARX, NSLOT>::NPW) * 32, MINB)
    k_maxwell_tmap(TM m, const __grid_constant__ M3Maps maps, double *__restrict__ y, int kchunk, int k_begin, int k_end,
                   int nzb0, int k_begin2, int k_end2, int l2hint, M3Halo hal) {
  typedef Mx3<XS, R, PARX, NSLOT> G;
  constexpr int S = NSLOT, W = G::W;
  static_assert(S % 2 == 0, "the row-parity bookkeeping assumes an even ring depth");
  // boxes must start on 128-byte boundaries.  (No pointer arithmetic through integers here: that would turn every tile
  // access into a generic LD / ST instead of LDS / STS.)
  extern __shared__ __align__(1024) unsigned char m3_smem_raw[];
  double *tiles = reinterpret_cast<double *>(m3_smem_raw);          // [S][STAGE]
  if ((smem_u32(tiles) & 127u) != 0u) __trap();
  double *xch = tiles + S * G::STAGE;                               // mailboxes
  uint64_t *full = reinterpret_cast<uint64_t *>(xch + G::XCH);      // [S]
  uint64_t *empty = full + S;                                       // [S]
  uint64_t *xready = empty + S;                                     // [S][XROWS][XS]
  double *s_hz = reinterpret_cast<double *>(xready + S * G::XROWS * XS);   // layers kb-1 .. ke-1 (0 outside)
  double *s_rzh = s_hz + (M3_MAXCHUNK + 2);                         // 0.5 / hz

  const int nx = (int)m.nx, ny = (int)m.ny, nz = (int)m.nz;
  const int nNx = nx + 1, nNy = ny + 1, nNz = nz + 1;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler
  const int j0 = blockIdx.y * R;
  const int Xb = blockIdx.x * (XS * M3_XOUT);
  // blocks z < nzb0 serve the plane range [k_begin, k_end), the others a second range [k_begin2, k_end2) (the two
  // boundary planes of a z-slab in one launch)
  const bool second = (int)blockIdx.z >= nzb0;
  const int kb = second ? k_begin2 + ((int)blockIdx.z - nzb0) * kchunk : k_begin + (int)blockIdx.z * kchunk;
  const int ke = min(kb + kchunk, second ? k_end2 : k_end);   // output planes kb .. ke-1 (layers kb .. min(ke, nz)-1)
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
      mbar_init(&full[q], G::NPW);
      mbar_init(&empty[q], G::NCONS + G::NHALO);
    }
    for (int q = 0; q < S * G::XROWS * XS; ++q) mbar_init(&xready[q], 1);
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

  // rows per plane of the five staged arrays; row J = plane * rpp + (row inside the plane)
  const int rpp[5] = {nNy, ny, nNy, ny, ny};
  // where plane / layer pl of array a is read from: 0 = this rank's array, 1 / 2 = the lower / upper neighbour's (then
  // `pl` becomes the plane index inside the neighbour's local array)
  auto source = [&](int a, int &pl) -> int {
    if (!PEER || a > 2) return 0;
    if (a < 2) {
      if (pl == hal.lo_plane) { pl = hal.lo_peer_plane; return 1; }
      if (pl == hal.hi_plane) { pl = hal.hi_peer_plane; return 2; }
      return 0;
    }
    if (pl == hal.lo_layer) { pl = hal.lo_peer_layer; return 1; }
    return 0;
  };

  if (warp >= G::NCONS + G::NHALO) {
    // ================= producer warp(s): one elected lane each, <= 7 box copies per stage =================
    const int pw = warp - (G::NCONS + G::NHALO);
    if (lane == 0) {
#pragma unroll
      for (int a = 0; a < 5; ++a) {
        if (!G::mine(a, pw)) continue;
        tmap_prefetch(&maps.m[a]);
        if (G::odd(a)) tmap_prefetch(&maps.o[a]);
      }
      if (PEER) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          if (hal.lo_plane >= 0) tmap_prefetch(&maps.pm[0][a]);
          if (hal.hi_plane >= 0) tmap_prefetch(&maps.pm[1][a]);
        }
      }
      const int npl[5] = {nNz, nNz, nz, nz, nz};
      const int gp[5] = {nx, nNx, nNx, nx, nx};
      const uint64_t pol = l2hint == 1 ? l2_policy_evict_last() : l2_policy_evict_first();
      auto load = [&](void *dst, const CUtensorMap *mp, int x, int yy, uint64_t *bar) {
        if (l2hint) tmap_load_2d_hint(dst, mp, x, yy, bar, pol);
        else tmap_load_2d(dst, mp, x, yy, bar);
      };
      unsigned announced = 0;   // neighbours whose epoch announcement this lane has seen
      for (int q = 0; q < nq; ++q) {
        const int slot = q % S, n = q / S;
        if (n >= 1) mbar_wait(&empty[slot], (n - 1) & 1);
        bool on[5];
        unsigned total = 0;
#pragma unroll
        for (int a = 0; a < 5; ++a) {
          const int pl = jfirst + q + (a < 2 ? 1 : 0);
          on[a] = G::mine(a, pw) && pl >= 0 && pl < npl[a] && !(q == 0 && a >= 2);   // ez / mui / sigma of the first stage are never read
          total += on[a] ? G::bytes(a) : 0u;
        }
        mbar_arrive_expect_tx(&full[slot], total);
        double *T = tiles + slot * G::STAGE;
        // the stage's copies; FROM_PEER: one of its planes is a halo plane read from a neighbour (at most two stages of a
        // block) -- everything else takes the path of the plain kernel, with compile-time map addresses (the producer
        // lane is on the critical path: with the neighbour tests in every stage the whole kernel ran 12-20 % slower)
        auto issue = [&](auto peer_tag) {
          constexpr bool FROM_PEER = decltype(peer_tag)::value;
#pragma unroll
          for (int a = 0; a < 5; ++a) {
            if (on[a]) {
              int pl = jfirst + q + (a < 2 ? 1 : 0);
              const CUtensorMap *mE = &maps.m[a], *mO = &maps.o[a];
              if (FROM_PEER) {
                const int src = source(a, pl);         // pl is now the plane index inside the array it is read from
                if (src) {
                  mE = &maps.pm[src - 1][a < 3 ? a : 0];
                  mO = &maps.po[src - 1][a < 3 ? a : 0];
                  if (hal.sync && !((announced >> src) & 1u)) {   // the neighbour's vector of this epoch must be complete
                    wait_signal_sys(hal.sync + (src - 1), epoch, hal.sync + 4);
                    fence_proxy_async_global();
                    announced |= 1u << src;
                  }
                }
              }
              const int J0 = pl * rpp[a] + j0 - 1;     // first staged row (may be -1: zero-filled / previous plane)
              if (G::odd(a)) {
                // sub-tile 0 holds the staged rows rr = 0, 2, 4, .., sub-tile 1 the rows rr = 1, 3, ..: global row J0 + rr
                // is even for the rr of parity t = J0 & 1, so the even-row box goes to sub-tile t, the odd-row box to 1 - t
                const int t = J0 & 1;
                load(T + G::off(a) + t * G::sub(a), mE, (Xb - 1) & ~1, (J0 + 1) >> 1, &full[slot]);
                load(T + G::off(a) + (t ^ 1) * G::sub(a), mO, (gp[a] + Xb - 1) & ~1, J0 >> 1, &full[slot]);
              } else {
                load(T + G::off(a), mE, (Xb - 1) & ~1, J0, &full[slot]);
              }
            }
          }
        };
        const int plane = jfirst + q + 1, layer = jfirst + q;
        if (PEER && (plane == hal.lo_plane || plane == hal.hi_plane || layer == hal.lo_layer)) issue(std::true_type{});
        else issue(std::false_type{});
      }
    }
    return;
  }

  // ================= consumer / halo warps =================
  const bool halo = warp >= G::NCONS;
  const int s = halo ? warp - G::NCONS : warp % XS;   // strip
  const int r = halo ? -1 : warp / XS;                // row inside the block (-1: the halo row below it)
  const int i = Xb + s * M3_XOUT + lane - 1;          // node column and cell column of this lane
  const int j = j0 + r;                               // node row and cell row
  const bool ci = (i >= 0) && (i < nx), cil = (i >= 1) && (i <= nx);
  const bool cj = (j >= 0) && (j < ny), cjd = (j >= 1) && (j <= ny);
  const double hx = ci ? __ldg(m.hx + i) : 0.0, rx = ci ? __ldg(m.rhx + i) : 0.0;
  const double hxl = cil ? __ldg(m.hx + i - 1) : 0.0;
  const double hy = cj ? __ldg(m.hy + j) : 0.0, ry = cj ? __ldg(m.rhy + j) : 0.0;
  const double hyd = cjd ? __ldg(m.hy + j - 1) : 0.0;
  const double cfz = 0.5 * rx * ry;

  // shared-memory addressing: thread part (column) + warp-uniform part (slot, array, row; for odd-pitch arrays the
  // sub-tile of a row depends on the parity of its global row number, which is fixed per ring slot because S is even)
  const double *tcol = tiles + (s * M3_XOUT + lane);
  // element offset of (column Xb-1, staged row rr) of array a inside the stage that holds its plane / layer pl (row 0 =
  // row j0-1).  Odd-pitch arrays keep their even / odd staged rows in two sub-tiles (the producer sorts the boxes
  // accordingly).  Boxes start on even columns: column Xb-1 sits at box column leadE = (Xb-1)&1 in boxes of even global
  // rows and -- the pitch being odd -- at 1-leadE in boxes of odd global rows.
  const int leadE = (Xb - 1) & 1;
  // `flip`: 1 where the plane was read from a neighbour's array and sits on a plane of the other parity there (the row
  // parity is that of the array the plane was read from); the plane index itself only enters through its parity, which
  // is the same in every ring cycle (S is even), so `pl` is jfirst + slot and everything but `flip` folds per slot
  auto rowoff = [&](int a, int rr, int pl, int flip) -> int {
    if (G::odd(a)) {
      const int Jodd = ((pl * rpp[a] + j0 - 1 + rr) & 1) ^ (flip & rpp[a] & 1);
      return G::off(a) + (rr & 1) * G::sub(a) + (rr >> 1) * W + (leadE ^ Jodd);
    }
    return G::off(a) + rr * W + leadE;
  };
  // parity flips of the node plane `pl` (ex, ey) / the cell layer `pl` (ez) when it comes from a neighbour
  // (kept to two compares per stage: the stages q whose plane / layer is a halo one, and their flips, are fixed per block)
  const int q_lo = PEER && hal.lo_plane >= 0 ? hal.lo_plane - 1 - jfirst : -1000000;
  const int q_hi = PEER && hal.hi_plane >= 0 ? hal.hi_plane - 1 - jfirst : -1000000;
  const int q_ll = PEER && hal.lo_layer >= 0 ? hal.lo_layer - jfirst : -1000000;
  const int f_lo = PEER ? (hal.lo_plane ^ hal.lo_peer_plane) & 1 : 0, f_hi = PEER ? (hal.hi_plane ^ hal.hi_peer_plane) & 1 : 0;
  const int f_ll = PEER ? (hal.lo_layer ^ hal.lo_peer_layer) & 1 : 0;
  auto flip_plane_q = [&](int q) -> int { return !PEER ? 0 : q == q_lo ? f_lo : q == q_hi ? f_hi : 0; };
  auto flip_layer_q = [&](int q) -> int { return !PEER ? 0 : q == q_ll ? f_ll : 0; };

  const long long nEx = (long long)nx * nNy * nNz, nEy = (long long)nNx * ny * nNz;
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
        const int fP = flip_plane_q(q), fL = flip_layer_q(q);
        mbar_wait(&full[u], par);
        const double *T = tcol + u * G::STAGE;
        const double ex1 = T[rowoff(0, 0, ku + 1, fP)], ex1u = T[rowoff(0, 1, ku + 1, fP)];
        const double *tEY = T + rowoff(1, 0, ku + 1, fP);
        const double ey1 = tEY[0], ey1r = tEY[1];
        if (q >= 1) {
          const double *tMU = T + rowoff(3, 0, ku, 0);
          const int t = q - 1;
          const double hz = s_hz[t], rzh = s_rzh[t];
          const double ez0 = T[rowoff(2, 0, ku, fL)], ezu = T[rowoff(2, 1, ku, fL)];
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
  double ex0 = 0.0, ex0u = 0.0, ey0 = 0.0, ey0r = 0.0;   // ex(i,j), ex(i,j+1), ey(i,j), ey(i+1,j) on plane k
  double gfxp = 0.0, gfyp = 0.0, mzp = 0.0, ap = 0.0, bp = 0.0;

  // output addressing: warp-uniform plane pointers (advanced on the uniform datapath) + a 32-bit in-plane offset per thread
  const int ic = min(max(i, 0), nx - 1), in_ = min(max(i, 0), nNx - 1), jc = min(j, ny - 1), jn = min(j, nNy - 1);
  const unsigned ox = (unsigned)(ic + nx * jn), oy = (unsigned)(in_ + nNx * jc), oz = (unsigned)(in_ + nNx * jn);
  double *yx = y + (long long)(kb - 1) * pEx;
  double *yy = y + (nEx + (long long)(kb - 1) * pEy);
  double *yz = y + (nEx + nEy + (long long)(kb - 1) * pEz);
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
      const int fP = flip_plane_q(q), fL = flip_layer_q(q);
      mbar_wait(&full[u], par);
      const double *T = tcol + u * G::STAGE;
      const double ex1 = T[rowoff(0, r + 1, ku + 1, fP)], ex1u = T[rowoff(0, r + 2, ku + 1, fP)];
      const double *tEY = T + rowoff(1, r + 1, ku + 1, fP);
      const double ey1 = tEY[0], ey1r = tEY[1];
      if (q >= 1) {
        const int k = jfirst + q;          // this step works on plane k / layer k
        const double *tEZ = T + rowoff(2, r + 1, ku, fL);
        const double *tMU = T + rowoff(3, r + 1, ku, 0);
        const int t = q - 1;
        const double hz = s_hz[t], rzh = s_rzh[t];
        const double ez0 = tEZ[0], ezu = T[rowoff(2, r + 2, ku, fL)];
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
        const double ezr = tEZ[1], mud = T[rowoff(3, r, ku, 0)];
        const double sg0 = T[rowoff(4, r + 1, ku, 0)], sgd = T[rowoff(4, r, ku, 0)];
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
          if (wx) yx[ox] = vx;
          if (wy) yy[oy] = vy;
          if (wz && k < nz) yz[oz] = vz;
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
