// K9c: y = C^T Mf(mui) C e + (a + ib) Me(sigma) e  for a COMPLEX edge field (interleaved re / im, numpy / torch
// complex128): the frequency-domain Maxwell operator SimPEG applies (alpha = i omega), in ONE launch.
//
// Same ring, same consumers, mailboxes and halo warps as tm_maxwell_tmap.cuh, with two differences:
//   * an edge value is a 16-byte (re, im) pair: the face weights (from mui) and the edge masses (from sigma) are computed
//     once and applied to both parts -- this is also the kernel's gain over two real applies: the two model arrays are
//     read once per complex apply, not once per part, and the weight arithmetic is shared;
//   * 16-byte elements make every row of ex / ey / ez start on a 16-byte boundary whatever nx is, so the three edge
//     components need no super-row maps and no lead bits: one plain 2-D tensor map each (the arrays described as float64
//     matrices with 2 * pitch columns).  The two real model arrays keep the scheme of tm_maxwell_tmap.cuh.
// The mass term takes the complex scalar:  Me e -> Me (a e_r - b e_i) + i Me (a e_i + b e_r).
#pragma once
#include "tm_maxwell_tmap.cuh"

namespace dzb {

struct cplx {
  double r, i;
};
__device__ __forceinline__ cplx operator+(cplx a, cplx b) { return {a.r + b.r, a.i + b.i}; }
__device__ __forceinline__ cplx operator-(cplx a, cplx b) { return {a.r - b.r, a.i - b.i}; }
__device__ __forceinline__ cplx operator*(double s, cplx a) { return {s * a.r, s * a.i}; }
__device__ __forceinline__ cplx shfl_up1(cplx a) {
  return {__shfl_up_sync(0xffffffffu, a.r, 1), __shfl_up_sync(0xffffffffu, a.i, 1)};
}
__device__ __forceinline__ cplx lds2(const double *p) {   // p is 16-byte aligned: one LDS.128
  const double2 v = *reinterpret_cast<const double2 *>(p);
  return {v.x, v.y};
}

// geometry: like Mx3, e arrays (a < 3) with 2 W doubles per row and never "odd"
template <int XS, int R, int PARX, int NSLOT> struct Mx3c {
  static constexpr int NCONS = XS * R, NHALO = XS, NWARPS = NCONS + NHALO + 1, NT = NWARPS * 32;
  static constexpr int RN = R + 2, RC = R + 1;
  static constexpr int WE = XS * M3_XOUT + 2;             // edge boxes: columns Xb-1 .. Xb+XS*31 (complex elements)
  static constexpr int W = (XS * M3_XOUT + 3 + 1) & ~1;   // model boxes: as in Mx3 (start on the even column <= Xb-1)
  static constexpr int pad16(int n) { return (n + 15) & ~15; }
  static constexpr bool odd(int a) { return a >= 3 && PARX != 0; }
  static constexpr int rows(int a) { return (a == 0 || a == 2) ? RN : RC; }
  static constexpr int hbox(int a) { return odd(a) ? (rows(a) + 1) / 2 : rows(a); }
  static constexpr int width(int a) { return a < 3 ? 2 * WE : W; }                       // doubles per staged row
  static constexpr int sub(int a) { return pad16(width(a) * hbox(a)); }
  static constexpr int size(int a) { return odd(a) ? 2 * sub(a) : sub(a); }
  static constexpr int off(int a) {
    return (a > 0 ? size(0) : 0) + (a > 1 ? size(1) : 0) + (a > 2 ? size(2) : 0) + (a > 3 ? size(3) : 0);
  }
  static constexpr int STAGE = size(0) + size(1) + size(2) + size(3) + size(4);
  static constexpr unsigned bytes(int a) { return (unsigned)(width(a) * hbox(a) * 8 * (odd(a) ? 2 : 1)); }
  static constexpr int XROWS = R;
  static constexpr int XCH = NSLOT * XROWS * XS * 128;    // doubles: [slot][rr][strip][G_FX | G_FZ][lane] of (re, im)
  static constexpr int NBAR = 2 * NSLOT + NSLOT * XROWS * XS;
  static constexpr size_t SMEM = (size_t)(NSLOT * STAGE + XCH) * 8 + (size_t)NBAR * 8 + 2 * (M3_MAXCHUNK + 2) * 8 + 128;
};

template <int XS, int R, int PARX, int NSLOT, int MINB>
__global__ void __launch_bounds__((XS *R + XS + 1) * 32, MINB)
    k_maxwell_cplx(TM m, const __grid_constant__ M3Maps maps, double *__restrict__ y, double ar, double ai, int kchunk,
                   int k_begin, int k_end) {
  typedef Mx3c<XS, R, PARX, NSLOT> G;
  constexpr int S = NSLOT, W = G::W, W2 = 2 * G::WE;
  static_assert(S % 2 == 0, "the row-parity bookkeeping assumes an even ring depth");
  extern __shared__ __align__(1024) unsigned char m3c_smem_raw[];
  double *tiles = reinterpret_cast<double *>(m3c_smem_raw);
  if ((smem_u32(tiles) & 127u) != 0u) __trap();
  double *xch = tiles + S * G::STAGE;
  uint64_t *full = reinterpret_cast<uint64_t *>(xch + G::XCH);
  uint64_t *empty = full + S;
  uint64_t *xready = empty + S;
  double *s_hz = reinterpret_cast<double *>(xready + S * G::XROWS * XS);
  double *s_rzh = s_hz + (M3_MAXCHUNK + 2);

  const int nx = (int)m.nx, ny = (int)m.ny, nz = (int)m.nz;
  const int nNx = nx + 1, nNy = ny + 1, nNz = nz + 1;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int j0 = blockIdx.y * R;
  const int Xb = blockIdx.x * (XS * M3_XOUT);
  const int kb = k_begin + (int)blockIdx.z * kchunk;
  const int ke = min(kb + kchunk, k_end);
  const int jfirst = kb - 2;
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
      mbar_init(&full[q], 1);
      mbar_init(&empty[q], G::NCONS + G::NHALO);
    }
    for (int q = 0; q < S * G::XROWS * XS; ++q) mbar_init(&xready[q], 1);
    fence_barrier_init();
  }
  fence_proxy_async_smem();
  __syncthreads();

  const int rpp[5] = {nNy, ny, nNy, ny, ny};

  if (warp == G::NWARPS - 1) {
    // ================= producer: one elected lane =================
    if (lane == 0) {
#pragma unroll
      for (int a = 0; a < 5; ++a) {
        tmap_prefetch(&maps.m[a]);
        if (G::odd(a)) tmap_prefetch(&maps.o[a]);
      }
      const int npl[5] = {nNz, nNz, nz, nz, nz};
      const uint64_t pol = l2_policy_evict_last();
      for (int q = 0; q < nq; ++q) {
        const int slot = q % S, n = q / S;
        if (n >= 1) mbar_wait(&empty[slot], (n - 1) & 1);
        bool on[5];
        unsigned total = 0;
#pragma unroll
        for (int a = 0; a < 5; ++a) {
          const int pl = jfirst + q + (a < 2 ? 1 : 0);
          on[a] = pl >= 0 && pl < npl[a] && !(q == 0 && a >= 2);
          total += on[a] ? G::bytes(a) : 0u;
        }
        mbar_arrive_expect_tx(&full[slot], total);
        double *T = tiles + slot * G::STAGE;
#pragma unroll
        for (int a = 0; a < 5; ++a) {
          if (on[a]) {
            const int pl = jfirst + q + (a < 2 ? 1 : 0);
            const int J0 = pl * rpp[a] + j0 - 1;
            if (a < 3) {
              tmap_load_2d_hint(T + G::off(a), &maps.m[a], 2 * (Xb - 1), J0, &full[slot], pol);   // complex columns: x in doubles
            } else if (G::odd(a)) {
              const int t = J0 & 1;
              tmap_load_2d_hint(T + G::off(a) + t * G::sub(a), &maps.m[a], (Xb - 1) & ~1, (J0 + 1) >> 1, &full[slot], pol);
              tmap_load_2d_hint(T + G::off(a) + (t ^ 1) * G::sub(a), &maps.o[a], (nx + Xb - 1) & ~1, J0 >> 1, &full[slot], pol);
            } else {
              tmap_load_2d_hint(T + G::off(a), &maps.m[a], (Xb - 1) & ~1, J0, &full[slot], pol);
            }
          }
        }
      }
    }
    return;
  }

  // ================= consumer / halo warps =================
  const bool halo = warp >= G::NCONS;
  const int s = halo ? warp - G::NCONS : warp % XS;
  const int r = halo ? -1 : warp / XS;
  const int i = Xb + s * M3_XOUT + lane - 1;
  const int j = j0 + r;
  const bool ci = (i >= 0) && (i < nx), cil = (i >= 1) && (i <= nx);
  const bool cj = (j >= 0) && (j < ny), cjd = (j >= 1) && (j <= ny);
  const double hx = ci ? __ldg(m.hx + i) : 0.0, rx = ci ? __ldg(m.rhx + i) : 0.0;
  const double hxl = cil ? __ldg(m.hx + i - 1) : 0.0;
  const double hy = cj ? __ldg(m.hy + j) : 0.0, ry = cj ? __ldg(m.rhy + j) : 0.0;
  const double hyd = cjd ? __ldg(m.hy + j - 1) : 0.0;
  const double cfz = 0.5 * rx * ry;

  const int col = s * M3_XOUT + lane;                       // staged column of this lane (column Xb-1 is 0)
  const double *tE = tiles + 2 * col;                       // edge tiles: (re, im) pairs, 16-byte aligned
  const double *tM = tiles + col;                           // model tiles
  const int leadE = (Xb - 1) & 1;
  auto erow = [&](int a, int rr) -> int { return G::off(a) + rr * W2; };
  auto mrow = [&](int a, int rr, int pl) -> int {
    if (G::odd(a)) {
      const int Jodd = (pl * rpp[a] + j0 - 1 + rr) & 1;
      return G::off(a) + (rr & 1) * G::sub(a) + (rr >> 1) * W + (leadE ^ Jodd);
    }
    return G::off(a) + rr * W + leadE;
  };

  if (halo) {
    cplx ex0 = {0.0, 0.0}, ex0u = {0.0, 0.0}, ey0 = {0.0, 0.0}, ey0r = {0.0, 0.0};
    double mzp = 0.0;
    double *xw = xch + s * 128 + 2 * lane;
    uint64_t *xw_bar = xready + s;
    for (int qb = 0; qb < nq; qb += S) {
      const unsigned par = (unsigned)(qb / S) & 1u;
#pragma unroll
      for (int u = 0; u < S; ++u) {
        const int q = qb + u;
        if (q >= nq) break;
        const int ku = jfirst + u;
        mbar_wait(&full[u], par);
        const double *TE = tE + u * G::STAGE, *TM = tM + u * G::STAGE;
        const cplx ex1 = lds2(TE + erow(0, 0)), ex1u = lds2(TE + erow(0, 1));
        const cplx ey1 = lds2(TE + erow(1, 0)), ey1r = lds2(TE + erow(1, 0) + 2);
        if (q >= 1) {
          const double *tMU = TM + mrow(3, 0, ku);
          const int t = q - 1;
          const double hz = s_hz[t], rzh = s_rzh[t];
          const cplx ez0 = lds2(TE + erow(2, 0)), ezu = lds2(TE + erow(2, 1));
          const double mu0 = tMU[0], mul = tMU[-1];
          const cplx circx = hz * (ezu - ez0) - hy * (ey1 - ey0);
          const cplx circz = hy * (ey0r - ey0) - hx * (ex0u - ex0);
          const double mz = hz * mu0;
          const cplx gfz = ((mzp + mz) * cfz) * circz;
          const cplx gfx = ((hxl * mul + hx * mu0) * (ry * rzh)) * circx;
          mzp = mz;
          double *X = xw + u * (G::XROWS * XS * 128);
          *reinterpret_cast<double2 *>(X) = make_double2(gfx.r, gfx.i);
          *reinterpret_cast<double2 *>(X + 64) = make_double2(gfz.r, gfz.i);
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

  const double vxy4 = 0.25 * hx * hy, vxyd4 = 0.25 * hx * hyd;
  cplx ex0 = {0.0, 0.0}, ex0u = {0.0, 0.0}, ey0 = {0.0, 0.0}, ey0r = {0.0, 0.0}, gfxp = {0.0, 0.0}, gfyp = {0.0, 0.0};
  double mzp = 0.0, ap = 0.0, bp = 0.0;

  const long long nEx = (long long)nx * nNy * nNz, nEy = (long long)nNx * ny * nNz;
  const long long pEx = (long long)nx * nNy, pEy = (long long)nNx * ny, pEz = (long long)nNx * nNy;
  const int ic = min(max(i, 0), nx - 1), in_ = min(max(i, 0), nNx - 1), jc = min(j, ny - 1), jn = min(j, nNy - 1);
  const unsigned ox = (unsigned)(ic + nx * jn), oy = (unsigned)(in_ + nNx * jc), oz = (unsigned)(in_ + nNx * jn);
  double2 *yc = reinterpret_cast<double2 *>(y);
  double2 *yx = yc + (long long)(kb - 1) * pEx;
  double2 *yy = yc + (nEx + (long long)(kb - 1) * pEy);
  double2 *yz = yc + (nEx + nEy + (long long)(kb - 1) * pEz);
  const bool lane_out = lane >= 1;
  const bool wx = lane_out && ci && (j < nNy);
  const bool wy = lane_out && (i >= 0) && (i < nNx) && cj;
  const bool wz = lane_out && (i >= 0) && (i < nNx) && (j < nNy);

  double *xw = xch + ((r + 1) * XS + s) * 128 + 2 * lane;
  const double *xr = xch + (r * XS + s) * 128 + 2 * lane;
  uint64_t *xw_bar = xready + (r + 1) * XS + s;
  uint64_t *xr_bar = xready + r * XS + s;
  const bool publish = r < R - 1;
  // (alpha Me) e = Me (a e_r - b e_i) + i Me (a e_i + b e_r)
  auto amul = [&](cplx e) -> cplx { return {ar * e.r - ai * e.i, ar * e.i + ai * e.r}; };

  for (int qb = 0; qb < nq; qb += S) {
    const unsigned par = (unsigned)(qb / S) & 1u;
#pragma unroll
    for (int u = 0; u < S; ++u) {
      const int q = qb + u;
      if (q >= nq) break;
      const int ku = jfirst + u;
      mbar_wait(&full[u], par);
      const double *TE = tE + u * G::STAGE, *TM = tM + u * G::STAGE;
      const cplx ex1 = lds2(TE + erow(0, r + 1)), ex1u = lds2(TE + erow(0, r + 2));
      const cplx ey1 = lds2(TE + erow(1, r + 1)), ey1r = lds2(TE + erow(1, r + 1) + 2);
      if (q >= 1) {
        const int k = jfirst + q;
        const double *tMU = TM + mrow(3, r + 1, ku);
        const int t = q - 1;
        const double hz = s_hz[t], rzh = s_rzh[t];
        const cplx ez0 = lds2(TE + erow(2, r + 1)), ezu = lds2(TE + erow(2, r + 2));
        const double mu0 = tMU[0], mul = tMU[-1];
        const cplx circx = hz * (ezu - ez0) - hy * (ey1 - ey0);
        const cplx circz = hy * (ey0r - ey0) - hx * (ex0u - ex0);
        const double mz = hz * mu0;
        const cplx gfz = ((mzp + mz) * cfz) * circz;
        const cplx gfx = ((hxl * mul + hx * mu0) * (ry * rzh)) * circx;
        mzp = mz;
        if (publish) {
          double *X = xw + u * (G::XROWS * XS * 128);
          *reinterpret_cast<double2 *>(X) = make_double2(gfx.r, gfx.i);
          *reinterpret_cast<double2 *>(X + 64) = make_double2(gfz.r, gfz.i);
          __syncwarp();
          if (lane == 0) mbar_arrive(xw_bar + u * (G::XROWS * XS));
        }
        const unsigned xpar = (u == 0) ? (par ^ 1u) : par;
        const bool xok = mbar_test_wait(xr_bar + u * (G::XROWS * XS), xpar);
        const cplx ezr = lds2(TE + erow(2, r + 1) + 2);
        const double mud = TM[mrow(3, r, ku)];
        const double sg0 = TM[mrow(4, r + 1, ku)], sgd = TM[mrow(4, r, ku)];
        const cplx circy = hx * (ex1 - ex0) - hz * (ezr - ez0);
        const cplx gfy = ((hyd * mud + hy * mu0) * (rx * rzh)) * circy;
        const cplx gfzl = shfl_up1(gfz);
        const cplx gfyl = shfl_up1(gfy);
        const double sp = sg0 * vxy4, sdp = sgd * vxyd4;
        const double a = hz * (sp + sdp);
        const double slp = __shfl_up_sync(0xffffffffu, sp, 1);
        const double b = hz * (sp + slp);
        const double al = __shfl_up_sync(0xffffffffu, a, 1);
        if (!xok) mbar_wait(xr_bar + u * (G::XROWS * XS), xpar);
        const double *X = xr + u * (G::XROWS * XS * 128);
        const cplx gfxd = lds2(X), gfzd = lds2(X + 64);
        const cplx vx = hx * ((gfyp - gfy) + (gfz - gfzd)) + (ap + a) * amul(ex0);
        const cplx vy = hy * ((gfx - gfxp) + (gfzl - gfz)) + (bp + b) * amul(ey0);
        const cplx vz = hz * ((gfxd - gfx) + (gfy - gfyl)) + (a + al) * amul(ez0);
        gfyp = gfy;
        gfxp = gfx;
        ap = a;
        bp = b;
        if (k >= kb) {
          if (wx) __stcs(yx + ox, make_double2(vx.r, vx.i));
          if (wy) __stcs(yy + oy, make_double2(vy.r, vy.i));
          if (wz && k < nz) __stcs(yz + oz, make_double2(vz.r, vz.i));
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
