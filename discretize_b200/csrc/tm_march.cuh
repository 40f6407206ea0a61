// Plane-marching kernels for the six core DiffOperators applies: face_divergence, nodal_gradient, edge_curl
// and their transposes (reference: operators/differential_operators.py:161-235, 593-664, 2091-2181).
//
// Why: the generic row-generator kernels (tm_rowkernels.cuh) spend 50-60 instructions per output row on 64-bit
// index arithmetic and bounds predicates; ncu (profiles/r1b_*) shows GRAD / CURL / the transposes ISSUE-bound
// (0.65-1.1 ms of pure issue time at 512^3) although their DRAM traffic equals the algorithmic bytes.  Here a
// thread owns one (i, j) column, walks KC planes along z with pointers that advance by constant strides, keeps
// the z-neighbour it will need again in a register, and reads x / y neighbours directly (L1 hits).  All
// bounds tests on i and j are hoisted out of the loop: 6-13 instructions per output row.
//
// Every kernel computes differences first, then scales by 1/h:  (x1 - x0) * rh  -- one rounding less than the
// reference's  (-rh) x0 + rh x1  and well inside the 1e-12 parity tolerance.
#pragma once
#include "tm_common.cuh"
#include "tm_rows.cuh"

namespace dzb {

#include <type_traits>

// launch configuration: BY rows of 32 lanes per block, KC planes per block, MINB = minimum resident blocks per SM
// (caps the registers per thread; 0 = let the compiler choose)
template <int CBY, int CKC, int CMINB> struct MarchCfg {
  static constexpr int BY = CBY, KC = CKC, MINB = CMINB, THREADS = 32 * CBY;
};
#define DZB_MARCH_BOUNDS(CFG) __launch_bounds__(CFG::THREADS, CFG::MINB)

struct MarchGeom {
  i64 i, j, k0, k1;
  template <class CFG> __device__ __forceinline__ bool init(i64 ext_x, i64 ext_y, i64 ext_z) {
    i = (i64)blockIdx.x * 32 + threadIdx.x;
    j = (i64)blockIdx.y * CFG::BY + threadIdx.y;
    k0 = (i64)blockIdx.z * CFG::KC;
    k1 = min(k0 + CFG::KC, ext_z);
    return i < ext_x && j < ext_y && k0 < k1;
  }
};

__device__ __forceinline__ double ldp(const double *__restrict__ p, bool ok) { return ok ? __ldg(p) : 0.0; }

// The z loop of every kernel below: planes k0 <= k < min(k1, nz) have a cell layer at k (HZ = true, no
// k-dependent predicate in the body, so the unrolled loads of several planes are issued together); the top
// node plane k == nz, if it is in the chunk, is one peeled step with HZ = false.
#define DZB_MARCH_LOOP(STEP)                                  \
  {                                                           \
    i64 k = g.k0;                                             \
    const i64 kmain = min(g.k1, nz);                          \
    _Pragma("unroll 4") for (; k < kmain; ++k) STEP(std::true_type{}, k);  \
    if (k < g.k1) STEP(std::false_type{}, k);                 \
  }

// ---- nodal_gradient: nodes -> edges --------------------------------------------------------------------------
// Ex(i,j,k) = (phi(i+1,j,k) - phi(i,j,k)) / hx_i   etc.; thread (i,j) over the node extent, k over node planes.
template <class CFG>
__global__ void DZB_MARCH_BOUNDS(CFG) k_grad_march(TM m, const double *__restrict__ phi, double *__restrict__ e) {
  MarchGeom g;
  const i64 nx = m.nx, ny = m.ny, nz = m.nz;
  if (!g.template init<CFG>(nx + 1, ny + 1, nz + 1)) return;
  const i64 i = g.i, j = g.j;
  const bool hx = i < nx, hy = j < ny;
  const double rx = hx ? __ldg(m.rhx + i) : 0.0, ry = hy ? __ldg(m.rhy + j) : 0.0;
  const i64 pN = nx + 1, plN = pN * (ny + 1), plEx = nx * (ny + 1), plEy = pN * ny;
  const double *p = phi + i + pN * j + plN * g.k0;
  double *ex = e + i + nx * j + plEx * g.k0;
  double *ey = e + m.off<LOC_EY>() + i + pN * j + plEy * g.k0;
  double *ez = e + m.off<LOC_EZ>() + i + pN * j + plN * g.k0;
  double a = __ldg(p);
  auto step = [&](auto hz_tag, i64 k) {
    constexpr bool HZ = decltype(hz_tag)::value;
    const double ax = ldp(p + 1, hx), ay = ldp(p + pN, hy);
    if (hx) *ex = (ax - a) * rx;
    if (hy) *ey = (ay - a) * ry;
    if (HZ) {
      const double az = __ldg(p + plN);
      *ez = (az - a) * __ldg(m.rhz + k);
      a = az;
    }
    p += plN; ex += plEx; ey += plEy; ez += plN;
  };
  DZB_MARCH_LOOP(step)
}

// ---- nodal_gradient^T: edges -> nodes -------------------------------------------------------------------------
template <class CFG>
__global__ void DZB_MARCH_BOUNDS(CFG) k_gradT_march(TM m, const double *__restrict__ e, double *__restrict__ out) {
  MarchGeom g;
  const i64 nx = m.nx, ny = m.ny, nz = m.nz;
  if (!g.template init<CFG>(nx + 1, ny + 1, nz + 1)) return;
  const i64 i = g.i, j = g.j;
  const bool hx = i < nx, hy = j < ny, lx = i >= 1, ly = j >= 1;
  const double rx = hx ? __ldg(m.rhx + i) : 0.0, ry = hy ? __ldg(m.rhy + j) : 0.0;
  const double rxm = lx ? __ldg(m.rhx + i - 1) : 0.0, rym = ly ? __ldg(m.rhy + j - 1) : 0.0;
  const i64 pN = nx + 1, plN = pN * (ny + 1), plEx = nx * (ny + 1), plEy = pN * ny;
  const double *ex = e + i + nx * j + plEx * g.k0;
  const double *ey = e + m.off<LOC_EY>() + i + pN * j + plEy * g.k0;
  const double *ez = e + m.off<LOC_EZ>() + i + pN * j + plN * g.k0;
  double *o = out + i + pN * j + plN * g.k0;
  // z-flux through the layer below the first plane of this chunk
  double zlo = g.k0 >= 1 ? __ldg(ez - plN) * __ldg(m.rhz + g.k0 - 1) : 0.0;
  auto step = [&](auto hz_tag, i64 k) {
    constexpr bool HZ = decltype(hz_tag)::value;
    double zhi = 0.0;
    if (HZ) zhi = __ldg(ez) * __ldg(m.rhz + k);
    double acc = zlo - zhi;
    acc += ldp(ex - 1, lx) * rxm - ldp(ex, hx) * rx;
    acc += ldp(ey - pN, ly) * rym - ldp(ey, hy) * ry;
    *o = acc;
    zlo = zhi;
    ex += plEx; ey += plEy; ez += plN; o += plN;
  };
  DZB_MARCH_LOOP(step)
}

// ---- face_divergence: faces -> cells ---------------------------------------------------------------------------
template <class CFG>
__global__ void DZB_MARCH_BOUNDS(CFG) k_div_march(TM m, const double *__restrict__ u, double *__restrict__ out) {
  MarchGeom g;
  const i64 nx = m.nx, ny = m.ny, nz = m.nz;
  if (!g.template init<CFG>(nx, ny, nz)) return;
  const i64 i = g.i, j = g.j;
  const double rx = __ldg(m.rhx + i), ry = __ldg(m.rhy + j);
  const i64 pFx = nx + 1, plFx = pFx * ny, plFy = nx * (ny + 1), plC = nx * ny;
  const double *fx = u + i + pFx * j + plFx * g.k0;
  const double *fy = u + m.off<LOC_FY>() + i + nx * j + plFy * g.k0;
  const double *fz = u + m.off<LOC_FZ>() + i + nx * j + plC * g.k0;
  double *o = out + i + nx * j + plC * g.k0;
  double z0 = __ldg(fz);
#pragma unroll 4
  for (i64 k = g.k0; k < g.k1; ++k) {
    const double z1 = __ldg(fz + plC);
    const double dx = __ldg(fx + 1) - __ldg(fx), dy = __ldg(fy + nx) - __ldg(fy);
    *o = dx * rx + dy * ry + (z1 - z0) * __ldg(m.rhz + k);
    z0 = z1;
    fx += plFx; fy += plFy; fz += plC; o += plC;
  }
}

// ---- face_divergence^T: cells -> faces --------------------------------------------------------------------------
// thread (i,j) over the node extent; k over node planes (Fz lives there, Fx / Fy on the cell layers k < nz).
template <class CFG>
__global__ void DZB_MARCH_BOUNDS(CFG) k_divT_march(TM m, const double *__restrict__ c, double *__restrict__ u) {
  MarchGeom g;
  const i64 nx = m.nx, ny = m.ny, nz = m.nz;
  if (!g.template init<CFG>(nx + 1, ny + 1, nz + 1)) return;
  const i64 i = g.i, j = g.j;
  const bool hx = i < nx, hy = j < ny, lx = i >= 1, ly = j >= 1;
  const double rx = hx ? __ldg(m.rhx + i) : 0.0, ry = hy ? __ldg(m.rhy + j) : 0.0;
  const double rxm = lx ? __ldg(m.rhx + i - 1) : 0.0, rym = ly ? __ldg(m.rhy + j - 1) : 0.0;
  const i64 pFx = nx + 1, plFx = pFx * ny, plFy = nx * (ny + 1), plC = nx * ny;
  const double *pc = c + i + nx * j + plC * g.k0;          // cell (i,j,k); only dereferenced where it exists
  double *fx = u + i + pFx * j + plFx * g.k0;
  double *fy = u + m.off<LOC_FY>() + i + nx * j + plFy * g.k0;
  double *fz = u + m.off<LOC_FZ>() + i + nx * j + plC * g.k0;
  const bool in_xy = hx && hy, lxy = lx && hy, lyx = ly && hx;
  double zlo = (g.k0 >= 1 && in_xy) ? __ldg(pc - plC) * __ldg(m.rhz + g.k0 - 1) : 0.0;
  auto step = [&](auto hz_tag, i64 k) {
    constexpr bool HZ = decltype(hz_tag)::value;
    double zhi = 0.0;
    if (HZ) {
      const double c0 = ldp(pc, in_xy), cxm = ldp(pc - 1, lxy), cym = ldp(pc - nx, lyx);
      if (hy) *fx = cxm * rxm - c0 * rx;      // x-face (i,j,k), j < ny   (c0 = 0 where i == nx)
      if (hx) *fy = cym * rym - c0 * ry;      // y-face (i,j,k), i < nx   (c0 = 0 where j == ny)
      zhi = c0 * __ldg(m.rhz + k);
    }
    if (in_xy) *fz = zlo - zhi;               // z-face (i,j,k), k <= nz
    zlo = zhi;
    pc += plC; fx += plFx; fy += plFy; fz += plC;
  };
  DZB_MARCH_LOOP(step)
}

// ---- edge_curl: edges -> faces ------------------------------------------------------------------------------------
// Fx = d_y Ez - d_z Ey,  Fy = d_z Ex - d_x Ez,  Fz = d_x Ey - d_y Ex.  Thread (i,j) over the node extent, k over
// node planes; the plane-k values of Ex(i,j), Ex(i,j+1), Ey(i,j), Ey(i+1,j) are carried from the previous step.
template <class CFG>
__global__ void DZB_MARCH_BOUNDS(CFG) k_curl_march(TM m, const double *__restrict__ e, double *__restrict__ u) {
  MarchGeom g;
  const i64 nx = m.nx, ny = m.ny, nz = m.nz;
  if (!g.template init<CFG>(nx + 1, ny + 1, nz + 1)) return;
  const i64 i = g.i, j = g.j;
  const bool hx = i < nx, hy = j < ny, hxy = hx && hy;
  const double rx = hx ? __ldg(m.rhx + i) : 0.0, ry = hy ? __ldg(m.rhy + j) : 0.0;
  const i64 pN = nx + 1, plN = pN * (ny + 1), plEx = nx * (ny + 1), plEy = pN * ny;
  const i64 plFx = pN * ny, plFy = nx * (ny + 1), plC = nx * ny;
  const double *ex = e + i + nx * j + plEx * g.k0;                           // Ex(i,j,k):   i < nx
  const double *ey = e + m.off<LOC_EY>() + i + pN * j + plEy * g.k0;         // Ey(i,j,k):   j < ny
  const double *ez = e + m.off<LOC_EZ>() + i + pN * j + plN * g.k0;          // Ez(i,j,k):   k < nz
  double *fx = u + i + pN * j + plFx * g.k0;                                 // Fx(i,j,k):   j < ny, k < nz
  double *fy = u + m.off<LOC_FY>() + i + nx * j + plFy * g.k0;               // Fy(i,j,k):   i < nx, k < nz
  double *fz = u + m.off<LOC_FZ>() + i + nx * j + plC * g.k0;                // Fz(i,j,k):   i < nx, j < ny
  double ex00 = ldp(ex, hx), ex01 = ldp(ex + nx, hxy), ey00 = ldp(ey, hy), ey10 = ldp(ey + 1, hxy);
  auto step = [&](auto hz_tag, i64 k) {
    constexpr bool HZ = decltype(hz_tag)::value;
    if (hxy) *fz = (ey10 - ey00) * rx - (ex01 - ex00) * ry;
    if (HZ) {
      // next plane of Ex / Ey (exists because k + 1 <= nz), this layer of Ez
      const double nex00 = ldp(ex + plEx, hx), nex01 = ldp(ex + plEx + nx, hxy);
      const double ney00 = ldp(ey + plEy, hy), ney10 = ldp(ey + plEy + 1, hxy);
      const double rz = __ldg(m.rhz + k);
      const double ez00 = __ldg(ez), ez01 = ldp(ez + pN, hy), ez10 = ldp(ez + 1, hx);
      if (hy) *fx = (ez01 - ez00) * ry - (ney00 - ey00) * rz;
      if (hx) *fy = (nex00 - ex00) * rz - (ez10 - ez00) * rx;
      ex00 = nex00; ex01 = nex01; ey00 = ney00; ey10 = ney10;
    }
    ex += plEx; ey += plEy; ez += plN; fx += plFx; fy += plFy; fz += plC;
  };
  DZB_MARCH_LOOP(step)
}

// ---- edge_curl^T: faces -> edges -------------------------------------------------------------------------------------
// with DT_a(F)(p) = F(p - 1) rh_a[p - 1] - F(p) rh_a[p] (terms outside the mesh dropped):
//   Ex = DT_z Fy - DT_y Fz,   Ey = DT_x Fz - DT_z Fx,   Ez = DT_y Fx - DT_x Fy.
template <class CFG>
__global__ void DZB_MARCH_BOUNDS(CFG) k_curlT_march(TM m, const double *__restrict__ u, double *__restrict__ e) {
  MarchGeom g;
  const i64 nx = m.nx, ny = m.ny, nz = m.nz;
  if (!g.template init<CFG>(nx + 1, ny + 1, nz + 1)) return;
  const i64 i = g.i, j = g.j;
  const bool hx = i < nx, hy = j < ny, lx = i >= 1, ly = j >= 1;
  const bool hxy = hx && hy, hx_ly = hx && ly, lx_hy = lx && hy;
  const double rx = hx ? __ldg(m.rhx + i) : 0.0, ry = hy ? __ldg(m.rhy + j) : 0.0;
  const double rxm = lx ? __ldg(m.rhx + i - 1) : 0.0, rym = ly ? __ldg(m.rhy + j - 1) : 0.0;
  const i64 pN = nx + 1, plN = pN * (ny + 1), plEx = nx * (ny + 1), plEy = pN * ny;
  const i64 plFx = pN * ny, plFy = nx * (ny + 1), plC = nx * ny;
  const double *fx = u + i + pN * j + plFx * g.k0;                           // Fx(i,j,k): j < ny, k < nz
  const double *fy = u + m.off<LOC_FY>() + i + nx * j + plFy * g.k0;         // Fy(i,j,k): i < nx, k < nz
  const double *fz = u + m.off<LOC_FZ>() + i + nx * j + plC * g.k0;          // Fz(i,j,k): i < nx, j < ny
  double *ex = e + i + nx * j + plEx * g.k0;
  double *ey = e + m.off<LOC_EY>() + i + pN * j + plEy * g.k0;
  double *ez = e + m.off<LOC_EZ>() + i + pN * j + plN * g.k0;
  // z-terms of the layer below this chunk: Fy(i,j,k0-1) rhz, Fx(i,j,k0-1) rhz
  double zy_lo = 0.0, zx_lo = 0.0;
  if (g.k0 >= 1) {
    const double rzm = __ldg(m.rhz + g.k0 - 1);
    zy_lo = ldp(fy - plFy, hx) * rzm;
    zx_lo = ldp(fx - plFx, hy) * rzm;
  }
  auto step = [&](auto hz_tag, i64 k) {
    constexpr bool HZ = decltype(hz_tag)::value;
    double zy_hi = 0.0, zx_hi = 0.0;
    // DT_y Fz and DT_x Fz on this node plane
    const double fz00 = ldp(fz, hxy), fz0m = ldp(fz - nx, hx_ly), fzm0 = ldp(fz - 1, lx_hy);
    if (HZ) {
      const double rz = __ldg(m.rhz + k);
      const double fx00 = ldp(fx, hy), fy00 = ldp(fy, hx), fx0m = ldp(fx - pN, ly), fym0 = ldp(fy - 1, lx);
      zy_hi = fy00 * rz;
      zx_hi = fx00 * rz;
      *ez = (fx0m * rym - fx00 * ry) - (fym0 * rxm - fy00 * rx);
    }
    if (hx) *ex = (zy_lo - zy_hi) - (fz0m * rym - fz00 * ry);
    if (hy) *ey = (fzm0 * rxm - fz00 * rx) - (zx_lo - zx_hi);
    zy_lo = zy_hi; zx_lo = zx_hi;
    fx += plFx; fy += plFy; fz += plC; ex += plEx; ey += plEy; ez += plN;
  };
  DZB_MARCH_LOOP(step)
}

// ---- the cell -> face family: face_divergence^T, (stencil_)cell_gradient with boundary conditions,
// average_cell_to_face (operators/differential_operators.py:41-82, 1413-1595, 2789-2828) -------------------------
// Every one of them is  face(p) = a_lo(p) * cell(p - 1) + a_hi(p) * cell(p)  along the face normal, with weights that
// depend on the position p along that axis only.  The x / y weights are per-thread constants, the z weights are
// computed once per plane; the walk is the one of k_divT_march.
enum C2FKind { C2F_DIVT = 0, C2F_CELL_GRAD = 1, C2F_AVE = 2 };

// weights at node position p of an axis with n cells; flags: GcFlags of that axis (C2F_CELL_GRAD only)
template <int KIND>
__device__ __forceinline__ void c2f_weights(i64 p, i64 n, const double *__restrict__ h, const double *__restrict__ rh,
                                            int flags, double &a_lo, double &a_hi) {
  const bool lo = p >= 1, hi = p < n;
  if (KIND == C2F_DIVT) {
    a_lo = lo ? __ldg(rh + p - 1) : 0.0;
    a_hi = hi ? -__ldg(rh + p) : 0.0;
  } else if (KIND == C2F_AVE) {
    a_lo = lo ? (hi ? 0.5 : 1.0) : 0.0;
    a_hi = hi ? (lo ? 0.5 : 1.0) : 0.0;
  } else {
    bool keep;
    a_lo = lo ? gc_value(p, p - 1, n, h, rh, flags, &keep) : 0.0;
    a_hi = hi ? gc_value(p, p, n, h, rh, flags, &keep) : 0.0;
  }
}

template <int KIND, class CFG>
__global__ void DZB_MARCH_BOUNDS(CFG) k_c2f_march(TM m, int fx, int fy, int fz, const double *__restrict__ c, double *__restrict__ u) {
  // the z weights depend on k only: the block computes them once (the cell-gradient ones cost an FP64 division each)
  __shared__ double s_lo[CFG::KC], s_hi[CFG::KC];
  MarchGeom g;
  const i64 nx = m.nx, ny = m.ny, nz = m.nz;
  const bool live = g.template init<CFG>(nx + 1, ny + 1, nz + 1);
  for (int t = threadIdx.y * 32 + threadIdx.x; t < CFG::KC; t += CFG::THREADS) {
    const i64 k = (i64)blockIdx.z * CFG::KC + t;
    double lo = 0.0, hi = 0.0;
    if (k <= nz) c2f_weights<KIND>(k, nz, m.hz, m.rhz, fz, lo, hi);
    s_lo[t] = lo;
    s_hi[t] = hi;
  }
  __syncthreads();
  if (!live) return;
  const i64 i = g.i, j = g.j;
  const bool hx = i < nx, hy = j < ny, lx = i >= 1, ly = j >= 1;
  double ax_lo, ax_hi, ay_lo, ay_hi;
  c2f_weights<KIND>(i, nx, m.hx, m.rhx, fx, ax_lo, ax_hi);
  c2f_weights<KIND>(j, ny, m.hy, m.rhy, fy, ay_lo, ay_hi);
  const i64 pFx = nx + 1, plFx = pFx * ny, plFy = nx * (ny + 1), plC = nx * ny;
  const double *pc = c + i + nx * j + plC * g.k0;          // cell (i,j,k); only dereferenced where it exists
  double *px = u + i + pFx * j + plFx * g.k0;
  double *py = u + m.off<LOC_FY>() + i + nx * j + plFy * g.k0;
  double *pz = u + m.off<LOC_FZ>() + i + nx * j + plC * g.k0;
  const bool in_xy = hx && hy, lxy = lx && hy, lyx = ly && hx;
  double cm = (g.k0 >= 1 && in_xy) ? __ldg(pc - plC) : 0.0;   // cell (i,j,k-1)
  auto step = [&](auto hz_tag, i64 k) {
    constexpr bool HZ = decltype(hz_tag)::value;
    const double az_lo = s_lo[k - g.k0], az_hi = s_hi[k - g.k0];
    double c0 = 0.0;
    if (HZ) {
      c0 = ldp(pc, in_xy);
      const double cxm = ldp(pc - 1, lxy), cym = ldp(pc - nx, lyx);
      if (hy) *px = ax_lo * cxm + ax_hi * c0;     // x-face (i,j,k), j < ny   (c0 = 0 where i == nx)
      if (hx) *py = ay_lo * cym + ay_hi * c0;     // y-face (i,j,k), i < nx   (c0 = 0 where j == ny)
    }
    if (in_xy) *pz = az_lo * cm + az_hi * c0;     // z-face (i,j,k), k <= nz
    cm = c0;
    pc += plC; px += plFx; py += plFy; pz += plC;
  };
  DZB_MARCH_LOOP(step)
}

// ---- averages to cell centres ------------------------------------------------------------------------------------
// average_face_to_cell: (1/3) (1/2) (the two faces per direction) -- the same walk as the divergence
template <class CFG>
__global__ void DZB_MARCH_BOUNDS(CFG) k_aveF2CC_march(TM m, const double *__restrict__ u, double *__restrict__ out) {
  MarchGeom g;
  const i64 nx = m.nx, ny = m.ny, nz = m.nz;
  if (!g.template init<CFG>(nx, ny, nz)) return;
  const i64 i = g.i, j = g.j;
  const i64 pFx = nx + 1, plFx = pFx * ny, plFy = nx * (ny + 1), plC = nx * ny;
  const double *fx = u + i + pFx * j + plFx * g.k0;
  const double *fy = u + m.off<LOC_FY>() + i + nx * j + plFy * g.k0;
  const double *fz = u + m.off<LOC_FZ>() + i + nx * j + plC * g.k0;
  double *o = out + i + nx * j + plC * g.k0;
  const double w = (1.0 / 3.0) * 0.5;
  double z0 = __ldg(fz);
#pragma unroll 4
  for (i64 k = g.k0; k < g.k1; ++k) {
    const double z1 = __ldg(fz + plC);
    const double sx = __ldg(fx) + __ldg(fx + 1), sy = __ldg(fy) + __ldg(fy + nx);
    *o = w * sx + w * sy + w * (z0 + z1);
    z0 = z1;
    fx += plFx; fy += plFy; fz += plC; o += plC;
  }
}

// average_edge_to_cell: (1/3) (1/4) (the four edges per direction); the plane sums of Ex / Ey are carried
template <class CFG>
__global__ void DZB_MARCH_BOUNDS(CFG) k_aveE2CC_march(TM m, const double *__restrict__ e, double *__restrict__ out) {
  MarchGeom g;
  const i64 nx = m.nx, ny = m.ny, nz = m.nz;
  if (!g.template init<CFG>(nx, ny, nz)) return;
  const i64 i = g.i, j = g.j;
  const i64 pN = nx + 1, plN = pN * (ny + 1), plEx = nx * (ny + 1), plEy = pN * ny, plC = nx * ny;
  const double *ex = e + i + nx * j + plEx * g.k0;
  const double *ey = e + m.off<LOC_EY>() + i + pN * j + plEy * g.k0;
  const double *ez = e + m.off<LOC_EZ>() + i + pN * j + plN * g.k0;
  double *o = out + i + nx * j + plC * g.k0;
  const double w = (1.0 / 3.0) * 0.25;
  double sx0 = __ldg(ex) + __ldg(ex + nx), sy0 = __ldg(ey) + __ldg(ey + 1);
#pragma unroll 4
  for (i64 k = g.k0; k < g.k1; ++k) {
    const double sx1 = __ldg(ex + plEx) + __ldg(ex + plEx + nx), sy1 = __ldg(ey + plEy) + __ldg(ey + plEy + 1);
    const double sz = (__ldg(ez) + __ldg(ez + 1)) + (__ldg(ez + pN) + __ldg(ez + pN + 1));
    *o = w * (sx0 + sx1) + w * (sy0 + sy1) + w * sz;
    sx0 = sx1; sy0 = sy1;
    ex += plEx; ey += plEy; ez += plN; o += plC;
  }
}

// average_node_to_cell: 1/8 of the eight corner nodes; the four-node sum of a plane is carried
template <class CFG>
__global__ void DZB_MARCH_BOUNDS(CFG) k_aveN2CC_march(TM m, const double *__restrict__ v, double *__restrict__ out) {
  MarchGeom g;
  const i64 nx = m.nx, ny = m.ny, nz = m.nz;
  if (!g.template init<CFG>(nx, ny, nz)) return;
  const i64 i = g.i, j = g.j;
  const i64 pN = nx + 1, plN = pN * (ny + 1), plC = nx * ny;
  const double *p = v + i + pN * j + plN * g.k0;
  double *o = out + i + nx * j + plC * g.k0;
  double s0 = (__ldg(p) + __ldg(p + 1)) + (__ldg(p + pN) + __ldg(p + pN + 1));
#pragma unroll 4
  for (i64 k = g.k0; k < g.k1; ++k) {
    p += plN;
    const double s1 = (__ldg(p) + __ldg(p + 1)) + (__ldg(p + pN) + __ldg(p + pN + 1));
    *o = 0.125 * (s0 + s1);
    s0 = s1;
    o += plC;
  }
}

template <class CFG> static dim3 march_grid(i64 ext_x, i64 ext_y, i64 ext_z) {
  return dim3((unsigned)((ext_x + 31) / 32), (unsigned)((ext_y + CFG::BY - 1) / CFG::BY),
              (unsigned)((ext_z + CFG::KC - 1) / CFG::KC));
}

// returns true when (op, transpose) has a marching kernel and it was launched
template <class CFG>
static bool launch_march(int op, int transpose, const TM &m, const double *x, double *y, cudaStream_t s) {
  const dim3 b(32, CFG::BY, 1);
  const dim3 gn = march_grid<CFG>(m.nx + 1, m.ny + 1, m.nz + 1);
  if (op == DZB_OP_GRAD && !transpose) k_grad_march<CFG><<<gn, b, 0, s>>>(m, x, y);
  else if (op == DZB_OP_GRAD) k_gradT_march<CFG><<<gn, b, 0, s>>>(m, x, y);
  else if (op == DZB_OP_DIV && !transpose) k_div_march<CFG><<<march_grid<CFG>(m.nx, m.ny, m.nz), b, 0, s>>>(m, x, y);
  else if (op == DZB_OP_DIV) k_divT_march<CFG><<<gn, b, 0, s>>>(m, x, y);
  else if (op == DZB_OP_CURL && !transpose) k_curl_march<CFG><<<gn, b, 0, s>>>(m, x, y);
  else if (op == DZB_OP_CURL) k_curlT_march<CFG><<<gn, b, 0, s>>>(m, x, y);
  else if (op == DZB_OP_AVE_F2CC && !transpose) k_aveF2CC_march<CFG><<<march_grid<CFG>(m.nx, m.ny, m.nz), b, 0, s>>>(m, x, y);
  else if (op == DZB_OP_AVE_E2CC && !transpose) k_aveE2CC_march<CFG><<<march_grid<CFG>(m.nx, m.ny, m.nz), b, 0, s>>>(m, x, y);
  else if (op == DZB_OP_AVE_N2CC && !transpose) k_aveN2CC_march<CFG><<<march_grid<CFG>(m.nx, m.ny, m.nz), b, 0, s>>>(m, x, y);
  else return false;
  return true;
}

// Per-operator launch configuration from the 512^3 sweep (profiles/r1b_sweep_march.log): the read-heavy kernels
// (D, G^T: 3 arrays in, 1 out) want 64 registers so that the loads of four planes are in flight together; the
// write-heavy ones (G, D^T) do best at the compiler's own 40 registers and more resident warps.
static bool launch_march_tuned(int op, int transpose, const TM &m, const double *x, double *y, cudaStream_t s) {
  typedef MarchCfg<8, 16, 4> Wide;    // 64 registers
  typedef MarchCfg<8, 16, 0> Mid;
  typedef MarchCfg<8, 32, 0> Mid32;
  typedef MarchCfg<4, 32, 0> Narrow;
  const dim3 gnW = march_grid<Wide>(m.nx + 1, m.ny + 1, m.nz + 1), bW(32, Wide::BY, 1);
  if (op == DZB_OP_DIV && !transpose) k_div_march<Wide><<<march_grid<Wide>(m.nx, m.ny, m.nz), bW, 0, s>>>(m, x, y);
  else if (op == DZB_OP_DIV) k_divT_march<Narrow><<<march_grid<Narrow>(m.nx + 1, m.ny + 1, m.nz + 1), dim3(32, Narrow::BY, 1), 0, s>>>(m, x, y);
  else if (op == DZB_OP_GRAD && !transpose) k_grad_march<Mid><<<march_grid<Mid>(m.nx + 1, m.ny + 1, m.nz + 1), dim3(32, Mid::BY, 1), 0, s>>>(m, x, y);
  else if (op == DZB_OP_GRAD) k_gradT_march<Wide><<<gnW, bW, 0, s>>>(m, x, y);
  else if (op == DZB_OP_CURL && !transpose) k_curl_march<Mid32><<<march_grid<Mid32>(m.nx + 1, m.ny + 1, m.nz + 1), dim3(32, Mid32::BY, 1), 0, s>>>(m, x, y);
  else if (op == DZB_OP_CURL) k_curlT_march<Wide><<<gnW, bW, 0, s>>>(m, x, y);
  else if (op == DZB_OP_AVE_F2CC && !transpose) k_aveF2CC_march<Wide><<<march_grid<Wide>(m.nx, m.ny, m.nz), bW, 0, s>>>(m, x, y);
  else if (op == DZB_OP_AVE_E2CC && !transpose) k_aveE2CC_march<Wide><<<march_grid<Wide>(m.nx, m.ny, m.nz), bW, 0, s>>>(m, x, y);
  else if (op == DZB_OP_AVE_N2CC && !transpose) k_aveN2CC_march<Wide><<<march_grid<Wide>(m.nx, m.ny, m.nz), bW, 0, s>>>(m, x, y);
  else return false;
  return true;
}

// cell -> face family with run-time flags (cell-gradient boundary conditions): returns true when launched
static bool launch_c2f_march(int op, int transpose, int flags, const TM &m, const double *x, double *y, cudaStream_t s) {
  typedef MarchCfg<4, 32, 0> CFG;
  if (transpose) return false;
  const dim3 gn = march_grid<CFG>(m.nx + 1, m.ny + 1, m.nz + 1), b(32, CFG::BY, 1);
  if (op == DZB_OP_AVE_CC2F) {
    k_c2f_march<C2F_AVE, CFG><<<gn, b, 0, s>>>(m, 0, 0, 0, x, y);
  } else if (op == DZB_OP_CELL_GRAD) {
    const int sc = (flags & DZB_CG_SCALED) ? GC_SCALED : 0;
    k_c2f_march<C2F_CELL_GRAD, CFG><<<gn, b, 0, s>>>(m, (flags & 3) | sc, ((flags >> 2) & 3) | sc, ((flags >> 4) & 3) | sc, x, y);
  } else {
    return false;
  }
  return true;
}

}  // namespace dzb
