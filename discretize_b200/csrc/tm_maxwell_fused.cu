// K9: y = C^T Mf(mui) C e + Me(sigma) e  (Maxwell, E-formulation) without temporaries.
//
// Reference composition (tests/boundaries/test_boundary_maxwell.py:95-99):
//   C  = edge_curl                      (operators/differential_operators.py:2169-2181)
//   Mf = get_face_inner_product(mui)    (operators/inner_products.py:35-67, isotropic)
//   Me = get_edge_inner_product(sigma)  (operators/inner_products.py:69-99, isotropic)
// A 13-point edge stencil: every edge couples to the edges of its (<=4) adjacent faces.
//
// Variant "gather": one thread per output edge evaluates the whole chain for its row
// from global memory (L1/L2 serve the re-reads).  It is the simple, obviously-correct
// form and the cross-check for the marching variant.
#include "tm_mass_common.cuh"
#include "tm_maxwell_rows.cuh"
#include "tm_maxwell_tmap.cuh"
#include "tm_maxwell_cplx.cuh"
#include <cudaTypedefs.h>
#include <cstdio>
#include <cstdlib>
#include "tm_rowkernels.cuh"

namespace dzb {

// Mf(f) * (C e)(f) for the face of component FC at (a,b,c)
template <int FC>
__device__ __forceinline__ double face_flux(const TM &m, const Model &mu, const double *__restrict__ e, i64 a, i64 b,
                                            i64 c) {
  double ce = 0.0;
  Rows<DZB_OP_CURL, false>::row<FC>(m, a, b, c, [&](i64 col, double v) { ce += v * __ldg(e + col); });
  return mass_diag_entry<false, FC>(m, mu, a, b, c) * ce;
}

template <int COMP>
__global__ void __launch_bounds__(BX *BY *BZ)
    k_maxwell_gather(TM m, Model mu, Model sg, const double *__restrict__ e, double *__restrict__ y) {
  constexpr int LOC = DofLoc<true, COMP>::LOC;
  const i64 i = (i64)blockIdx.x * BX + threadIdx.x;
  const i64 j = (i64)blockIdx.y * BY + threadIdx.y;
  const i64 k = (i64)blockIdx.z * BZ + threadIdx.z;
  if (i >= m.sx<LOC>() || j >= m.sy<LOC>() || k >= m.sz<LOC>()) return;
  const i64 row = m.idx<LOC>(i, j, k);
  double acc = mass_diag_entry<true, COMP>(m, sg, i, j, k) * __ldg(e + row);
  // rows of C^T (cf. Rows<DZB_OP_CURL, true>)
  if (COMP == 0) {
    term<AX_ID, AX_ID, AX_DT>(m, i, j, k, 1.0, [&](i64 a, i64 b, i64 c, double w) { acc += w * face_flux<1>(m, mu, e, a, b, c); });
    term<AX_ID, AX_DT, AX_ID>(m, i, j, k, -1.0, [&](i64 a, i64 b, i64 c, double w) { acc += w * face_flux<2>(m, mu, e, a, b, c); });
  } else if (COMP == 1) {
    term<AX_ID, AX_ID, AX_DT>(m, i, j, k, -1.0, [&](i64 a, i64 b, i64 c, double w) { acc += w * face_flux<0>(m, mu, e, a, b, c); });
    term<AX_DT, AX_ID, AX_ID>(m, i, j, k, 1.0, [&](i64 a, i64 b, i64 c, double w) { acc += w * face_flux<2>(m, mu, e, a, b, c); });
  } else {
    term<AX_ID, AX_DT, AX_ID>(m, i, j, k, 1.0, [&](i64 a, i64 b, i64 c, double w) { acc += w * face_flux<0>(m, mu, e, a, b, c); });
    term<AX_DT, AX_ID, AX_ID>(m, i, j, k, -1.0, [&](i64 a, i64 b, i64 c, double w) { acc += w * face_flux<1>(m, mu, e, a, b, c); });
  }
  y[row] = acc;
}

}  // namespace dzb

using namespace dzb;

static int g_mx_variant = 3;   // 0 = gather (one thread per edge, cross-check), 2 = row-per-warp ring fed by row-wise bulk copies, 3 = same ring fed by tensor maps
static int g_mx_last_path = -1;   // which variant served the last call (tests assert that even-sized meshes take the tensor-map path)
extern "C" int dzb_tm_maxwell_last_path(void) { return g_mx_last_path; }
static int g_m2_kchunk = 0;    // 0 = automatic (~128 planes per block)
static int g_m2_dbg = 0;   // experiments only: 1 = consumers skip the arithmetic, 2 = producers skip the copies
extern "C" void dzb_tune_maxwell_debug(int d) { g_m2_dbg = d; }
static int g_m2_xs = 2, g_m2_r = 4, g_m2_ns = 4, g_m2_np = 2;
extern "C" void dzb_tune_maxwell(int variant, int kchunk) {
  g_mx_variant = variant;
  g_m2_kchunk = kchunk > 0 ? (kchunk > M2_MAXCHUNK ? M2_MAXCHUNK : kchunk) : 0;
}
// block shape of variant 2: strips x rows, one of (1,8) (2,4) (4,2) (4,4); producer warps 1 or 2
extern "C" void dzb_tune_maxwell_rows(int xs, int r, int nslot, int nprod) {
  g_m2_xs = xs;
  g_m2_r = r;
  g_m2_ns = nslot;
  g_m2_np = nprod;
}

template <int XS, int R, int PARX, int NSLOT, int NPROD, int MINB>
static int launch_m2_t(const TM &m, const double *mui, const double *sigma, const double *e, double *y, int k_begin,
                       int k_end, cudaStream_t s) {
  typedef Mx2<XS, R, PARX, NSLOT, NPROD> G;
  // the attribute is per device; setting it is cheap, so no process-wide "done" flag
  if (cudaFuncSetAttribute(k_maxwell_rows<XS, R, PARX, NSLOT, NPROD, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)G::SMEM) != cudaSuccess)
    return -1;
  const int nNx = (int)m.nx + 1, nNy = (int)m.ny + 1, planes = k_end - k_begin;
  const int xblocks = (nNx + XS * M2_XOUT - 1) / (XS * M2_XOUT), yblocks = (nNy + R - 1) / R;
  // ~128 planes per block amortise the two warm-up stages; short plane ranges (z-slabs on many GPUs, small meshes) are
  // cut finer until the grid has about four waves of blocks
  int nchunks = g_m2_kchunk > 0 ? (planes + g_m2_kchunk - 1) / g_m2_kchunk : (planes + 143) / 144;
  if (g_m2_kchunk == 0) {
    const long long nxy = (long long)xblocks * yblocks;
    const long long want = (4ll * 148 * MINB + nxy - 1) / nxy;
    if (nchunks < want) nchunks = (int)(want < (planes + 7) / 8 ? want : (planes + 7) / 8);
    if (nchunks < 1) nchunks = 1;
  }
  const int kchunk = (planes + nchunks - 1) / nchunks;
  dim3 grid((unsigned)xblocks, (unsigned)yblocks, (unsigned)((planes + kchunk - 1) / kchunk));
  k_maxwell_rows<XS, R, PARX, NSLOT, NPROD, MINB><<<grid, dim3(G::NT), G::SMEM, s>>>(m, mui, sigma, e, y, kchunk, k_begin, k_end, g_m2_dbg);
  return 0;
}

#define M2_LAUNCH(XS, R, NS, NP, MINB)                                                         \
  return parx ? launch_m2_t<XS, R, 1, NS, NP, MINB>(m, mui, sigma, e, y, k_begin, k_end, s)    \
              : launch_m2_t<XS, R, 0, NS, NP, MINB>(m, mui, sigma, e, y, k_begin, k_end, s)

static int launch_m2(const TM &m, const double *mui, const double *sigma, const double *e, double *y, int k_begin,
                     int k_end, cudaStream_t s) {
  const int parx = (int)(m.nx & 1);
  const int key = g_m2_xs * 1000 + g_m2_r * 100 + g_m2_ns * 10 + g_m2_np;
  switch (key) {
    case 2441: M2_LAUNCH(2, 4, 4, 1, 2);
    case 2442: M2_LAUNCH(2, 4, 4, 2, 2);
    case 2444: M2_LAUNCH(2, 4, 4, 4, 2);
    case 2461: M2_LAUNCH(2, 4, 6, 1, 2);
    case 2462: M2_LAUNCH(2, 4, 6, 2, 2);
    case 4241: M2_LAUNCH(4, 2, 4, 1, 2);
    case 4242: M2_LAUNCH(4, 2, 4, 2, 2);
    case 4244: M2_LAUNCH(4, 2, 4, 4, 2);
    case 4442: M2_LAUNCH(4, 4, 4, 2, 1);
    case 4444: M2_LAUNCH(4, 4, 4, 4, 1);
    case 4464: M2_LAUNCH(4, 4, 6, 4, 1);
    case 1842: M2_LAUNCH(1, 8, 4, 2, 2);
    default: break;
  }
  M2_LAUNCH(2, 4, 4, 2, 2);
}

// ---------------------------------------------------------------------------------------------------------------
// variant 3: the same ring fed by tensor-map TMA (tm_maxwell_tmap.cuh)
// ---------------------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled tmap_encoder() {
  static PFN_cuTensorMapEncodeTiled fn = []() -> PFN_cuTensorMapEncodeTiled {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess ||
        qr != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<PFN_cuTensorMapEncodeTiled>(p);
  }();
  return fn;
}

static int g_m3_xs = 2, g_m3_r = 8, g_m3_ns = 6;
static int g_m3_l2hint = 1;   // L2 cache hint of the box loads: 0 none, 1 evict_last, 2 evict_first
extern "C" void dzb_tune_maxwell_l2hint(int h) { g_m3_l2hint = h; }
static CUtensorMapL2promotion g_m3_l2promo = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
extern "C" void dzb_tune_maxwell_l2promo(int bytes) {
  g_m3_l2promo = bytes == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : bytes == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                 : bytes == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
}
extern "C" void dzb_tune_maxwell_tmap(int xs, int r, int nslot) {
  g_m3_xs = xs;
  g_m3_r = r;
  g_m3_ns = nslot;
}

// one (pointers, mesh, box shape) -> tensor maps entry per host thread: re-encoding costs a few microseconds per map
struct M3Key {
  const void *e, *mu, *sg;
  i64 nx, ny, nz;
  int w, r;
  bool operator==(const M3Key &o) const {
    return e == o.e && mu == o.mu && sg == o.sg && nx == o.nx && ny == o.ny && nz == o.nz && w == o.w && r == o.r;
  }
};

// The tensor maps of one array (pitch elements per row, `rows` rows): even pitch: the array as it is.  Odd pitch: m = its
// EVEN rows (width = one row, row stride = two rows), o = super-rows of two rows, of which the boxes take the second
// half: both stay inside the array whatever the parity of its row count, and a box never reads past the end of a row it
// was not asked for.
static int m3_encode_array(PFN_cuTensorMapEncodeTiled enc, const double *base, i64 pitch, i64 rows, int W, int hbox,
                           CUtensorMap *m, CUtensorMap *o) {
  if (reinterpret_cast<uintptr_t>(base) & 15) return -1;
  if (rows >= ((i64)1 << 31) || 2 * pitch >= ((i64)1 << 31)) return -1;
  const bool odd = (pitch & 1) != 0;
  cuuint32_t box[2] = {(cuuint32_t)W, (cuuint32_t)hbox};
  cuuint32_t estr[2] = {1, 1};
  cuuint64_t gdim[2] = {(cuuint64_t)pitch, (cuuint64_t)(odd ? (rows + 1) / 2 : rows)};
  cuuint64_t gstr[1] = {(cuuint64_t)(odd ? 16 * pitch : 8 * pitch)};
  if (enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double *>(base), gdim, gstr, box, estr,
          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, g_m3_l2promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) !=
      CUDA_SUCCESS)
    return -1;
  if (odd) {
    cuuint64_t odim[2] = {(cuuint64_t)(2 * pitch), (cuuint64_t)(rows / 2)};
    if (odim[1] == 0) odim[1] = 1, odim[0] = (cuuint64_t)pitch;   // a one-row array has no odd rows: everything the box asks for is out of bounds
    if (enc(o, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double *>(base), odim, gstr, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, g_m3_l2promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) !=
        CUDA_SUCCESS)
      return -1;
  } else {
    *o = *m;
  }
  return 0;
}

// returns 0 and fills `maps`, or -1 if this mesh / these pointers cannot be described (caller falls back to bulk copies)
template <typename G>
static int m3_encode(const TM &m, const double *mui, const double *sigma, const double *e, M3Maps &maps) {
  static thread_local M3Key last_key = {};
  static thread_local M3Maps last_maps;
  static thread_local bool have = false;
  const M3Key key{e, mui, sigma, m.nx, m.ny, m.nz, G::W, G::RN + 1000 * (int)g_m3_l2promo};
  if (have && key == last_key) {
    maps = last_maps;
    return 0;
  }
  PFN_cuTensorMapEncodeTiled enc = tmap_encoder();
  if (!enc) return -1;
  const i64 nx = m.nx, ny = m.ny, nz = m.nz, nNx = nx + 1, nNy = ny + 1, nNz = nz + 1;
  const i64 nEx = nx * nNy * nNz, nEy = nNx * ny * nNz;
  const double *base[5] = {e, e + nEx, e + nEx + nEy, mui, sigma};
  const i64 pitch[5] = {nx, nNx, nNx, nx, nx};
  const i64 rows[5] = {nNy * nNz, ny * nNz, nNy * nz, ny * nz, ny * nz};
  for (int a = 0; a < 5; ++a)
    if (m3_encode_array(enc, base[a], pitch[a], rows[a], G::W, G::hbox(a), &maps.m[a], &maps.o[a]) != 0) return -1;
  last_key = key;
  last_maps = maps;
  have = true;
  return 0;
}

// peer-read halos of a z-slab: the neighbours' local edge vectors and where the halo planes sit in them
struct M3Peers {
  const double *e_lo, *e_hi;   // NULL: no such neighbour
  i64 nz_lo, nz_hi;            // cell layers of the neighbours' local meshes (their edge vectors' layout)
  M3Halo halo;
};

template <typename G>
static int m3_encode_peers(const TM &m, const M3Peers &pr, M3Maps &maps) {
  PFN_cuTensorMapEncodeTiled enc = tmap_encoder();
  if (!enc) return -1;
  const i64 nx = m.nx, ny = m.ny, nNx = nx + 1, nNy = ny + 1;
  const double *pe[2] = {pr.e_lo, pr.e_hi};
  const i64 pnz[2] = {pr.nz_lo, pr.nz_hi};
  for (int side = 0; side < 2; ++side) {
    if (!pe[side]) {
      for (int a = 0; a < 3; ++a) maps.pm[side][a] = maps.m[a], maps.po[side][a] = maps.o[a];
      continue;
    }
    const i64 nz = pnz[side], nNz = nz + 1;
    const i64 nEx = nx * nNy * nNz, nEy = nNx * ny * nNz;
    const double *base[3] = {pe[side], pe[side] + nEx, pe[side] + nEx + nEy};
    const i64 pitch[3] = {nx, nNx, nNx};
    const i64 rows[3] = {nNy * nNz, ny * nNz, nNy * nz};
    for (int a = 0; a < 3; ++a)
      if (m3_encode_array(enc, base[a], pitch[a], rows[a], G::W, G::hbox(a), &maps.pm[side][a], &maps.po[side][a]) != 0)
        return -1;
  }
  return 0;
}

template <int XS, int R, int PARX, int NSLOT, int MINB>
static int launch_m3_t(const TM &m, const double *mui, const double *sigma, const double *e, double *y, int k_begin,
                       int k_end, int k_begin2, int k_end2, const M3Peers *peers, cudaStream_t s) {
  typedef Mx3<XS, R, PARX, NSLOT> G;
  M3Maps maps;
  if (m3_encode<G>(m, mui, sigma, e, maps) != 0) return -1;
  if (peers && m3_encode_peers<G>(m, *peers, maps) != 0) return -1;
  const void *kern = peers ? (const void *)k_maxwell_tmap<XS, R, PARX, NSLOT, MINB, true>
                           : (const void *)k_maxwell_tmap<XS, R, PARX, NSLOT, MINB, false>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM) != cudaSuccess) return -1;
  const int nNx = (int)m.nx + 1, nNy = (int)m.ny + 1, planes = k_end - k_begin, planes2 = k_end2 - k_begin2;
  const int longest = planes > planes2 ? planes : planes2;
  const int xblocks = (nNx + XS * M3_XOUT - 1) / (XS * M3_XOUT), yblocks = (nNy + R - 1) / R;
  int nchunks = g_m2_kchunk > 0 ? (longest + g_m2_kchunk - 1) / g_m2_kchunk : (longest + 143) / 144;
  if (g_m2_kchunk == 0) {
    const long long nxy = (long long)xblocks * yblocks;
    const long long want = (4ll * 148 * MINB + nxy - 1) / nxy;
    if (nchunks < want) nchunks = (int)(want < (longest + 7) / 8 ? want : (longest + 7) / 8);
    if (nchunks < 1) nchunks = 1;
  }
  const int kchunk = (longest + nchunks - 1) / nchunks;
  const int nzb0 = (planes + kchunk - 1) / kchunk, nzb1 = (planes2 + kchunk - 1) / kchunk;
  dim3 grid((unsigned)xblocks, (unsigned)yblocks, (unsigned)(nzb0 + nzb1));
  const M3Halo none = {-1, -1, -1, -1, -1, -1, nullptr, nullptr, nullptr};
  if (peers)
    k_maxwell_tmap<XS, R, PARX, NSLOT, MINB, true><<<grid, dim3(G::NT), G::SMEM, s>>>(m, maps, y, kchunk, k_begin, k_end, nzb0,
                                                                                      k_begin2, k_end2, g_m3_l2hint, peers->halo);
  else
    k_maxwell_tmap<XS, R, PARX, NSLOT, MINB, false><<<grid, dim3(G::NT), G::SMEM, s>>>(m, maps, y, kchunk, k_begin, k_end, nzb0,
                                                                                       k_begin2, k_end2, g_m3_l2hint, none);
  return 0;
}

#define M3_LAUNCH(XS, R, NS, MINB)                                                                                   \
  return parx ? launch_m3_t<XS, R, 1, NS, MINB>(m, mui, sigma, e, y, k_begin, k_end, k_begin2, k_end2, peers, s)     \
              : launch_m3_t<XS, R, 0, NS, MINB>(m, mui, sigma, e, y, k_begin, k_end, k_begin2, k_end2, peers, s)

static int launch_m3(const TM &m, const double *mui, const double *sigma, const double *e, double *y, int k_begin,
                     int k_end, int k_begin2, int k_end2, cudaStream_t s, const M3Peers *peers = nullptr) {
  static const bool env_read = []() {       // DZB_K9_SHAPE=xs,r,slots: experiment knob, like dzb_tune_maxwell_tmap
    int a, b, c;
    const char *v = getenv("DZB_K9_SHAPE");
    if (v && sscanf(v, "%d,%d,%d", &a, &b, &c) == 3) g_m3_xs = a, g_m3_r = b, g_m3_ns = c;
    return true;
  }();
  (void)env_read;
  const int parx = (int)(m.nx & 1);
  switch (g_m3_r < 10 ? g_m3_xs * 100 + g_m3_r * 10 + g_m3_ns : g_m3_xs * 1000 + g_m3_r * 10 + g_m3_ns) {
    case 144: M3_LAUNCH(1, 4, 4, 3);
    case 146: M3_LAUNCH(1, 4, 6, 3);
    case 154: M3_LAUNCH(1, 5, 4, 3);
    case 156: M3_LAUNCH(1, 5, 6, 3);
    case 164: M3_LAUNCH(1, 6, 4, 2);
    case 166: M3_LAUNCH(1, 6, 6, 2);
    case 168: M3_LAUNCH(1, 6, 8, 2);
    case 184: M3_LAUNCH(1, 8, 4, 2);
    case 186: M3_LAUNCH(1, 8, 6, 2);
    case 244: M3_LAUNCH(2, 4, 4, 2);
    case 246: M3_LAUNCH(2, 4, 6, 2);
    case 264: M3_LAUNCH(2, 6, 4, 1);
    case 266: M3_LAUNCH(2, 6, 6, 1);
    case 284: M3_LAUNCH(2, 8, 4, 1);
    case 286: M3_LAUNCH(2, 8, 6, 1);
    case 444: M3_LAUNCH(4, 4, 4, 1);
    case 1126: M3_LAUNCH(1, 12, 6, 1);
    case 1164: M3_LAUNCH(1, 16, 4, 1);
    case 1166: M3_LAUNCH(1, 16, 6, 1);
    case 2104: M3_LAUNCH(2, 10, 4, 1);
    default: break;
  }
  M3_LAUNCH(2, 8, 6, 1);
}

extern "C" int dzb_tm_maxwell_range(const DzbTensorMesh *mm, const double *mui, const double *sigma, const double *e,
                                    double *y, int64_t k_begin, int64_t k_end, void *stream);

// ---------------------------------------------------------------------------------------------------------------
// complex edge fields: y = C^T Mf(mui) C e + (a + ib) Me(sigma) e in one launch (tm_maxwell_cplx.cuh)
// ---------------------------------------------------------------------------------------------------------------
template <int XS, int R, int PARX, int NSLOT, int MINB>
static int launch_m3c_t(const TM &m, const double *mui, const double *sigma, double ar, double ai, const double *e, double *y,
                        cudaStream_t s) {
  typedef Mx3c<XS, R, PARX, NSLOT> G;
  PFN_cuTensorMapEncodeTiled enc = tmap_encoder();
  if (!enc) return -1;
  const i64 nx = m.nx, ny = m.ny, nz = m.nz, nNx = nx + 1, nNy = ny + 1, nNz = nz + 1;
  const i64 nEx = nx * nNy * nNz, nEy = nNx * ny * nNz;
  M3Maps maps;
  // the three edge components as float64 matrices with 2 * pitch columns (interleaved re / im): every row starts on a
  // 16-byte boundary, so one plain map each
  const double *eb[3] = {e, e + 2 * nEx, e + 2 * (nEx + nEy)};
  const i64 epitch[3] = {2 * nx, 2 * nNx, 2 * nNx};
  const i64 erows[3] = {nNy * nNz, ny * nNz, nNy * nz};
  for (int a = 0; a < 3; ++a)
    if (m3_encode_array(enc, eb[a], epitch[a], erows[a], 2 * G::WE, G::rows(a), &maps.m[a], &maps.o[a]) != 0) return -1;
  if (m3_encode_array(enc, mui, nx, ny * nz, G::W, G::hbox(3), &maps.m[3], &maps.o[3]) != 0) return -1;
  if (m3_encode_array(enc, sigma, nx, ny * nz, G::W, G::hbox(4), &maps.m[4], &maps.o[4]) != 0) return -1;
  for (int side = 0; side < 2; ++side)
    for (int a = 0; a < 3; ++a) maps.pm[side][a] = maps.m[a], maps.po[side][a] = maps.o[a];
  if (cudaFuncSetAttribute(k_maxwell_cplx<XS, R, PARX, NSLOT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)G::SMEM) != cudaSuccess)
    return -1;
  const int planes = (int)nNz;
  const int xblocks = (int)((nNx + XS * M3_XOUT - 1) / (XS * M3_XOUT)), yblocks = (int)((nNy + R - 1) / R);
  int nchunks = (planes + 143) / 144;
  const long long nxy = (long long)xblocks * yblocks;
  const long long want = (4ll * 148 * MINB + nxy - 1) / nxy;
  if (nchunks < want) nchunks = (int)(want < (planes + 7) / 8 ? want : (planes + 7) / 8);
  if (nchunks < 1) nchunks = 1;
  const int kchunk = (planes + nchunks - 1) / nchunks;
  dim3 grid((unsigned)xblocks, (unsigned)yblocks, (unsigned)((planes + kchunk - 1) / kchunk));
  k_maxwell_cplx<XS, R, PARX, NSLOT, MINB><<<grid, dim3(G::NT), G::SMEM, s>>>(m, maps, y, ar, ai, kchunk, 0, planes);
  return 0;
}

static int g_m3c_xs = 1, g_m3c_r = 6;
extern "C" void dzb_tune_maxwell_cplx(int xs, int r) {
  g_m3c_xs = xs;
  g_m3c_r = r;
}

// e, y: interleaved complex (2 * nE doubles).  Returns -2 where the tensor-map path does not apply (the caller then applies
// the real kernels to the two parts).
extern "C" int dzb_tm_maxwell_cplx(const DzbTensorMesh *mm, const double *mui, const double *sigma, double alpha_re,
                                   double alpha_im, const double *e, double *y, void *stream) {
  if (!valid_mesh(mm)) return -1;
  TM m(*mm);
  const bool ring_ok = m.ny >= 1 && (m.nx + 1) * (m.ny + 1) < ((i64)1 << 31);
  if (ring_ok) {
    const int parx = (int)(m.nx & 1);
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    if (g_m3c_xs == 2 && g_m3c_r == 4)
      rc = parx ? launch_m3c_t<2, 4, 1, 4, 1>(m, mui, sigma, alpha_re, alpha_im, e, y, s)
                : launch_m3c_t<2, 4, 0, 4, 1>(m, mui, sigma, alpha_re, alpha_im, e, y, s);
    else if (g_m3c_xs == 1 && g_m3c_r == 6)
      rc = parx ? launch_m3c_t<1, 6, 1, 4, 2>(m, mui, sigma, alpha_re, alpha_im, e, y, s)
                : launch_m3c_t<1, 6, 0, 4, 2>(m, mui, sigma, alpha_re, alpha_im, e, y, s);
    else
      rc = parx ? launch_m3c_t<1, 8, 1, 4, 1>(m, mui, sigma, alpha_re, alpha_im, e, y, s)
                : launch_m3c_t<1, 8, 0, 4, 1>(m, mui, sigma, alpha_re, alpha_im, e, y, s);
    if (rc == 0) return check_launch("dzb_tm_maxwell_cplx");
    cudaGetLastError();
  }
  set_error("dzb_tm_maxwell_cplx: needs the tensor-map path (16-byte aligned e, mui, sigma)");
  return -2;
}

// A z-slab's apply with PEER-READ halos: ex, ey on the local node planes lo_plane / hi_plane and ez in the local layer
// lo_plane are not read from `e` but from the neighbours' local edge vectors e_lo / e_hi (device pointers into peer memory
// mapped into this process: symmetric allocation / CUDA IPC), at plane lo_peer_plane / hi_peer_plane of those arrays, whose
// layout is that of a TensorMesh with nz_lo / nz_hi cell layers.  One launch covers all owned planes: no send / recv and no
// separate launch for the planes next to a halo.  The caller orders the apply after the neighbours' writes to their e
// (a barrier across the ranks).  Tensor-map path only: returns -2 where that path does not apply.
extern "C" int dzb_tm_maxwell_range_peer_sync(const DzbTensorMesh *mm, const double *mui, const double *sigma,
                                              const double *e, const double *e_lo, int64_t nz_lo, int64_t lo_plane,
                                              int64_t lo_peer_plane, const double *e_hi, int64_t nz_hi, int64_t hi_plane,
                                              int64_t hi_peer_plane, uint32_t *sync, uint32_t *lo_signal,
                                              uint32_t *hi_signal, double *y, int64_t k_begin, int64_t k_end,
                                              void *stream) {
  if (!valid_mesh(mm)) return -1;
  TM m(*mm);
  if (k_begin < 0 || k_end > m.nz + 1 || k_begin > k_end) {
    set_error("dzb_tm_maxwell_range_peer: plane range outside [0, nz+1]");
    return -1;
  }
  if (k_begin == k_end) return 0;
  const bool ring_ok = m.ny >= 1 && (m.nx + 1) * (m.ny + 1) < ((i64)1 << 31);
  M3Peers pr;
  pr.e_lo = e_lo;
  pr.e_hi = e_hi;
  pr.nz_lo = nz_lo;
  pr.nz_hi = nz_hi;
  pr.halo = {e_lo ? (int)lo_plane : -1, (int)lo_peer_plane, e_lo ? (int)lo_plane : -1, (int)lo_peer_plane,
             e_hi ? (int)hi_plane : -1, (int)hi_peer_plane, sync, sync ? lo_signal : nullptr, sync ? hi_signal : nullptr};
  if (g_mx_variant == 3 && ring_ok) {
    g_mx_last_path = 3;
    if (launch_m3(m, mui, sigma, e, y, (int)k_begin, (int)k_end, 0, 0, (cudaStream_t)stream, &pr) == 0)
      return check_launch("dzb_tm_maxwell_range_peer");
    cudaGetLastError();
  }
  set_error("dzb_tm_maxwell_range_peer: needs the tensor-map path (16-byte aligned component offsets: nx, ny even)");
  return -2;
}

extern "C" int dzb_tm_maxwell_range_peer(const DzbTensorMesh *mm, const double *mui, const double *sigma, const double *e,
                                         const double *e_lo, int64_t nz_lo, int64_t lo_plane, int64_t lo_peer_plane,
                                         const double *e_hi, int64_t nz_hi, int64_t hi_plane, int64_t hi_peer_plane,
                                         double *y, int64_t k_begin, int64_t k_end, void *stream) {
  return dzb_tm_maxwell_range_peer_sync(mm, mui, sigma, e, e_lo, nz_lo, lo_plane, lo_peer_plane, e_hi, nz_hi, hi_plane,
                                        hi_peer_plane, nullptr, nullptr, nullptr, y, k_begin, k_end, stream);
}

// two disjoint plane ranges in ONE launch (the two boundary planes of a z-slab, once its halos have arrived)
extern "C" int dzb_tm_maxwell_range2(const DzbTensorMesh *mm, const double *mui, const double *sigma, const double *e,
                                     double *y, int64_t k_begin, int64_t k_end, int64_t k_begin2, int64_t k_end2,
                                     void *stream) {
  if (!valid_mesh(mm)) return -1;
  TM m(*mm);
  if (k_begin < 0 || k_end > m.nz + 1 || k_begin > k_end || k_begin2 < 0 || k_end2 > m.nz + 1 || k_begin2 > k_end2) {
    set_error("dzb_tm_maxwell_range2: plane range outside [0, nz+1]");
    return -1;
  }
  if (k_begin == k_end) return dzb_tm_maxwell_range(mm, mui, sigma, e, y, k_begin2, k_end2, stream);
  if (k_begin2 == k_end2) return dzb_tm_maxwell_range(mm, mui, sigma, e, y, k_begin, k_end, stream);
  const bool ring_ok = m.ny >= 1 && (m.nx + 1) * (m.ny + 1) < ((i64)1 << 31);
  if (g_mx_variant == 3 && ring_ok) {
    g_mx_last_path = 3;
    if (launch_m3(m, mui, sigma, e, y, (int)k_begin, (int)k_end, (int)k_begin2, (int)k_end2, (cudaStream_t)stream) == 0)
      return check_launch("dzb_tm_maxwell");
    cudaGetLastError();
  }
  const int rc = dzb_tm_maxwell_range(mm, mui, sigma, e, y, k_begin, k_end, stream);
  return rc ? rc : dzb_tm_maxwell_range(mm, mui, sigma, e, y, k_begin2, k_end2, stream);
}

extern "C" int dzb_tm_maxwell_range(const DzbTensorMesh *mm, const double *mui, const double *sigma, const double *e,
                                    double *y, int64_t k_begin, int64_t k_end, void *stream) {
  if (!valid_mesh(mm)) return -1;
  TM m(*mm);
  if (k_begin < 0 || k_end > m.nz + 1 || k_begin > k_end) {
    set_error("dzb_tm_maxwell_range: plane range outside [0, nz+1]");
    return -1;
  }
  if (k_begin == k_end) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const bool ring_ok = m.ny >= 1 && (m.nx + 1) * (m.ny + 1) < ((i64)1 << 31);
  if (g_mx_variant == 3 && ring_ok) {
    g_mx_last_path = 3;
    if (launch_m3(m, mui, sigma, e, y, (int)k_begin, (int)k_end, 0, 0, s) == 0) return check_launch("dzb_tm_maxwell");
    cudaGetLastError();
  }
  if (g_mx_variant >= 2 && ring_ok) {
    g_mx_last_path = 2;
    if (launch_m2(m, mui, sigma, e, y, (int)k_begin, (int)k_end, s) == 0) return check_launch("dzb_tm_maxwell");
    cudaGetLastError();
  }
  g_mx_last_path = 0;
  if (k_begin != 0 || k_end != m.nz + 1) {
    set_error("dzb_tm_maxwell_range: the gather variant only supports the full plane range");
    return -1;
  }
  Model mu{DZB_MODEL_ISO, mui, 0.0, m.nC(), false};
  Model sg{DZB_MODEL_ISO, sigma, 0.0, m.nC(), false};
  dim3 b(BX, BY, BZ);
  k_maxwell_gather<0><<<grid_for<LOC_EX>(m), b, 0, s>>>(m, mu, sg, e, y);
  k_maxwell_gather<1><<<grid_for<LOC_EY>(m), b, 0, s>>>(m, mu, sg, e, y);
  k_maxwell_gather<2><<<grid_for<LOC_EZ>(m), b, 0, s>>>(m, mu, sg, e, y);
  return check_launch("dzb_tm_maxwell");
}

extern "C" int dzb_tm_maxwell(const DzbTensorMesh *mm, const double *mui, const double *sigma, const double *e,
                              double *y, void *stream) {
  if (!valid_mesh(mm)) return -1;
  return dzb_tm_maxwell_range(mm, mui, sigma, e, y, 0, mm->nz + 1, stream);
}
