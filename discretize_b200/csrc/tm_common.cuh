// Shared device helpers for the TensorMesh kernels (sm_100a).
// Index conventions follow the reference (base/base_regular_mesh.py:601-762):
// x fastest; faces [Fx;Fy;Fz], edges [Ex;Ey;Ez].
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/dzb.h"

namespace dzb {

typedef long long i64;

// Locations of degrees of freedom.  The low three bits say along which axes the
// location sits on node planes (extent n+1) rather than cell centres (extent n).
enum Loc : int {
  LOC_C = 0,
  LOC_FX = 1, LOC_FY = 2, LOC_FZ = 4,
  LOC_EX = 6, LOC_EY = 5, LOC_EZ = 3,
  LOC_N = 7
};

struct TM {
  i64 nx, ny, nz;
  const double *hx, *hy, *hz, *rhx, *rhy, *rhz;

  __host__ __device__ TM() {}
  __host__ TM(const DzbTensorMesh &m)
      : nx(m.nx), ny(m.ny), nz(m.nz), hx(m.hx), hy(m.hy), hz(m.hz), rhx(m.rhx), rhy(m.rhy), rhz(m.rhz) {}

  template <int LOC> __host__ __device__ i64 sx() const { return nx + ((LOC >> 0) & 1); }
  template <int LOC> __host__ __device__ i64 sy() const { return ny + ((LOC >> 1) & 1); }
  template <int LOC> __host__ __device__ i64 sz() const { return nz + ((LOC >> 2) & 1); }
  template <int LOC> __host__ __device__ i64 count() const { return sx<LOC>() * sy<LOC>() * sz<LOC>(); }
  __host__ __device__ i64 nC() const { return nx * ny * nz; }
  __host__ __device__ i64 nN() const { return count<LOC_N>(); }
  __host__ __device__ i64 nF() const { return count<LOC_FX>() + count<LOC_FY>() + count<LOC_FZ>(); }
  __host__ __device__ i64 nE() const { return count<LOC_EX>() + count<LOC_EY>() + count<LOC_EZ>(); }

  // offset of a component block inside the stacked face / edge vector
  template <int LOC> __host__ __device__ i64 off() const {
    if (LOC == LOC_FY) return count<LOC_FX>();
    if (LOC == LOC_FZ) return count<LOC_FX>() + count<LOC_FY>();
    if (LOC == LOC_EY) return count<LOC_EX>();
    if (LOC == LOC_EZ) return count<LOC_EX>() + count<LOC_EY>();
    return 0;
  }
  // flat index inside the component block
  template <int LOC> __host__ __device__ i64 lidx(i64 i, i64 j, i64 k) const {
    return i + sx<LOC>() * (j + sy<LOC>() * k);
  }
  // flat index inside the stacked vector
  template <int LOC> __host__ __device__ i64 idx(i64 i, i64 j, i64 k) const { return off<LOC>() + lidx<LOC>(i, j, k); }
};

// error plumbing (defined in api.cu)
void set_error(const char *msg);
int check_launch(const char *what);

static inline int valid_mesh(const DzbTensorMesh *m) {
  if (!m || m->nx <= 0 || m->ny <= 0 || m->nz <= 0) {
    set_error("invalid mesh (3-D TensorMesh with positive cell counts required)");
    return 0;
  }
  return 1;
}

__device__ __forceinline__ double ldg(const double *p) { return __ldg(p); }

}  // namespace dzb
