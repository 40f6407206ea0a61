/* Calling libdzb.so from plain C: the discrete divergence of a linear field on a non-uniform 3-D tensor mesh.
 *
 * The face vector holds u = (x, 2y, 3z) sampled at the face centres (normal components, the reference's layout
 * [Fx; Fy; Fz], x fastest).  The finite-volume divergence of a linear field is exact, so D u = 1 + 2 + 3 = 6 in every
 * cell up to rounding.  Everything the library touches is device memory; it never allocates or synchronises.
 *
 * build: gcc -std=c99 -Iinclude -I/usr/local/cuda/include examples/c_abi_divergence.c \
 *            -Ldiscretize_b200 -ldzb -L/usr/local/cuda/lib64 -lcudart -lm -Wl,-rpath,$PWD/discretize_b200 -o c_abi_divergence
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <cuda_runtime_api.h>

#include "dzb.h"

#define CHECK_CUDA(call)                                                          \
  do {                                                                            \
    cudaError_t e_ = (call);                                                      \
    if (e_ != cudaSuccess) {                                                      \
      fprintf(stderr, "CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); \
      return 2;                                                                   \
    }                                                                             \
  } while (0)

static double *upload(const double *host, size_t n) {
  double *dev = NULL;
  if (cudaMalloc((void **)&dev, n * sizeof(double)) != cudaSuccess) return NULL;
  if (cudaMemcpy(dev, host, n * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) return NULL;
  return dev;
}

int main(void) {
  enum { NX = 24, NY = 17, NZ = 20 };
  const int64_t n[3] = {NX, NY, NZ};
  double *h[3], *rh[3], *nodes[3];
  for (int a = 0; a < 3; ++a) { /* widths 1, 1.1, 1.2, ... scaled per axis */
    h[a] = malloc(n[a] * sizeof(double));
    rh[a] = malloc(n[a] * sizeof(double));
    nodes[a] = malloc((n[a] + 1) * sizeof(double));
    nodes[a][0] = 0.0;
    for (int64_t i = 0; i < n[a]; ++i) {
      h[a][i] = (1.0 + 0.1 * (double)(i % 7)) * (a + 1);
      rh[a][i] = 1.0 / h[a][i];
      nodes[a][i + 1] = nodes[a][i] + h[a][i];
    }
  }
  const size_t nC = (size_t)NX * NY * NZ;
  const size_t nFx = (size_t)(NX + 1) * NY * NZ, nFy = (size_t)NX * (NY + 1) * NZ, nFz = (size_t)NX * NY * (NZ + 1);
  double *u = malloc((nFx + nFy + nFz) * sizeof(double));
  size_t p = 0;
  for (int k = 0; k < NZ; ++k) for (int j = 0; j < NY; ++j) for (int i = 0; i <= NX; ++i) u[p++] = nodes[0][i];
  for (int k = 0; k < NZ; ++k) for (int j = 0; j <= NY; ++j) for (int i = 0; i < NX; ++i) u[p++] = 2.0 * nodes[1][j];
  for (int k = 0; k <= NZ; ++k) for (int j = 0; j < NY; ++j) for (int i = 0; i < NX; ++i) u[p++] = 3.0 * nodes[2][k];

  DzbTensorMesh mesh;
  mesh.nx = NX; mesh.ny = NY; mesh.nz = NZ;
  mesh.hx = upload(h[0], NX);   mesh.hy = upload(h[1], NY);   mesh.hz = upload(h[2], NZ);
  mesh.rhx = upload(rh[0], NX); mesh.rhy = upload(rh[1], NY); mesh.rhz = upload(rh[2], NZ);
  double *u_dev = upload(u, nFx + nFy + nFz), *div_dev = NULL;
  CHECK_CUDA(cudaMalloc((void **)&div_dev, nC * sizeof(double)));
  if (!mesh.hx || !mesh.hy || !mesh.hz || !mesh.rhx || !mesh.rhy || !mesh.rhz || !u_dev) {
    fprintf(stderr, "device allocation failed\n");
    return 2;
  }

  if (dzb_tm_apply(&mesh, DZB_OP_DIV, 0, u_dev, div_dev, NULL) != 0) {
    fprintf(stderr, "dzb_tm_apply: %s\n", dzb_last_error());
    return 1;
  }
  CHECK_CUDA(cudaDeviceSynchronize());
  double *div = malloc(nC * sizeof(double));
  CHECK_CUDA(cudaMemcpy(div, div_dev, nC * sizeof(double), cudaMemcpyDeviceToHost));
  double worst = 0.0;
  for (size_t c = 0; c < nC; ++c) worst = fmax(worst, fabs(div[c] - 6.0));
  printf("libdzb ABI %d: max |div u - 6| over %zu cells = %.3e\n", dzb_version(), nC, worst);
  return worst < 1e-11 ? 0 : 1;
}
