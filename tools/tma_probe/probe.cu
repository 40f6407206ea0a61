// Minimal tensor-map TMA probe (FLOAT64, 2-D):  ./probe mode
//   bit 0: prefetch.tensormap first   bit 1: negative start coordinate   bit 2: descriptor from a global buffer instead of the kernel parameter
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Maps { CUtensorMap m[2]; };

__global__ void k(const __grid_constant__ Maps maps, const CUtensorMap *gmap, double *out, int W, int H, int x0, int y0, int mode) {
  extern __shared__ unsigned char raw[];
  double *tile = reinterpret_cast<double *>((reinterpret_cast<uintptr_t>(raw) + 127) & ~(uintptr_t)127);
  uint64_t *bar = reinterpret_cast<uint64_t *>(tile + W * H);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const CUtensorMap *mp = (mode & 4) ? gmap : &maps.m[1];
  if (threadIdx.x == 0) {
    if (mode & 1) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(mp)) : "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(W * H * 8) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(tile)),
        "l"(reinterpret_cast<uint64_t>(mp)), "r"(x0), "r"(y0), "r"(smem_u32(bar))
        : "memory");
  }
  asm volatile(
      "{\n.reg .pred p;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(
          smem_u32(bar)),
      "r"(0)
      : "memory");
  for (int t = threadIdx.x; t < W * H; t += blockDim.x) out[t] = tile[t];
}

int main(int argc, char **argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  const int P = 1024, ROWS = 64, W = (mode & 16) ? 16 : 64, H = (mode & 16) ? 4 : 6;
  const int f32 = (mode & 8) ? 2 : 1;   // describe the doubles as pairs of 32-bit elements
  std::vector<double> h((size_t)P * ROWS);
  for (size_t t = 0; t < h.size(); ++t) h[t] = (double)t;
  double *d, *out;
  cudaMalloc(&d, h.size() * 8);
  cudaMalloc(&out, W * H * 8);
  cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  void *p = nullptr;
  cudaDriverEntryPointQueryResult qr;
  cudaError_t ge = (mode & 32) ? cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &qr)
                               : cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr);
  if (ge != cudaSuccess || !p) { printf("no encoder\n"); return 2; }
  auto enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(p);
  Maps maps;
  cuuint64_t gdim[2] = {(cuuint64_t)P * f32, (cuuint64_t)ROWS};
  cuuint64_t gstr[1] = {(cuuint64_t)P * 8};
  cuuint32_t box[2] = {(cuuint32_t)W * f32, (cuuint32_t)H};
  const CUtensorMapDataType dt = (mode & 8) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : (mode & 64) ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
  cuuint32_t estr[2] = {1, 1};
  for (int a = 0; a < 2; ++a) {
    CUresult r = enc(&maps.m[a], dt, 2, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 3; }
  }
  CUtensorMap *gmap;
  cudaMalloc(&gmap, sizeof(CUtensorMap));
  cudaMemcpy(gmap, &maps.m[1], sizeof(CUtensorMap), cudaMemcpyHostToDevice);
  const int x0 = (mode & 2) ? -1 : 3, y0 = (mode & 2) ? -1 : 5;
  { const unsigned long long *w = reinterpret_cast<const unsigned long long *>(&maps.m[1]); printf("map: %016llx %016llx %016llx %016llx\n", w[0], w[1], w[2], w[3]); }
  k<<<1, 128, W * H * 8 + 256>>>(maps, gmap, out, W, H, x0 * f32, y0, mode);
  cudaError_t e = cudaDeviceSynchronize();
  printf("mode %d: %s\n", mode, cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<double> o(W * H);
  cudaMemcpy(o.data(), out, W * H * 8, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int r = 0; r < H; ++r)
    for (int c = 0; c < W; ++c) {
      const int gx = x0 + c, gy = y0 + r;
      const double want = (gx < 0 || gy < 0 || gx >= P || gy >= ROWS) ? 0.0 : (double)((size_t)gy * P + gx);
      if (o[r * W + c] != want) ++bad;
    }
  printf("mode %d: mismatches %d (o[0]=%g o[last]=%g)\n", mode, bad, o[0], o[W * H - 1]);
  return bad != 0;
}
