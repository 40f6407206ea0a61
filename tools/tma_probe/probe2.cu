// The CUDA programming guide's 2-D TMA example, verbatim in structure (libcu++ wrappers, int32 data).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda/barrier>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;

constexpr int GW = 1024, GH = 1024, SW = 32, SH = 32;

__global__ void kernel(const __grid_constant__ CUtensorMap tensor_map, int x, int y, int *out) {
  __shared__ alignas(128) int smem_buffer[SH][SW];
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ barrier bar;
  if (threadIdx.x == 0) {
    init(&bar, blockDim.x);
    cde::fence_proxy_async_shared_cta();
  }
  __syncthreads();
  barrier::arrival_token token;
  if (threadIdx.x == 0) {
    cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
    token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
  } else {
    token = bar.arrive();
  }
  bar.wait(std::move(token));
  for (int t = threadIdx.x; t < SW * SH; t += blockDim.x) out[t] = smem_buffer[t / SW][t % SW];
}

int main(int argc, char **argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  int *d, *out;
  cudaMalloc(&d, (size_t)GW * GH * 4);
  cudaMalloc(&out, SW * SH * 4);
  int *h = new int[(size_t)GW * GH];
  for (size_t t = 0; t < (size_t)GW * GH; ++t) h[t] = (int)t;
  cudaMemcpy(d, h, (size_t)GW * GH * 4, cudaMemcpyHostToDevice);
  void *p = nullptr;
  cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr);
  auto enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  CUtensorMap tm{};
  uint64_t size[2] = {GW, GH};
  uint64_t stride[1] = {GW * sizeof(int)};
  uint32_t box[2] = {SW, SH};
  uint32_t es[2] = {1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, d, size, stride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, (mode & 2) ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d\n", (int)r);
  const int x = (mode & 1) ? 3 : (mode & 4) ? -4 : (mode & 8) ? -1 : 64;
  kernel<<<1, 128>>>(tm, x, 32, out);
  cudaError_t e = cudaDeviceSynchronize();
  printf("probe2 mode %d: %s\n", mode, cudaGetErrorString(e));
  if (e == cudaSuccess) {
    int o[4];
    cudaMemcpy(o, out, 16, cudaMemcpyDeviceToHost);
    printf("o[0]=%d o[1]=%d (want %d)\n", o[0], o[1], 32 * GW + x);
  }
  return e != cudaSuccess;
}
