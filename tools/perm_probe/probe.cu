// How fast can a fixed random permutation of 1e8 doubles be applied?  (un-permute step of the bucketed interpolation)
//   A: one pass of random reads   out[q] = s[inv[q]]                      (what k_gather_rows does), unroll 4 / 8 / 16
//   B: two passes through blocks of BS destinations:  t[pos1[i]] = s[i]  (scattered 8-byte writes into ~N/BS open
//      streams, merged in L2), then out[q] = t[blk(q)*BS + loc[q]] (random reads inside a window that sits in L2)
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o probe probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
typedef long long i64;

template <int U>
__global__ void __launch_bounds__(256) k_gather(i64 n, const double* __restrict__ src, const int* __restrict__ idx, double* __restrict__ dst) {
  const i64 stride = (i64)gridDim.x * blockDim.x;
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (U - 1) * stride < n; i += U * stride) {
    int a[U]; double v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) a[u] = __ldg(idx + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = __ldg(src + a[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) __stcs(dst + i + u * stride, v[u]);
  }
  for (; i < n; i += stride) dst[i] = __ldg(src + __ldg(idx + i));
}
__global__ void __launch_bounds__(256) k_scatter(i64 n, const double* __restrict__ src, const int* __restrict__ pos, double* __restrict__ dst) {
  const i64 stride = (i64)gridDim.x * blockDim.x;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[__ldg(pos + i)] = __ldg(src + i);
}
// stage 2: a block of threads owns whole destination blocks, so that its reads stay inside one window
template <typename LOC>
__global__ void __launch_bounds__(256) k_local(i64 n, int bs, const double* __restrict__ t, const LOC* __restrict__ loc, double* __restrict__ dst) {
  const i64 nblk = (n + bs - 1) / bs;
  for (i64 b = blockIdx.x; b < nblk; b += gridDim.x) {
    const i64 lo = b * bs, hi = min(n, lo + bs);
    for (i64 q = lo + threadIdx.x; q < hi; q += 4 * blockDim.x) {
      double v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { const i64 qq = q + u * blockDim.x; v[u] = qq < hi ? __ldg(t + lo + (i64)loc[qq]) : 0.0; }
#pragma unroll
      for (int u = 0; u < 4; ++u) { const i64 qq = q + u * blockDim.x; if (qq < hi) __stcs(dst + qq, v[u]); }
    }
  }
}
static float timeit(void (*fn)(void*), void* ctx, int reps = 5) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  fn(ctx); fn(ctx); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) { cudaEventRecord(a); fn(ctx); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); best = std::min(best, ms); }
  return best;
}
struct Ctx { i64 n; double *s, *t, *out; int *inv, *pos1; void* loc; int bs, blocks; };
template <int U> static void runA(void* p) { Ctx* c = (Ctx*)p; k_gather<U><<<c->blocks, 256>>>(c->n, c->s, c->inv, c->out); }
static void runS(void* p) { Ctx* c = (Ctx*)p; k_scatter<<<148 * 32, 256>>>(c->n, c->s, c->pos1, c->t); }
static void runL16(void* p) { Ctx* c = (Ctx*)p; k_local<unsigned short><<<c->blocks, 256>>>(c->n, c->bs, c->t, (unsigned short*)c->loc, c->out); }
static void runL32(void* p) { Ctx* c = (Ctx*)p; k_local<int><<<c->blocks, 256>>>(c->n, c->bs, c->t, (int*)c->loc, c->out); }

int main(int argc, char** argv) {
  const i64 n = argc > 1 ? atoll(argv[1]) : 100000000LL;
  std::vector<int> inv(n);
  for (i64 i = 0; i < n; ++i) inv[i] = (int)i;
  uint64_t st = 88172645463325252ULL;
  for (i64 i = n - 1; i > 0; --i) { st ^= st << 13; st ^= st >> 7; st ^= st << 17; std::swap(inv[i], inv[st % (uint64_t)(i + 1)]); }
  std::vector<double> s(n);
  for (i64 i = 0; i < n; ++i) s[i] = (double)i;
  Ctx c; c.n = n;
  cudaMalloc(&c.s, n * 8); cudaMalloc(&c.t, n * 8); cudaMalloc(&c.out, n * 8); cudaMalloc(&c.inv, n * 4); cudaMalloc(&c.pos1, n * 4); cudaMalloc(&c.loc, n * 4);
  cudaMemcpy(c.s, s.data(), n * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(c.inv, inv.data(), n * 4, cudaMemcpyHostToDevice);
  for (int waves : {8, 16, 32}) {
    c.blocks = 148 * waves;
    printf("A gather waves %2d: U4 %.3f  U8 %.3f  U16 %.3f ms\n", waves, timeit(runA<4>, &c), timeit(runA<8>, &c), timeit(runA<16>, &c));
  }
  std::vector<double> out(n);
  cudaMemcpy(out.data(), c.out, n * 8, cudaMemcpyDeviceToHost);
  for (i64 q = 0; q < n; q += 99991) if (out[q] != (double)inv[q]) { printf("A wrong at %lld\n", q); return 1; }
  std::vector<int> perm(n), pos1(n), loc(n);
  for (i64 q = 0; q < n; ++q) perm[inv[q]] = (int)q;
  for (int bs : {1 << 14, 1 << 16, 1 << 18, 1 << 20}) {
    const i64 nblk = (n + bs - 1) / bs;
    std::vector<i64> fill(nblk, 0);
    for (i64 i = 0; i < n; ++i) { const i64 b = perm[i] / bs; pos1[i] = (int)(b * bs + fill[b]++); }
    for (i64 q = 0; q < n; ++q) loc[q] = pos1[inv[q]] - (int)((q / bs) * bs);
    cudaMemcpy(c.pos1, pos1.data(), n * 4, cudaMemcpyHostToDevice);
    c.bs = bs;
    float tl;
    if (bs <= 65536) {
      std::vector<unsigned short> l16(n);
      for (i64 q = 0; q < n; ++q) l16[q] = (unsigned short)loc[q];
      cudaMemcpy(c.loc, l16.data(), n * 2, cudaMemcpyHostToDevice);
      c.blocks = 148 * 8;
      tl = timeit(runL16, &c);
    } else {
      cudaMemcpy(c.loc, loc.data(), n * 4, cudaMemcpyHostToDevice);
      c.blocks = 148 * 8;
      tl = timeit(runL32, &c);
    }
    const float ts = timeit(runS, &c);
    runS(&c); if (bs <= 65536) runL16(&c); else runL32(&c);
    cudaMemcpy(out.data(), c.out, n * 8, cudaMemcpyDeviceToHost);
    bool ok = true;
    for (i64 q = 0; q < n; q += 9973) if (out[q] != (double)inv[q]) { ok = false; break; }
    printf("B bs %8d: scatter %.3f + local gather %.3f = %.3f ms  %s\n", bs, ts, tl, ts + tl, ok ? "ok" : "WRONG");
  }
  return 0;
}
