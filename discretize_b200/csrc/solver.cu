// Solver-side kernels: the vector work of a Jacobi-preconditioned conjugate-gradient iteration on the
// fused operators, with every scalar (alpha, beta, the dot products) kept ON THE DEVICE, so that an iteration is
// "operator apply + 3 launches" and the host only looks at the residual norm every few iterations.
//
//   scal[0] = r.z (current)   scal[1] = p.Ap   scal[2] = r.z (new)   scal[3] = r.r   scal[4..5] scratch
//
// Reductions are deterministic: every block writes one partial, the last block to finish (atomic ticket) sums the
// partials in block order.  Grid = a few waves of the 148 SMs, grid-stride loops, 8-byte coalesced accesses; each
// kernel is a pure stream over its vectors (HBM-bound: 16 B/entry for the dot, 64 B/entry for the update, 24 for
// the direction).
#include "tm_common.cuh"

namespace dzb {

constexpr int SOLVER_THREADS = 256;
constexpr int SOLVER_MAX_BLOCKS = 148 * 8;

__device__ __forceinline__ double block_sum(double v, double *smem) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) smem[w] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < SOLVER_THREADS / 32) t = smem[threadIdx.x];
  if (w == 0) {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  __syncthreads();
  return t;  // valid in thread 0
}

// the last block to arrive sums partials[0 .. gridDim.x) of each of the NRED quantities in block order
template <int NRED>
__device__ __forceinline__ void finish_reduction(const double (&mine)[NRED], double *partials, unsigned *ticket,
                                                 double *scal, const int (&slot)[NRED]) {
  __shared__ bool last;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < NRED; ++q) partials[q * SOLVER_MAX_BLOCKS + blockIdx.x] = mine[q];
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  __shared__ double smem[SOLVER_THREADS / 32];
#pragma unroll
  for (int q = 0; q < NRED; ++q) {
    double v = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += SOLVER_THREADS) v += partials[q * SOLVER_MAX_BLOCKS + b];
    v = block_sum(v, smem);
    if (threadIdx.x == 0) scal[slot[q]] = v;
  }
  if (threadIdx.x == 0) *ticket = 0u;  // ready for the next launch
}

// scal[slot] = a . b
__global__ void __launch_bounds__(SOLVER_THREADS)
    k_dot(i64 n, const double *__restrict__ a, const double *__restrict__ b, double *partials, unsigned *ticket,
          double *scal, int slot) {
  __shared__ double smem[SOLVER_THREADS / 32];
  double acc = 0.0;
  const i64 stride = (i64)gridDim.x * SOLVER_THREADS;
  for (i64 t = (i64)blockIdx.x * SOLVER_THREADS + threadIdx.x; t < n; t += stride) acc += __ldg(a + t) * __ldg(b + t);
  const double s = block_sum(acc, smem);
  const double mine[1] = {s};
  const int slots[1] = {slot};
  finish_reduction<1>(mine, partials, ticket, scal, slots);
}

// alpha = scal[0] / scal[1];  x += alpha p;  r -= alpha Ap;  z = dinv ? dinv * r : r;  scal[2] = r.z, scal[3] = r.r
__global__ void __launch_bounds__(SOLVER_THREADS)
    k_cg_update(i64 n, double *__restrict__ x, double *r, const double *__restrict__ p,
                const double *__restrict__ Ap, const double *__restrict__ dinv, double *z,
                double *partials, unsigned *ticket, double *scal) {
  __shared__ double smem[SOLVER_THREADS / 32];
  // breakdown guard: once r.z or p.Ap is zero (exact solution reached inside a blind window of iterations, or an
  // underflow) the update becomes a no-op instead of 0/0 = NaN poisoning x, r and p
  const double rz0 = scal[0], pAp = scal[1];
  const double alpha = (rz0 == 0.0 || pAp == 0.0) ? 0.0 : rz0 / pAp;
  double rz = 0.0, rr = 0.0;
  const i64 stride = (i64)gridDim.x * SOLVER_THREADS;
  for (i64 t = (i64)blockIdx.x * SOLVER_THREADS + threadIdx.x; t < n; t += stride) {
    x[t] += alpha * __ldg(p + t);
    const double rt = r[t] - alpha * __ldg(Ap + t);
    r[t] = rt;
    const double zt = dinv ? __ldg(dinv + t) * rt : rt;
    if (z != r) z[t] = zt;
    rz += rt * zt;
    rr += rt * rt;
  }
  const double s0 = block_sum(rz, smem);
  const double s1 = block_sum(rr, smem);
  const double mine[2] = {s0, s1};
  const int slots[2] = {2, 3};
  finish_reduction<2>(mine, partials, ticket, scal, slots);
}

// beta = scal[2] / scal[0];  p = z + beta p;  then scal[0] = scal[2]
__global__ void __launch_bounds__(SOLVER_THREADS)
    k_cg_direction(i64 n, double *__restrict__ p, const double *__restrict__ z, double *scal, unsigned *ticket) {
  const double beta = scal[0] == 0.0 ? 0.0 : scal[2] / scal[0];   // converged: keep p finite (see k_cg_update)
  const i64 stride = (i64)gridDim.x * SOLVER_THREADS;
  for (i64 t = (i64)blockIdx.x * SOLVER_THREADS + threadIdx.x; t < n; t += stride) p[t] = __ldg(z + t) + beta * p[t];
  // every block has read scal[0] before the last one overwrites it
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    if (last) {
      scal[0] = scal[2];
      *ticket = 0u;
    }
  }
}

static unsigned solver_blocks(i64 n) {
  const i64 need = (n + SOLVER_THREADS - 1) / SOLVER_THREADS;
  return (unsigned)(need < SOLVER_MAX_BLOCKS ? (need > 0 ? need : 1) : SOLVER_MAX_BLOCKS);
}

}  // namespace dzb

using namespace dzb;

// workspace: (2 * SOLVER_MAX_BLOCKS) doubles of partials followed by 2 unsigned tickets (zero-initialised by the caller)
extern "C" int64_t dzb_solver_workspace_bytes(void) { return (int64_t)(2 * SOLVER_MAX_BLOCKS * sizeof(double) + 16); }

static inline unsigned *ws_ticket(void *ws, int which) {
  return (unsigned *)((char *)ws + 2 * SOLVER_MAX_BLOCKS * sizeof(double)) + which;
}

extern "C" int dzb_dot(int64_t n, const double *a, const double *b, void *workspace, double *scal, int slot, void *stream) {
  k_dot<<<solver_blocks(n), SOLVER_THREADS, 0, (cudaStream_t)stream>>>(n, a, b, (double *)workspace, ws_ticket(workspace, 0),
                                                                        scal, slot);
  return check_launch("dzb_dot");
}

extern "C" int dzb_cg_update(int64_t n, double *x, double *r, const double *p, const double *Ap, const double *dinv,
                             double *z, void *workspace, double *scal, void *stream) {
  k_cg_update<<<solver_blocks(n), SOLVER_THREADS, 0, (cudaStream_t)stream>>>(n, x, r, p, Ap, dinv, z, (double *)workspace,
                                                                              ws_ticket(workspace, 0), scal);
  return check_launch("dzb_cg_update");
}

extern "C" int dzb_cg_direction(int64_t n, double *p, const double *z, void *workspace, double *scal, void *stream) {
  k_cg_direction<<<solver_blocks(n), SOLVER_THREADS, 0, (cudaStream_t)stream>>>(n, p, z, scal, ws_ticket(workspace, 1));
  return check_launch("dzb_cg_direction");
}
