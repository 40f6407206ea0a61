// Elementary TensorMesh operators: y = A x, y = A^T x and CSR materialisation,
// all driven by the row generators in tm_rows.cuh (one thread per output row,
// 3-D thread blocks so that the y/z neighbours of a row are served from L1).
#include "tm_rows.cuh"
#include "tm_rowkernels.cuh"
#include "tm_march.cuh"

namespace dzb {

int g_apply_cfg = 2;
// marching kernels for DIV / GRAD / CURL and transposes: 0 = off (generic row kernels), -1 = per-operator tuned
// configuration (default), > 0 = one configuration for all six (sweep builds)
static int g_march_cfg = -1;

static bool try_march(int op, int transpose, int flags, const TM &m, const double *x, double *y, cudaStream_t s) {
  if (g_march_cfg != 0 && launch_c2f_march(op, transpose, flags, m, x, y, s)) return true;
  if (flags != 0) return false;
  switch (g_march_cfg) {
    case 0: return false;
#ifdef DZB_APPLY_SWEEP
    case 1: return launch_march<MarchCfg<4, 32, 0>>(op, transpose, m, x, y, s);
    case 2: return launch_march<MarchCfg<8, 32, 0>>(op, transpose, m, x, y, s);
    case 3: return launch_march<MarchCfg<8, 16, 0>>(op, transpose, m, x, y, s);
    case 4: return launch_march<MarchCfg<8, 16, 8>>(op, transpose, m, x, y, s);
    case 5: return launch_march<MarchCfg<8, 32, 8>>(op, transpose, m, x, y, s);
    case 6: return launch_march<MarchCfg<4, 32, 16>>(op, transpose, m, x, y, s);
    case 7: return launch_march<MarchCfg<4, 16, 16>>(op, transpose, m, x, y, s);
    case 8: return launch_march<MarchCfg<8, 8, 8>>(op, transpose, m, x, y, s);
    case 9: return launch_march<MarchCfg<16, 16, 4>>(op, transpose, m, x, y, s);
    case 10: return launch_march<MarchCfg<8, 16, 6>>(op, transpose, m, x, y, s);
    case 11: return launch_march<MarchCfg<8, 64, 8>>(op, transpose, m, x, y, s);
    case 12: return launch_march<MarchCfg<8, 16, 4>>(op, transpose, m, x, y, s);
    case 13: return launch_march<MarchCfg<8, 16, 2>>(op, transpose, m, x, y, s);
    case 14: return launch_march<MarchCfg<4, 16, 1>>(op, transpose, m, x, y, s);
#endif
    default: return launch_march_tuned(op, transpose, m, x, y, s);
  }
}

template <class R> static void shape_of(const TM &m, i64 *nr, i64 *nc) {
  *nr = R::nrows(m);
  *nc = R::ncols(m);
}

#define DZB_FOR_OP(OPID, BODY)                                                     \
  case OPID:                                                                       \
    if (transpose) { typedef Rows<OPID, true> R; BODY; }                           \
    else { typedef Rows<OPID, false> R; BODY; }                                    \
    break;

#define DZB_DISPATCH(BODY)                                                         \
  switch (op) {                                                                    \
    DZB_FOR_OP(DZB_OP_DIV, BODY) DZB_FOR_OP(DZB_OP_CURL, BODY) DZB_FOR_OP(DZB_OP_GRAD, BODY)            \
    DZB_FOR_OP(DZB_OP_AVE_F2CC, BODY) DZB_FOR_OP(DZB_OP_AVE_F2CCV, BODY) DZB_FOR_OP(DZB_OP_AVE_FX2CC, BODY) \
    DZB_FOR_OP(DZB_OP_AVE_FY2CC, BODY) DZB_FOR_OP(DZB_OP_AVE_FZ2CC, BODY) DZB_FOR_OP(DZB_OP_AVE_E2CC, BODY) \
    DZB_FOR_OP(DZB_OP_AVE_E2CCV, BODY) DZB_FOR_OP(DZB_OP_AVE_EX2CC, BODY) DZB_FOR_OP(DZB_OP_AVE_EY2CC, BODY) \
    DZB_FOR_OP(DZB_OP_AVE_EZ2CC, BODY) DZB_FOR_OP(DZB_OP_AVE_N2CC, BODY)                                 \
    DZB_FOR_OP(DZB_OP_AVE_CC2F, BODY) DZB_FOR_OP(DZB_OP_AVE_CCV2F, BODY) DZB_FOR_OP(DZB_OP_AVE_CC2E, BODY) \
    DZB_FOR_OP(DZB_OP_AVE_E2F, BODY) DZB_FOR_OP(DZB_OP_AVE_N2E, BODY) DZB_FOR_OP(DZB_OP_AVE_N2F, BODY)    \
    DZB_FOR_OP(DZB_OP_CELL_GRAD, BODY) DZB_FOR_OP(DZB_OP_CELL_GRAD_X, BODY) DZB_FOR_OP(DZB_OP_CELL_GRAD_Y, BODY) \
    DZB_FOR_OP(DZB_OP_CELL_GRAD_Z, BODY) DZB_FOR_OP(DZB_OP_DIV_X, BODY) DZB_FOR_OP(DZB_OP_DIV_Y, BODY)    \
    DZB_FOR_OP(DZB_OP_DIV_Z, BODY) DZB_FOR_OP(DZB_OP_GRAD_SQ, BODY) DZB_FOR_OP(DZB_OP_CURL_SQ, BODY)      \
    default:                                                                       \
      set_error("unknown operator id");                                            \
      return -1;                                                                   \
  }

}  // namespace dzb

using namespace dzb;

// tuning knob for experiments (not part of the stable ABI): launch configuration of the apply kernels
extern "C" void dzb_tune_elementary(int cfg) { dzb::g_apply_cfg = cfg; }
extern "C" void dzb_tune_march(int cfg) { dzb::g_march_cfg = cfg; }

extern "C" int dzb_tm_shape(const DzbTensorMesh *mm, int op, int64_t *nrows, int64_t *ncols) {
  if (!valid_mesh(mm)) return -1;
  TM m(*mm);
  const int transpose = 0;
  i64 nr = 0, nc = 0;
  DZB_DISPATCH(shape_of<R>(m, &nr, &nc));
  *nrows = nr;
  *ncols = nc;
  return 0;
}

template <class R> static R with_flags(int flags) {
  R gen;
  gen.flags = flags;
  return gen;
}

extern "C" int dzb_tm_apply_flags(const DzbTensorMesh *mm, int op, int transpose, int flags, const double *x, double *y,
                                  void *stream) {
  if (!valid_mesh(mm)) return -1;
  TM m(*mm);
  Args a{x, y, nullptr, nullptr, nullptr, nullptr};
  cudaStream_t s = (cudaStream_t)stream;
  if (try_march(op, transpose, flags, m, x, y, s)) return check_launch("dzb_tm_apply (march)");
  DZB_DISPATCH(launch_all(with_flags<R>(flags), MODE_APPLY, m, a, s));
  return check_launch("dzb_tm_apply");
}

extern "C" int dzb_tm_csr_count_flags(const DzbTensorMesh *mm, int op, int transpose, int flags, int64_t *row_nnz,
                                      void *stream) {
  if (!valid_mesh(mm)) return -1;
  TM m(*mm);
  Args a{nullptr, nullptr, (i64 *)row_nnz, nullptr, nullptr, nullptr};
  cudaStream_t s = (cudaStream_t)stream;
  DZB_DISPATCH(launch_all(with_flags<R>(flags), MODE_COUNT, m, a, s));
  return check_launch("dzb_tm_csr_count");
}

extern "C" int dzb_tm_csr_fill_flags(const DzbTensorMesh *mm, int op, int transpose, int flags, const int64_t *indptr,
                                     int64_t *indices, double *data, void *stream) {
  if (!valid_mesh(mm)) return -1;
  TM m(*mm);
  Args a{nullptr, nullptr, nullptr, (const i64 *)indptr, (i64 *)indices, data};
  cudaStream_t s = (cudaStream_t)stream;
  DZB_DISPATCH(launch_all(with_flags<R>(flags), MODE_FILL, m, a, s));
  return check_launch("dzb_tm_csr_fill");
}

extern "C" int dzb_tm_apply(const DzbTensorMesh *mm, int op, int transpose, const double *x, double *y, void *stream) {
  return dzb_tm_apply_flags(mm, op, transpose, 0, x, y, stream);
}

extern "C" int dzb_tm_csr_count(const DzbTensorMesh *mm, int op, int transpose, int64_t *row_nnz, void *stream) {
  return dzb_tm_csr_count_flags(mm, op, transpose, 0, row_nnz, stream);
}

extern "C" int dzb_tm_csr_fill(const DzbTensorMesh *mm, int op, int transpose, const int64_t *indptr, int64_t *indices,
                               double *data, void *stream) {
  return dzb_tm_csr_fill_flags(mm, op, transpose, 0, indptr, indices, data, stream);
}
