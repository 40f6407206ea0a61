/*
 * dzb.h -- C ABI of libdzb.so, the B200 (sm_100a) operator-application library.
 *
 * Drop-in boundary for the operator-application hot path of simpeg/discretize.
 * The reference has no FFI on this path: its operators are scipy.sparse matrices
 * built by python properties and applied by scipy's csr_matvec.  Each entry point
 * below replaces "build that matrix, then A @ x" for one reference property; the
 * reference interface it replaces is cited as file:line (relative to the
 * reference tree).  INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Rules of the ABI
 *   - plain C: pointers, sizes, ints.  No torch / C++ types cross it.
 *   - every pointer is a DEVICE pointer unless the name ends in _host.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *   - nothing allocates: the caller owns inputs, outputs and workspaces.
 *   - nothing synchronises: kernels are enqueued on `stream` and the call returns.
 *   - return 0 on success, <0 on error; dzb_last_error() gives the message.
 *   - all reals are float64; vector layouts are the reference's (x fastest,
 *     faces = [Fx;Fy;Fz], edges = [Ex;Ey;Ez]; base/base_regular_mesh.py:601-762).
 */
#ifndef DZB_H
#define DZB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DZB_ABI_VERSION 1

/* 3-D tensor mesh: cell counts and DEVICE pointers to the 1-D cell widths and
 * their reciprocals (reference: base/base_tensor_mesh.py:86-114 `h`). */
typedef struct {
  int64_t nx, ny, nz;
  const double *hx, *hy, *hz;    /* lengths nx, ny, nz */
  const double *rhx, *rhy, *rhz; /* 1/h, same lengths  */
} DzbTensorMesh;

/* Elementary operators (all 3-D TensorMesh). */
enum DzbOp {
  DZB_OP_DIV = 0,        /* face_divergence            operators/differential_operators.py:225-235 */
  DZB_OP_CURL = 1,       /* edge_curl                  :2169-2181 */
  DZB_OP_GRAD = 2,       /* nodal_gradient             :658-664   */
  DZB_OP_AVE_F2CC = 3,   /* average_face_to_cell       :2478-2492 */
  DZB_OP_AVE_F2CCV = 4,  /* average_face_to_cell_vector:2494-2508 */
  DZB_OP_AVE_FX2CC = 5,  /* average_face_x_to_cell     :2510-2787 */
  DZB_OP_AVE_FY2CC = 6,
  DZB_OP_AVE_FZ2CC = 7,
  DZB_OP_AVE_E2CC = 8,   /* average_edge_to_cell       :2894-2908 */
  DZB_OP_AVE_E2CCV = 9,  /* average_edge_to_cell_vector:2910-2924 */
  DZB_OP_AVE_EX2CC = 10, /* average_edge_x_to_cell     :2926-3203 */
  DZB_OP_AVE_EY2CC = 11,
  DZB_OP_AVE_EZ2CC = 12,
  DZB_OP_AVE_N2CC = 13,  /* average_node_to_cell       :3236-3253 */
  /* cell-centred side */
  DZB_OP_AVE_CC2F = 14,  /* average_cell_to_face        :2789-2828 */
  DZB_OP_AVE_CCV2F = 15, /* average_cell_vector_to_face :2830-2865 */
  DZB_OP_AVE_CC2E = 16,  /* average_cell_to_edge        :2867-2892 */
  DZB_OP_AVE_E2F = 17,   /* average_edge_to_face        :3205-3234 */
  DZB_OP_AVE_N2E = 18,   /* average_node_to_edge        :3254-3319 */
  DZB_OP_AVE_N2F = 19,   /* average_node_to_face        :3321-3382 */
  DZB_OP_CELL_GRAD = 20, /* stencil_cell_gradient :1413-1443 / cell_gradient :1445-1595 (flags: DzbCellGradFlags) */
  DZB_OP_CELL_GRAD_X = 21, /* stencil_cell_gradient_x / cell_gradient_x :1049-1181, 1762-1877 (always Neumann) */
  DZB_OP_CELL_GRAD_Y = 22,
  DZB_OP_CELL_GRAD_Z = 23,
  DZB_OP_DIV_X = 24,     /* face_x_divergence           :236-348 */
  DZB_OP_DIV_Y = 25,     /* face_y_divergence           :350-462 */
  DZB_OP_DIV_Z = 26,     /* face_z_divergence           :464-591 */
  /* entrywise squares: diag(G^T W G) = (G o G)^T w, diag(C^T W C) = (C o C)^T w (Jacobi preconditioners) */
  DZB_OP_GRAD_SQ = 27,
  DZB_OP_CURL_SQ = 28,
  DZB_OP_COUNT = 29
};

/* `flags` of the DZB_OP_CELL_GRAD* operators (set_cell_gradient_BC, :975-1046; _ddxCellGrad, :41-82).
 * Default (0) is zero Neumann on all six sides and the unscaled stencil. */
enum DzbCellGradFlags {
  DZB_CG_DIRICHLET_X_LO = 1, DZB_CG_DIRICHLET_X_HI = 2,
  DZB_CG_DIRICHLET_Y_LO = 4, DZB_CG_DIRICHLET_Y_HI = 8,
  DZB_CG_DIRICHLET_Z_LO = 16, DZB_CG_DIRICHLET_Z_HI = 32,
  DZB_CG_SCALE = 64 /* cell_gradient = diag(S / aveCC2F V) stencil; Neumann boundary rows are then empty */
};

/* Physical-property layouts (utils/matrix_utils.py:1048-1068 TensorType). */
enum DzbModelKind {
  DZB_MODEL_SCALAR = 0, /* one value, passed in `model_scalar`              */
  DZB_MODEL_ISO = 1,    /* nC values                                        */
  DZB_MODEL_ANISO = 2,  /* 3*nC values, columns xx,yy,zz (Fortran order)    */
  DZB_MODEL_TENSOR = 3  /* 6*nC values, columns xx,yy,zz,xy,xz,yz           */
};

enum DzbFlags {
  DZB_INVERT_MODEL = 1,  /* use 1/model (3x3 inverse for TENSOR)             */
  DZB_INVERT_MATRIX = 2, /* use the reciprocal of the (diagonal) mass matrix */
  DZB_EDGES = 4          /* edge inner product; default is faces             */
};

int dzb_version(void);
const char *dzb_last_error(void);

/* ---- elementary operators --------------------------------------------------- */
/* y = A x (transpose == 0) or y = A^T x (transpose != 0), A = mesh.<op>. */
int dzb_tm_apply(const DzbTensorMesh *m, int op, int transpose, const double *x, double *y, void *stream);
/* Shape of mesh.<op> (host call). */
int dzb_tm_shape(const DzbTensorMesh *m_host_counts_only, int op, int64_t *nrows, int64_t *ncols);
/* CSR materialisation of mesh.<op> (or its transpose): pass 1 writes the number of
 * stored entries of each row into row_nnz[nrows]; the caller turns that into indptr
 * (exclusive scan); pass 2 writes sorted column indices and values. */
int dzb_tm_csr_count(const DzbTensorMesh *m, int op, int transpose, int64_t *row_nnz, void *stream);
int dzb_tm_csr_fill(const DzbTensorMesh *m, int op, int transpose, const int64_t *indptr, int64_t *indices,
                    double *data, void *stream);
/* The same three calls for operators that take runtime `flags` (DzbCellGradFlags); flags = 0 is the call above. */
int dzb_tm_apply_flags(const DzbTensorMesh *m, int op, int transpose, int flags, const double *x, double *y,
                       void *stream);
int dzb_tm_csr_count_flags(const DzbTensorMesh *m, int op, int transpose, int flags, int64_t *row_nnz, void *stream);
int dzb_tm_csr_fill_flags(const DzbTensorMesh *m, int op, int transpose, int flags, const int64_t *indptr,
                          int64_t *indices, double *data, void *stream);

/* ---- mass matrices: get_face_inner_product / get_edge_inner_product ---------- */
/* operators/inner_products.py:35-99 -> base/base_tensor_mesh.py:792-863 (diagonal)
 * and inner_products.py:147-272 (full tensor).  flags: DzbFlags. */
/* diag[n] = diagonal of M(model)  (SCALAR / ISO / ANISO only). */
int dzb_tm_mass_diag(const DzbTensorMesh *m, int flags, int model_kind, const double *model, double model_scalar,
                     double *diag, void *stream);
/* y = M(model) x for any model kind, generated on the fly (no stored matrix). */
int dzb_tm_mass_apply(const DzbTensorMesh *m, int flags, int model_kind, const double *model, double model_scalar,
                      const double *x, double *y, void *stream);
/* out = the inverses of the n_cells symmetric 3x3 tensors of a full-tensor property (six columns xx, yy, zz, xy, xz, yz;
 * closed form of utils/matrix_utils.py:734-754): `invert_model` for MODEL_TENSOR done once, so that dzb_tm_mass_apply
 * can take its marching kernels on the result (without DZB_INVERT_MODEL). */
int dzb_tm_tensor_invert(int64_t n_cells, const double *model, double *out, void *stream);
/* CSR of the full-tensor mass matrix (<= 9 entries per row). */
int dzb_tm_mass_tensor_csr_count(const DzbTensorMesh *m, int flags, int64_t *row_nnz, void *stream);
int dzb_tm_mass_tensor_csr_fill(const DzbTensorMesh *m, int flags, const double *model, const int64_t *indptr,
                                int64_t *indices, double *data, void *stream);

/* ---- model derivatives: get_{face,edge}_inner_product_deriv ------------------- */
/* inner_products.py:274-326, 405-565; base_tensor_mesh.py:865-988.
 * F(v) = d(M(model) v)/d(model) is n x (nC*k), k = 1 (SCALAR: n x 1; ISO), 3, 6.
 * transpose == 0: y[n]      = F(v) dm      (dm has nC*k entries, 1 for SCALAR)
 * transpose != 0: y[nC*k]   = F(v)^T w     (w has n entries). */
int dzb_tm_mass_deriv_apply(const DzbTensorMesh *m, int flags, int model_kind, const double *model,
                            double model_scalar, const double *v, int transpose, const double *x, double *y,
                            void *stream);

/* ---- fused composites -------------------------------------------------------- */
/* y = G^T Me(sigma) G phi  (DC resistivity, nodal; tutorials/pde/3_nodal_dirichlet_poisson.py:146-160):
 * nodal_gradient (:658-664) + get_edge_inner_product (isotropic sigma, nC) in ONE kernel.
 * phi and sigma only need their natural 8-byte alignment and nothing outside phi[0, nN) / sigma[0, nC) is read
 * (rows are staged with 16-byte bulk copies that stop at the aligned bounds inside the arrays; the at most
 * one element left at an odd-aligned end is staged by a plain load). */
int dzb_tm_dc_nodal(const DzbTensorMesh *m, const double *sigma, const double *phi, double *y, void *stream);
/* A z-slab's apply with PEER-READ halos (one process per GPU, NVLink): phi on the local node planes lo_plane / hi_plane
 * is fetched by the kernel's producer warp from the neighbours' local vectors phi_lo / phi_hi (peer memory mapped into
 * this process; NULL: no such neighbour) at plane lo_peer_plane / hi_peer_plane of those arrays, which hold nz_lo + 1 /
 * nz_hi + 1 node planes of the same nx, ny.  ONE launch per apply and rank: no send / recv, no separate launches for the
 * planes next to a halo.  Returns -2 where the staged-ring kernel does not apply (use dzb_tm_dc_nodal_range + an exchange).
 * Ordering after the neighbours' writes to their phi: sync == NULL -- the caller's job (a barrier across the ranks in
 * front of the call); else IN the kernel: sync = 5 zero-initialised words of this rank ([0] / [1] written by the lower /
 * upper neighbour, [2] launches completed, [3] scratch, [4] set if a wait gave up after 2 s), lo_signal / hi_signal = the
 * addresses of the lower neighbour's sync[1] / the upper neighbour's sync[0] (peer memory; NULL: no such neighbour).  Each
 * launch announces "phi of epoch sync[2] + 1 is complete" to both neighbours and its producer warps wait for the
 * neighbours' announcement of the same epoch before they copy a halo plane, so every rank must call this the same number
 * of times; nothing but the stream order of each rank is needed around the call. */
int dzb_tm_dc_nodal_range_peer(const DzbTensorMesh *m, const double *sigma, const double *phi, const double *phi_lo,
                               int64_t nz_lo, int64_t lo_plane, int64_t lo_peer_plane, const double *phi_hi,
                               int64_t nz_hi, int64_t hi_plane, int64_t hi_peer_plane, uint32_t *sync,
                               uint32_t *lo_signal, uint32_t *hi_signal, double *y, int64_t k_begin, int64_t k_end,
                               void *stream);
/* The same operator on nrhs right-hand sides in ONE launch (multi-source DC surveys: scipy's matmat): right-hand side r is
 * phi + r * ldx, its result y + r * ldx (ldx >= nN).  sigma is read from DRAM once per tile, not once per right-hand
 * side: the nrhs blocks that stage the same tile are scheduled side by side and share it through L2. */
int dzb_tm_dc_nodal_multi(const DzbTensorMesh *m, const double *sigma, const double *phi, double *y, int nrhs, int64_t ldx,
                          void *stream);
/* y = A phi and dot[0] = phi . y in ONE pass over the data (the p.Ap of a conjugate-gradient iteration): the
 * kernel's consumers hold phi and y of every node they write.  workspace: scratch for the per-(block, warp) partials,
 * workspace_doubles >= blocks x consumer warps of the launch (2^20 doubles cover 512^3); the partials are summed in index
 * order (deterministic).  dot: one device double.  Returns -2 where the staged-ring kernel does not apply or the
 * workspace is too small (then: dzb_tm_dc_nodal + dzb_dot). */
int dzb_tm_dc_nodal_dot(const DzbTensorMesh *m, const double *sigma, const double *phi, double *y, double *workspace,
                        int64_t workspace_doubles, double *dot, void *stream);
/* Same, writing only the node planes k_begin <= k < k_end of y (0 <= k <= nz).  This is the
 * building block of the z-slab decomposition: a rank applies it to its local sub-mesh (its cell
 * layers plus one halo layer / node plane per neighbour) for the planes it owns, interior
 * planes first while the halo planes are still in flight. */
int dzb_tm_dc_nodal_range(const DzbTensorMesh *m, const double *sigma, const double *phi, double *y, int64_t k_begin,
                          int64_t k_end, void *stream);

/* y = D Mf(rho)^-1 D^T x  (DC resistivity, cell-centred: examples/plot_dc_resistivity.py:23-30;
 * face_divergence :225-235 + get_face_inner_product(rho, invert_matrix=True), isotropic rho of nC values) in ONE
 * kernel.  flags: DZB_DC_CELL_NEUMANN drops the boundary faces (zero-flux boundary) -- default keeps them, which is
 * what the plain product D Mf^-1 D^T does (homogeneous Dirichlet ghost values). */
enum DzbDcCellFlags { DZB_DC_CELL_NEUMANN = 1 };
int dzb_tm_dc_cell(const DzbTensorMesh *m, const double *rho, int flags, const double *x, double *y, void *stream);
/* The same with mi[nF] = 1 / diag(Mf(rho)) precomputed (dzb_tm_mass_diag with DZB_INVERT_MATRIX): no divisions in the
 * apply, 40 B per cell. */
int dzb_tm_dc_cell_mi(const DzbTensorMesh *m, const double *mi, int flags, const double *x, double *y, void *stream);

/* y = C^T Mf(mui) C e + Me(sigma) e  (Maxwell, E-formulation;
 * tests/boundaries/test_boundary_maxwell.py:95-99): edge_curl (:2169-2181) +
 * get_face_inner_product(mui) + get_edge_inner_product(sigma), isotropic models. */
int dzb_tm_maxwell(const DzbTensorMesh *m, const double *mui, const double *sigma, const double *e, double *y,
                   void *stream);
/* Same, writing only what lives on node planes / cell layers k_begin <= k < k_end (Ex, Ey on planes k,
 * Ez in layers k < nz): the z-slab building block, like dzb_tm_dc_nodal_range.  Nothing outside
 * e[0, nE), mui[0, nC), sigma[0, nC) is read. */
int dzb_tm_maxwell_range(const DzbTensorMesh *m, const double *mui, const double *sigma, const double *e, double *y,
                         int64_t k_begin, int64_t k_end, void *stream);
/* Two disjoint plane ranges in one launch: the two boundary planes of a z-slab once its halos have arrived. */
int dzb_tm_maxwell_range2(const DzbTensorMesh *m, const double *mui, const double *sigma, const double *e, double *y,
                          int64_t k_begin, int64_t k_end, int64_t k_begin2, int64_t k_end2, void *stream);
/* A z-slab's apply with PEER-READ halos (one process per GPU, NVLink): ex, ey on the local node planes lo_plane /
 * hi_plane and ez in the local layer lo_plane are fetched by the kernel itself from the neighbours' local edge vectors
 * e_lo / e_hi (peer memory mapped into this process; NULL: no such neighbour) at plane lo_peer_plane / hi_peer_plane of
 * those arrays, laid out like a TensorMesh with nz_lo / nz_hi cell layers.  ONE launch per apply and rank: no send / recv,
 * no separate launch for the planes next to a halo.  The caller orders the call after the neighbours' writes to their e.
 * Returns -2 where the tensor-map path does not apply (use dzb_tm_maxwell_range + an exchange there). */
int dzb_tm_maxwell_range_peer(const DzbTensorMesh *m, const double *mui, const double *sigma, const double *e,
                              const double *e_lo, int64_t nz_lo, int64_t lo_plane, int64_t lo_peer_plane,
                              const double *e_hi, int64_t nz_hi, int64_t hi_plane, int64_t hi_peer_plane, double *y,
                              int64_t k_begin, int64_t k_end, void *stream);
/* The same with the ordering against the neighbours' writes done INSIDE the kernel: sync / lo_signal / hi_signal as for
 * dzb_tm_dc_nodal_range_peer (5 zero-initialised words of this rank; the lower neighbour's sync[1], the upper neighbour's
 * sync[0]).  No barrier, no other launch between two distributed applies; every rank calls it the same number of times. */
int dzb_tm_maxwell_range_peer_sync(const DzbTensorMesh *m, const double *mui, const double *sigma, const double *e,
                                   const double *e_lo, int64_t nz_lo, int64_t lo_plane, int64_t lo_peer_plane,
                                   const double *e_hi, int64_t nz_hi, int64_t hi_plane, int64_t hi_peer_plane,
                                   uint32_t *sync, uint32_t *lo_signal, uint32_t *hi_signal, double *y, int64_t k_begin,
                                   int64_t k_end, void *stream);
/* y = C^T Mf(mui) C e + (alpha_re + i alpha_im) Me(sigma) e for a COMPLEX edge field in ONE launch: e and y are
 * interleaved (re, im) arrays of 2 nE doubles (numpy / torch complex128).  This is the frequency-domain Maxwell operator
 * SimPEG applies (alpha = i omega).  Face weights and edge masses are computed once for both parts.
 * Returns -2 where the tensor-map path does not apply (e, mui, sigma not 16-byte aligned). */
int dzb_tm_maxwell_cplx(const DzbTensorMesh *m, const double *mui, const double *sigma, double alpha_re, double alpha_im,
                        const double *e, double *y, void *stream);

/* ---- interpolation: get_interpolation_matrix --------------------------------- */
/* base/base_tensor_mesh.py:684-790 -> utils/interpolation_utils.py:44-172 ->
 * _extensions/interputils_cython.pyx:42-182.
 * tx,ty,tz: the three 1-D sample arrays of the location grid (device, lengths
 * ntx,nty,ntz), exactly the float64 arrays the host built (bit-exact bisection).
 * pts: (npts,3) row-major.  cols[npts*8] (F-order flat index + col_offset; int32 entries when
 * cols_is64 == 0 -- allowed while every index fits, like scipy's index arrays -- else int64) and
 * w[npts*8] in the reference's entry order. */
int dzb_interp_locate(const double *tx, int64_t ntx, const double *ty, int64_t nty, const double *tz, int64_t ntz,
                      const double *pts, int64_t npts, int64_t col_offset, void *cols, int cols_is64, double *w,
                      void *stream);
/* y[npts] = Q v from stored cols/w;  row_scale (may be NULL) multiplies each row (0 for zeros_outside rows). */
int dzb_interp_apply(const void *cols, int cols_is64, const double *w, const double *row_scale, int64_t npts,
                     const double *v, double *y, void *stream);
/* y[ncols] += Q^T x (y must be zeroed by the caller). */
int dzb_interp_apply_t(const void *cols, int cols_is64, const double *w, const double *row_scale, int64_t npts,
                       const double *x, double *y, void *stream);
/* The same two applies on BUCKETED tables: the host side sorts the table rows once by the flat index of their first
 * grid sample (so that consecutive rows gather from neighbouring grid entries) and passes perm[p] = the point row p
 * of the table belongs to; row_scale is in table order.  perm == NULL: tables in point order (the calls above). */
int dzb_interp_apply_perm(const void *cols, int cols_is64, const double *w, const double *row_scale, const int32_t *perm,
                          int64_t npts, const double *v, double *y, void *stream);
int dzb_interp_apply_t_perm(const void *cols, int cols_is64, const double *w, const double *row_scale,
                            const int32_t *perm, int64_t npts, const double *x, double *y, void *stream);
/* dst[i] = src[idx[i]], i < n: returns a result computed in bucket order (dzb_interp_apply on bucketed tables with
 * perm == NULL) to point order with random READS, which cost half of what the scattered writes of *_perm cost. */
int dzb_gather_rows(int64_t n, const double *src, const int32_t *idx, double *dst, void *stream);
/* Compact tables (K10c / K11c): a row keeps the flat index of its first grid sample (`base`, int32 or int64) and the
 * three per-axis weights w1 (28 bytes instead of 96); a weight stored as -0.5 marks an axis where the reference clamps
 * (i2 == i1, both weights 0.5).  Columns base + {0,1} + {0,ntx} + {0,ntx*nty} and weights ((wx*wy)*wz), w2 = 1 - w1, are
 * rebuilt in the kernels with the reference's expressions (interputils_cython.pyx:60-67, 145-180): dzb_interp_expand
 * writes tables bit-identical to dzb_interp_locate's.  Rows may be in any order (the host side sorts them by `base`);
 * `perm` as in dzb_interp_apply_t_perm. */
int dzb_interp_locate_compact(const double *tx, int64_t ntx, const double *ty, int64_t nty, const double *tz, int64_t ntz,
                              const double *pts, int64_t npts, int64_t col_offset, void *base, int base_is64, double *wx,
                              double *wy, double *wz, void *stream);
int dzb_interp_apply_compact(const void *base, int base_is64, const double *wx, const double *wy, const double *wz,
                             int64_t ntx, int64_t nty, const double *row_scale, int64_t npts, const double *v, double *y,
                             void *stream);
int dzb_interp_apply_t_compact(const void *base, int base_is64, const double *wx, const double *wy, const double *wz,
                               int64_t ntx, int64_t nty, const double *row_scale, const int32_t *perm, int64_t npts,
                               const double *x, double *y, void *stream);
int dzb_interp_expand(const void *base, int base_is64, const double *wx, const double *wy, const double *wz, int64_t ntx,
                      int64_t nty, int64_t npts, void *cols, int cols_is64, double *w, void *stream);
/* Fused locate + apply: y = Q v without storing Q. */
int dzb_interp_fused(const double *tx, int64_t ntx, const double *ty, int64_t nty, const double *tz, int64_t ntz,
                     const double *pts, int64_t npts, int64_t col_offset, const double *row_scale, const double *v,
                     double *y, void *stream);
/* inside[npts] = 1 if the point is inside the mesh (is_inside, base_tensor_mesh.py:638-682);
 * lo/hi/tol: 3 doubles each on the HOST. */
int dzb_points_inside(const double *pts, int64_t npts, const double *lo_host, const double *hi_host,
                      const double *tol_host, uint8_t *inside, void *stream);

/* ---- TreeMesh (OcTree) operators ------------------------------------------------ */
/* y = A x for a deflated OcTree operator (face_divergence, edge_curl, nodal_gradient, average_*:
 * _extensions/tree_ext.pyx:3114-4762) stored as CSR connectivity + one code byte per entry.
 * value(entry) = tab_mult[code] * (tab_sel[code] == 0 ? 1 : g{tab_sel[code]-1}[row])
 * (geom_by_col != 0: the geometry is indexed by the column instead -- used for A^T).
 * indptr: int32 or int64 (indptr_is64); indices int32; tab_* are DEVICE arrays of ntab entries;
 * g0..g2 device arrays or NULL; lanes_per_row in {1,2,4,8,16,32}, or 0 for the warp-staged kernel (a warp copies
 * the contiguous entries of its 32 rows to shared memory with coalesced loads, then each lane walks one row). */
int dzb_tree_spmv(int64_t nrows, const void *indptr, int indptr_is64, const int32_t *indices, const uint8_t *codes,
                  const double *tab_mult, const uint8_t *tab_sel, int ntab, const double *g0, const double *g1,
                  const double *g2, int geom_by_col, const double *x, double *y, int lanes_per_row, void *stream);
/* The same operator in ELL layout: entry e of row r at indices[e*stride + r] / codes[e*stride + r], stride >= nrows,
 * row_len slots per row.  padded == 0: every row holds exactly row_len entries (on a balanced OcTree: every
 * cell-rowed operator); padded != 0: shorter rows are filled with slots of code 255 (edge_curl: 4..6 entries).
 * ntab <= 255. */
int dzb_tree_ell_spmv(int64_t nrows, int row_len, int64_t stride, const int32_t *indices, const uint8_t *codes,
                      const double *tab_mult, const uint8_t *tab_sel, int ntab, const double *g0, const double *g1,
                      const double *g2, int geom_by_col, int padded, const double *x, double *y, void *stream);
/* The same with the table rows (indices, codes, row geometry g*) stored in an order that follows the COLUMN numbering
 * (cells are numbered along the tree's Z-order, faces / edges / nodes by Cantor key: in cell order every gather of a
 * warp touches ~25 sectors, sorted by first column ~12): table row r is row perm[r] of the operator, where y is written. */
int dzb_tree_ell_spmv_perm(int64_t nrows, int row_len, int64_t stride, const int32_t *indices, const uint8_t *codes,
                           const double *tab_mult, const uint8_t *tab_sel, int ntab, const double *g0, const double *g1,
                           const double *g2, int geom_by_col, int padded, const int32_t *perm, const double *x, double *y,
                           void *stream);
/* Sliced ELL (SELL-32) for rows of different lengths (transposed operators, nodal_gradient): slice s = rows 32 s ..
 * 32 s + 31 stores entry e of its row l at slice_ptr[s] + 32 e + l with the slice's own width
 * (slice_ptr[s+1] - slice_ptr[s]) / 32; padding slots carry code 255.  slice_ptr: ceil(nrows / 32) + 1 offsets (int32 or
 * int64); perm as in dzb_tree_ell_spmv_perm (NULL: table row r is row r). */
int dzb_tree_sell_spmv(int64_t nrows, const void *slice_ptr, int slice_ptr_is64, const int32_t *indices,
                       const uint8_t *codes, const double *tab_mult, const uint8_t *tab_sel, int ntab, const double *g0,
                       const double *g1, const double *g2, int geom_by_col, const int32_t *perm, const double *x,
                       double *y, void *stream);
/* Hybrid layout for ragged rows: dzb_tree_ell_spmv(_perm) on the first K0 entries of every row (padded = 1), then this
 * on the few rows that have more: y[rows[r]] += the entries indptr[r] .. indptr[r+1] of overflow row r.  Same code table
 * and geometry arrays as the ELL part; same stream, after it. */
int dzb_tree_rows_add(int64_t nrem, const int32_t *rows, const void *indptr, int indptr_is64, const int32_t *indices,
                      const uint8_t *codes, const double *tab_mult, const uint8_t *tab_sel, int ntab, const double *g0,
                      const double *g1, const double *g2, int geom_by_col, const double *x, double *y, void *stream);

/* OcTree point location (TreeMesh.point2index / get_containing_cells; _extensions/tree.cpp:758-766, 1499-1516): cells[p] =
 * the leaf that contains point p (ties and outside points as in the reference: root boundaries go high, cell centres go
 * low, outside points end in a boundary cell).  nodes_*: fine-grid node coordinates per axis (n + 1 entries, device);
 * roots: root cells per axis (3 host ints); root_width: fine cells per root (power of two); leaf_keys: the sorted keys
 * root_id * root_width^3 + Z-order code of every leaf's low corner, in cell order (device); pts: (npts, 3) row-major. */
int dzb_tree_locate(const double *nodes_x, const double *nodes_y, const double *nodes_z, const int32_t *roots, int root_width,
                    const int64_t *leaf_keys, int64_t n_leaves, const double *pts, int64_t npts, int64_t *cells,
                    void *stream);
/* y = A x for a CSR matrix with stored float64 values (int32 columns, int32 / int64 indptr): scipy
 * matrices mixed into operator algebra and the full-tensor OcTree mass matrices
 * (operators/inner_products.py:147-272 with tree_ext.pyx:5651-5763 projections). */
int dzb_csr_spmv(int64_t nrows, const void *indptr, int indptr_is64, const int32_t *indices, const double *values,
                 const double *x, double *y, int lanes_per_row, void *stream);
/* y = d .* x (invert == 0) or x ./ d: diagonal mass matrices on a TreeMesh
 * (base/base_tensor_mesh.py:792-863 with the tree averages). */
int dzb_diag_scale(int64_t n, const double *d, const double *x, double *y, int invert, void *stream);
/* Full-tensor mass matrices on an OcTree, matrix-free: M x = R^T A (R x) with R the deflation matrix (applied by
 * dzb_csr_spmv around this call) and A = sum over the 8 cell corners of P_n^T (V/8) Sigma P_n on TOTAL entities
 * (operators/inner_products.py:147-272, corner tables of _extensions/tree_ext.pyx:5651-5763).
 * ent: int32 [slots][n_cells] (slot-major) total entity numbers of every cell, slots = fx0 fx1 fy0 fy1 fz0 fz1 (edges == 0)
 * or ex(b+2c) x4, ey(a+2c) x4, ez(a+2b) x4 (edges != 0); model6: [6][n_cells] = xx yy zz xy xz yz;
 * y_total must be zeroed by the caller (the kernel accumulates with atomics). */
int dzb_tree_mass_tensor(int64_t n_cells, int edges, const int32_t *ent, const double *model6, const double *cell_volumes,
                         const double *x_total, double *y_total, void *stream);
/* Adjoint of the model derivative of the above: out6[q][C] = (V_C/8) sum_n (u_n(v) (x) u_n(w))_q for the 6 tensor slots
 * (operators/inner_products.py:459-565); the forward derivative F(v) dm is dzb_tree_mass_tensor with dm as the model. */
int dzb_tree_mass_tensor_adj(int64_t n_cells, int edges, const int32_t *ent, const double *cell_volumes,
                             const double *v_total, const double *w_total, double *out6, void *stream);

/* ---- solver-side vector kernels ------------------------------------ */
/* One Jacobi-preconditioned CG iteration on an operator A is: Ap = A p (any apply above), dzb_dot(p, Ap -> scal[1]),
 * dzb_cg_update, dzb_cg_direction.  All scalars live in the DEVICE array scal[6]:
 * scal[0] = r.z (current), scal[1] = p.Ap, scal[2] = r.z (new), scal[3] = r.r.  `workspace`: device memory of
 * dzb_solver_workspace_bytes() bytes, zeroed once by the caller.  Reductions are deterministic. */
int64_t dzb_solver_workspace_bytes(void);
/* scal[slot] = a . b */
int dzb_dot(int64_t n, const double *a, const double *b, void *workspace, double *scal, int slot, void *stream);
/* alpha = scal[0]/scal[1]; x += alpha p; r -= alpha Ap; z = dinv ? dinv*r : r (z may alias r when dinv == NULL);
 * scal[2] = r.z, scal[3] = r.r */
int dzb_cg_update(int64_t n, double *x, double *r, const double *p, const double *Ap, const double *dinv, double *z,
                  void *workspace, double *scal, void *stream);
/* beta = scal[2]/scal[0]; p = z + beta p; scal[0] = scal[2] */
int dzb_cg_direction(int64_t n, double *p, const double *z, void *workspace, double *scal, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DZB_H */
