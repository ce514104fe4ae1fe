/*
 * symtensor_b200 -- C-ABI of the B200-native backend for symtensor's symmetrized-contraction hot path.
 *
 * The reference (Eike-Flath/symtensor) is pure Python and has no FFI: its hot path is reached through
 * class-level registries (`SymmetricTensor.implements` / `implements_ufunc.outer`,
 * symtensor/base.py:259-322, 1057-1063).  The functions below are what a backend mixin registered there
 * binds (INTEGRATION.md shows the ctypes stub); each entry cites the reference code it replaces, paths
 * relative to the reference repository.
 *
 * Conventions
 *  - plain C types only; every function returns an `st_status` (0 = ok) and never throws;
 *    `st_last_error()` returns a thread-local message for the last non-zero status;
 *  - pointers named `d_*` are DEVICE pointers owned by the caller (e.g. torch CUDA tensors), `h_*` are host
 *    pointers; `stream` is a `cudaStream_t` passed as `void*` (NULL = default stream); all device work is
 *    stream-ordered and the call returns without synchronising unless stated;
 *  - the library keeps no per-call state; it caches one immutable "plan" (class table, binomial table) per
 *    (device, rank, dim), mirroring the reference's lazily built `pos_dict[(rank, dim)]`
 *    (symtensor/permcls_symtensor.py:422-445);
 *  - `rank` <= ST_MAX_RANK.  Positions / sizes are int64.
 *
 * Packed layouts
 *  ST_LAYOUT_PERMCLS  one contiguous buffer holding the reference's `_data` dict
 *                     (symtensor/permcls_symtensor.py:546-547): classes in `_perm_classes(rank)` order
 *                     (symtensor/utils.py:839-856, 1000-1002); inside a class the `σindex_iter` order
 *                     (symtensor/permcls_symtensor.py:288-347).  Class c starts at `offsets[c]`
 *                     (multiple of ST_CLASS_ALIGN elements, zero padded) as returned by st_class_table.
 *  ST_LAYOUT_FLAT     the single 1-D array of `FlatSymmetricTensor`, in
 *                     `itertools.combinations_with_replacement(range(dim), rank)` order
 *                     (symtensor/flat_symtensor.py:39-50, 219-220).
 */
#ifndef SYMTENSOR_B200_H
#define SYMTENSOR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ST_MAX_RANK 16
#define ST_CLASS_ALIGN 32 /* elements; class starts are padded to this (256 B fp64 / 128 B fp32) */

#define ST_LAYOUT_PERMCLS 0
#define ST_LAYOUT_FLAT 1

/* combiner of the symmetrized outer ops (st_outer_op_*): symalg.multiply / add / subtract (symtensor/symalg.py:193-195) */
#define ST_OUTER_MULTIPLY 0
#define ST_OUTER_ADD 1
#define ST_OUTER_SUBTRACT 2

typedef enum {
  ST_OK = 0,
  ST_ERR_INVALID = 1,     /* bad argument (maps to ValueError in the Python mixin) */
  ST_ERR_CUDA = 2,        /* a CUDA runtime call failed (RuntimeError) */
  ST_ERR_UNSUPPORTED = 3, /* valid request this build cannot serve (NotImplementedError) */
  ST_ERR_OVERFLOW = 4     /* a size does not fit int64 (OverflowError) */
} st_status;

int st_version(void);
const char* st_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * Class tables (host side, no GPU needed).
 * Replaces utils._perm_classes / _all_index_counts (symtensor/utils.py:839-856, 1000-1002),
 * utils._get_permclass_size (:925-933), utils.get_permclass_multiplicity / multinom (:207-223, 760-776)
 * and SymmetricTensor.indep_size (symtensor/base.py:833-844).  Exact integer arithmetic.
 * ------------------------------------------------------------------------------------------------ */
/* number of permutation classes p(rank); negative on error */
int st_num_classes(int rank);
/* parts   [ncls * rank]  class c = parts[c*rank .. c*rank+nparts[c]) (descending), rest zero
 * nparts  [ncls]         number of distinct index values l of class c
 * sizes   [ncls]         stored components s_c (0 when l > dim)
 * mults   [ncls]         multiplicity gamma_c = rank!/prod(m_k!)
 * offsets [ncls + 1]     start of class c in the ST_LAYOUT_PERMCLS buffer; offsets[ncls] = padded total
 * Any output pointer may be NULL. */
int st_class_table(int rank, int64_t dim, int32_t* parts, int32_t* nparts, int64_t* sizes, int64_t* mults,
                   int64_t* offsets);
/* C(dim + rank - 1, rank) */
int st_indep_size(int rank, int64_t dim, int64_t* out);

/* Host rank / unrank of single indices (used by the mixin's __getitem__/__setitem__, replaces the
 * position registry PosRegistry/_convert_dense_index, symtensor/permcls_symtensor.py:422-479, and
 * get_index_representative :375-381, utils._get_permclass symtensor/utils.py:878-889).
 * idx: `rank` values in [0, dim), any order.  cls: class ordinal in st_class_table order. */
int st_host_permcls_rank(int rank, int64_t dim, const int32_t* idx, int32_t* cls, int64_t* pos);
/* writes the representative multi-index (rank values) of component `pos` of class `cls` */
int st_host_permcls_unrank(int rank, int64_t dim, int32_t cls, int64_t pos, int32_t* idx);
/* flat_symtensor.index_of_multicombination (symtensor/flat_symtensor.py:39-50); idx any order */
int st_host_flat_rank(int rank, int64_t dim, const int32_t* idx, int64_t* pos);
int st_host_flat_unrank(int rank, int64_t dim, int64_t pos, int32_t* idx);

/* ------------------------------------------------------------------------------------------------
 * GPU index-class enumerator (bulk).  Replaces the Python generators σindex_iter / _sub_σindex_iter
 * (symtensor/permcls_symtensor.py:288-347), indep_iter_repindex (:958-960), the PosRegistry lookups
 * (:422-479) and flat indep_iter_repindex / index_of_multicombination (symtensor/flat_symtensor.py:39-50,
 * 219-220).  Bit-exact with the reference's storage order.
 * ------------------------------------------------------------------------------------------------ */
/* d_idx_out[(p - begin) * rank + k] = k-th entry of the representative multi-index of component p of
 * class `cls`, for p in [begin, begin + count). */
int st_permcls_unrank(int rank, int64_t dim, int32_t cls, int64_t begin, int64_t count, int32_t* d_idx_out,
                      void* stream);
/* n arbitrary multi-indices d_idx[n * rank] -> class ordinal and position inside the class */
int st_permcls_rank(int rank, int64_t dim, int64_t n, const int32_t* d_idx, int32_t* d_cls_out,
                    int64_t* d_pos_out, void* stream);
int st_flat_unrank(int rank, int64_t dim, int64_t begin, int64_t count, int32_t* d_idx_out, void* stream);
int st_flat_rank(int rank, int64_t dim, int64_t n, const int32_t* d_idx, int64_t* d_pos_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * contract_all_indices_with_vector  (symtensor/symalg.py:505-527):
 *     s = sum_{i1..ir} A[i1..ir] x[i1]...x[ir]
 *       = sum_classes gamma_c * sum_p A_c[p] * prod_j x[v_j(p)]^{m_j}
 * over the packed coordinates [begin, end) of the buffer (whole tensor: begin = 0, end = total).  Ranges
 * let callers shard the tensor over GPUs or stream it from the host; `d_packed` points at coordinate
 * `begin` (i.e. the local shard), begin must be a multiple of ST_CLASS_ALIGN.
 * d_out receives ONE value (the partial sum of the range).  The result is deterministic: every tile of the
 * range has its own partial-sum slot in the workspace and the slots are added in index order by the last CTA
 * of the same launch (permcls layout; the flat layout uses a second single-CTA pass).
 * d_workspace: st_contract_vec_workspace_bytes() bytes (4 MiB) of device scratch, no initialisation needed;
 *              launches that may overlap (different streams) need different workspaces.
 * The caller handles the reference's early exits (len(x) != dim -> ValueError, all-zero x -> 0).
 * ------------------------------------------------------------------------------------------------ */
int64_t st_contract_vec_workspace_bytes(void);
int st_contract_vec_f64(int layout, int rank, int64_t dim, const double* d_packed, int64_t begin, int64_t end,
                        const double* d_x, double* d_out, void* d_workspace, void* stream);
int st_contract_vec_f32(int layout, int rank, int64_t dim, const float* d_packed, int64_t begin, int64_t end,
                        const float* d_x, float* d_out, void* d_workspace, void* stream);
/* The same launch with `flags`.  ST_VEC_OVERLAP: the call is one of a BATCH of contractions of operands that were complete
 * before the previous operation on `stream` was enqueued (e.g. one resident tensor contracted with many vectors, or the
 * steps of a loop over resident tensors): the kernel is then launched as a programmatic dependent of the previous launch
 * on the stream and starts streaming while that one -- another vector contraction -- drains its tail (ramp-up and tail
 * of consecutive launches overlap; at most two launches are in flight).  With the flag d_workspace holds
 * 2 * st_contract_vec_workspace_bytes() bytes (the library alternates between the halves), and d_out is written in
 * stream order as always.  Without the promise the flag must not be set (the kernel reads d_x / d_packed early). */
#define ST_VEC_OVERLAP 1
int st_contract_vec_ex_f64(int layout, int rank, int64_t dim, const double* d_packed, int64_t begin, int64_t end,
                           const double* d_x, double* d_out, void* d_workspace, int flags, void* stream);
int st_contract_vec_ex_f32(int layout, int rank, int64_t dim, const float* d_packed, int64_t begin, int64_t end,
                           const float* d_x, float* d_out, void* d_workspace, int flags, void* stream);
/* Same op with HOST buffers: streams the packed range through pinned staging buffers in chunks, overlapping
 * the host->device copies with the kernel, and returns the value on the host (synchronises). */
int st_contract_vec_host_f64(int layout, int rank, int64_t dim, const double* h_packed, int64_t total,
                             const double* h_x, double* h_out);
int st_contract_vec_host_f32(int layout, int rank, int64_t dim, const float* h_packed, int64_t total,
                             const float* h_x, float* h_out);

/* ------------------------------------------------------------------------------------------------
 * Layout converters: permcls <-> flat re-ordering of the packed components (the step either side of the
 * ops; replaces the Python gathers of PermClsSymmetricTensor._validate_data / todense,
 * symtensor/permcls_symtensor.py:599-618, 883-887, and FlatSymmetricTensor.__init__,
 * symtensor/flat_symtensor.py:100-110).  d_flat has C(dim+rank-1, rank) elements; d_permcls is the
 * ST_LAYOUT_PERMCLS buffer (padding is written as zeros); [begin, end) is a range of permcls coordinates
 * and d_permcls points at coordinate `begin`.
 * ------------------------------------------------------------------------------------------------ */
int st_permcls_to_flat_f64(int rank, int64_t dim, const double* d_permcls, double* d_flat, void* stream);
int st_permcls_to_flat_f32(int rank, int64_t dim, const float* d_permcls, float* d_flat, void* stream);
int st_flat_to_permcls_f64(int rank, int64_t dim, const double* d_flat, double* d_permcls, int64_t begin, int64_t end, void* stream);
int st_flat_to_permcls_f32(int rank, int64_t dim, const float* d_flat, float* d_permcls, int64_t begin, int64_t end, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Dense <-> packed (the constructor / todense step either side of the ops).  d_dense is the row-major
 * dim^rank array.  st_pack_dense_* fills the packed coordinates [begin, end) (d_packed points at `begin`;
 * permcls padding is written as zeros): symmetrize = 0 takes the representative entry of every component
 * and checks the symmetry of the dense array with numpy.allclose(rtol, atol) semantics over all
 * permutations, setting *d_flag (device int, caller-zeroed) to 1 on a violation -- the reference raises
 * ValueError("Data array is not symmetric.") then (PermClsSymmetricTensor._validate_data,
 * symtensor/permcls_symtensor.py:599-618; utils.is_symmetric, symtensor/utils.py:563-578); symmetrize = 1
 * stores the mean over all axis permutations (utils.symmetrize, symtensor/utils.py:507-532).
 * st_unpack_dense_* is todense (symtensor/permcls_symtensor.py:883-887, torch_symtensor.py:564-568,
 * flat_symtensor.py:251-256).
 * ------------------------------------------------------------------------------------------------ */
int st_pack_dense_f64(int layout, int rank, int64_t dim, const double* d_dense, double* d_packed, int64_t begin, int64_t end,
                      int symmetrize, double rtol, double atol, int* d_flag, void* stream);
int st_pack_dense_f32(int layout, int rank, int64_t dim, const float* d_dense, float* d_packed, int64_t begin, int64_t end,
                      int symmetrize, double rtol, double atol, int* d_flag, void* stream);
int st_unpack_dense_f64(int layout, int rank, int64_t dim, const double* d_packed, double* d_dense, void* stream);
int st_unpack_dense_f32(int layout, int rank, int64_t dim, const float* d_packed, float* d_dense, void* stream);

/* ------------------------------------------------------------------------------------------------
 * multiply.outer, symmetrized  (symtensor/symalg.py:294-316 via symmetrized_op :206-283):
 *     C_K = C(ra+rb, ra)^-1 * sum over position subsets S, |S| = ra, of A[K_S] * B[K_S^c]
 * Operands in ST_LAYOUT_FLAT (rank ra / rb, same dim); the output is the ST_LAYOUT_PERMCLS buffer of rank
 * ra + rb, coordinates [begin, end) (d_out points at `begin`): output ranges shard over GPUs with no
 * collective.  st_outer_vec_* is the fused outer -> contract_all_indices_with_vector: it returns
 * sum_K gamma_K C_K prod x[K] over the range without ever storing the rank-(ra+rb) tensor.
 * ------------------------------------------------------------------------------------------------ */
int st_outer_f64(int ra, int rb, int64_t dim, const double* d_a_flat, const double* d_b_flat, double* d_out, int64_t begin,
                 int64_t end, void* stream);
int st_outer_f32(int ra, int rb, int64_t dim, const float* d_a_flat, const float* d_b_flat, float* d_out, int64_t begin,
                 int64_t end, void* stream);
/* add.outer / subtract.outer / multiply.outer with an explicit combiner `op` (ST_OUTER_*): the reference registers the
 * same symmetrized outer for the three ufunc wrappers (symtensor/symalg.py:294-316):
 *     C_K = C(ra+rb, ra)^-1 * sum over position subsets S of  A[K_S] (op) B[K_S^c]                                   */
int st_outer_op_f64(int op, int ra, int rb, int64_t dim, const double* d_a_flat, const double* d_b_flat, double* d_out,
                    int64_t begin, int64_t end, void* stream);
int st_outer_op_f32(int op, int ra, int rb, int64_t dim, const float* d_a_flat, const float* d_b_flat, float* d_out,
                    int64_t begin, int64_t end, void* stream);
int64_t st_outer_vec_workspace_bytes(void);
int st_outer_vec_f64(int ra, int rb, int64_t dim, const double* d_a_flat, const double* d_b_flat, const double* d_x,
                     double* d_out, void* d_workspace, int64_t begin, int64_t end, void* stream);
int st_outer_vec_f32(int ra, int rb, int64_t dim, const float* d_a_flat, const float* d_b_flat, const float* d_x, float* d_out,
                     void* d_workspace, int64_t begin, int64_t end, void* stream);

/* ------------------------------------------------------------------------------------------------
 * tensordot, symmetrized, k contracted index pairs  (symtensor/symalg.py:427-459):
 *     C_K = C(n, ra-k)^-1 sum_S sum_{J in [d]^k} A[K_S, J] B[J, K_S^c],   n = ra + rb - 2k
 * computed as a pair-packed Gram matrix G = Aexp . Bexp^T (rows: packed free indices, columns: packed
 * contracted tuples weighted by their multiplicity) followed by the gather of the C(n, ra-k) splits.
 * Operands ST_LAYOUT_FLAT; output ST_LAYOUT_PERMCLS of rank n over [begin, end) (rank 0: one component,
 * the reference returns dim 1).  d_workspace: st_tensordot_workspace_bytes() bytes.
 * ------------------------------------------------------------------------------------------------ */
int st_tensordot_workspace_bytes(int ra, int rb, int k, int64_t dim, int elem_size, int64_t* out_bytes);
/* 1 when st_tensordot_* serves (ra, rb, k, dim, elem_size) with the tiled NON-MATERIALISING kernel (fp32, two free indices on
 * each side: BASELINE config 3): the six Gram terms of a 16 x 16 x 16 x 8 output tile are three 128 x 256 tcgen05 GEMMs over
 * TMA-staged operand boxes, added straight into the packed output (which is zeroed first); the workspace then holds the
 * expanded, hi/lo-split pair matrices (16 dim^2 Kp bytes) and its first int is an error flag the kernel raises if a
 * bounded barrier wait expired (read it after synchronising the stream). */
int st_tensordot_is_tiled(int ra, int rb, int k, int64_t dim, int elem_size);
int st_tensordot_f64(int ra, int rb, int k, int64_t dim, const double* d_a_flat, const double* d_b_flat, double* d_out,
                     int64_t begin, int64_t end, void* d_workspace, void* stream);
int st_tensordot_f32(int ra, int rb, int k, int64_t dim, const float* d_a_flat, const float* d_b_flat, float* d_out,
                     int64_t begin, int64_t end, void* d_workspace, void* stream);

/* Several DISJOINT output ranges in one call (fp32): range q of the packed output goes to d_outs[q] (which starts at coordinate
 * begins[q]); at most 8 ranges.  For the tiled kernel a tile that serves several ranges runs once -- the multi-GPU shard of
 * sharding.tensordot22_shards: a GPU's part of every class with repeated indices plus its part of class (1,1,1,1), so that the
 * tiles that hold a diagonal are spread over the GPUs instead of all landing on the one that owns the start of the buffer.
 * begins / ends / d_outs are HOST arrays. */
int st_tensordot_ranges_f32(int ra, int rb, int k, int64_t dim, const float* d_a_flat, const float* d_b_flat, int nranges,
                            const int64_t* begins, const int64_t* ends, float* const* d_outs, void* d_workspace, void* stream);

/* ------------------------------------------------------------------------------------------------
 * contract_all_indices_with_matrix  (symtensor/symalg.py:475-496):
 *     C[j1..jr] = sum A[i1..ir] W[i1,j1] ... W[ir,jr]      (W is dim x dim, row-major, contracted on axis 0)
 * as the partially-symmetric mode chain T_{k+1}[j1..j_{k+1}; I] = sum_a W[a, j_{k+1}] T_k[j1..jk; sort(a, I)];
 * every intermediate is stored packed x packed.  Input and output ST_LAYOUT_FLAT of the same rank / dim.
 * d_workspace: st_contract_mat_workspace_bytes() bytes (two ping-pong intermediates).
 * ------------------------------------------------------------------------------------------------ */
int st_contract_mat_workspace_bytes(int rank, int64_t dim, int elem_size, int64_t* out_bytes);
int st_contract_mat_f64(int rank, int64_t dim, const double* d_a_flat, const double* d_W, double* d_out_flat, void* d_workspace,
                        void* stream);
int st_contract_mat_f32(int rank, int64_t dim, const float* d_a_flat, const float* d_W, float* d_out_flat, void* d_workspace,
                        void* stream);

/* The same contraction restricted to the output components whose FIRST (smallest) mode j1 lies in [jlo, jhi) -- the multi-GPU
 * partition of the matrix contraction "by output permutation classes / first output mode" (SURVEY.md 8e): with the flat
 * (lexicographic) order these components are the contiguous range [*flat_begin, *flat_end) of the output
 * (st_contract_mat_range_bounds), d_out_slice starts at *flat_begin, and every intermediate of the mode chain shards with
 * it, so a GPU holds and computes only its slice of the chain (workspace: st_contract_mat_range_workspace_bytes).  A and W are
 * replicated; no collective -- the result stays sharded (concatenate the slices for the whole tensor). */
int st_contract_mat_range_bounds(int rank, int64_t dim, int64_t jlo, int64_t jhi, int64_t* flat_begin, int64_t* flat_end);
int st_contract_mat_range_workspace_bytes(int rank, int64_t dim, int64_t jlo, int64_t jhi, int elem_size, int64_t* out_bytes);
int st_contract_mat_range_f64(int rank, int64_t dim, const double* d_a_flat, const double* d_W, double* d_out_slice, int64_t jlo, int64_t jhi,
                              void* d_workspace, void* stream);
int st_contract_mat_range_f32(int rank, int64_t dim, const float* d_a_flat, const float* d_W, float* d_out_slice, int64_t jlo, int64_t jhi,
                              void* d_workspace, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Rows next to the path, so that whole expressions stay on the GPU.  `layout` / `rank` / `dim` describe
 * the packed buffers (same layout for operands and result; permcls padding is written as zeros).
 * Elementwise ufuncs on the packed components (SymmetricTensor.default_unary_ufunc / default_binary_ufunc,
 * symtensor/base.py:1146-1362).  Binary `mode`: 0 = a (op) b, 1 = a (op) scalar, 2 = scalar (op) a.
 * ------------------------------------------------------------------------------------------------ */
#define ST_UN_NEGATIVE 0
#define ST_UN_ABS 1
#define ST_UN_SQRT 2
#define ST_UN_SQUARE 3
#define ST_UN_EXP 4
#define ST_UN_LOG 5
#define ST_UN_RECIPROCAL 6
#define ST_BIN_ADD 0
#define ST_BIN_SUBTRACT 1
#define ST_BIN_MULTIPLY 2
#define ST_BIN_DIVIDE 3
#define ST_BIN_MAXIMUM 4
#define ST_BIN_MINIMUM 5
#define ST_BIN_POWER 6
int st_elementwise_unary_f64(int op, int layout, int rank, int64_t dim, const double* d_a, double* d_out, void* stream);
int st_elementwise_unary_f32(int op, int layout, int rank, int64_t dim, const float* d_a, float* d_out, void* stream);
int st_elementwise_binary_f64(int op, int mode, int layout, int rank, int64_t dim, const double* d_a, const double* d_b, double scalar, double* d_out,
                              void* stream);
int st_elementwise_binary_f32(int op, int mode, int layout, int rank, int64_t dim, const float* d_a, const float* d_b, double scalar, float* d_out,
                              void* stream);
/* isclose / allclose / array_equal (symtensor/base.py:1521-1684) of two tensors of the same layout, or of a tensor and a scalar
 * (b_is_scalar).  mode 0: a == b, mode 1: numpy.isclose(a, b, rtol, atol, equal_nan).  *d_all (device int, caller-set to 1) is
 * cleared when a component fails (allclose / array_equal); d_mask (optional) receives 1 / 0 per component (isclose). */
int st_compare_f64(int mode, int layout, int rank, int64_t dim, const double* d_a, const double* d_b, int b_is_scalar, double scalar, double rtol,
                   double atol, int equal_nan, double* d_mask, int* d_all, void* stream);
int st_compare_f32(int mode, int layout, int rank, int64_t dim, const float* d_a, const float* d_b, int b_is_scalar, double scalar, double rtol, double atol,
                   int equal_nan, float* d_mask, int* d_all, void* stream);
/* Partial indexing A[i_1, ..., i_n] (symtensor/permcls_symtensor.py:750-781): the rank-(rank - nfixed) tensor B[K] = A[K + fixed],
 * one gather per packed coordinate of B.  d_fixed: `nfixed` device int32 indices. */
int st_slice_f64(int layout, int rank, int64_t dim, int nfixed, const int32_t* d_fixed, const double* d_a, double* d_out, void* stream);
int st_slice_f32(int layout, int rank, int64_t dim, int nfixed, const int32_t* d_fixed, const float* d_a, float* d_out, void* stream);

/* kernel variant selection for benchmarking / tests: 0 = auto (small classes per component, the rest through the
 * ring kernel), 1 = generic per-element enumerator, 2 = every class through the ring kernel.  Process-wide. */
int st_set_vec_variant(int variant);
/* tuning hooks (process-wide; defaults are the tuned values): "vec_force_tau" in [0, ST_MAX_RANK] (0 = cost model),
 * "vec_small_class", "vec_use_dir", and for the ring kernel "vec_ring_warps", "vec_ring_slots", "vec_ring_bytes",
 * "vec_ring_bytes_max", "vec_ring_tile_bytes", "vec_ring_table_max", "vec_ring_direct", "vec_ring_dynamic"
 * (0: static tile deal), "vec_timeline" (debug stamps, see st_debug_vec_timeline), "vec_short_launch_bytes" /
 * "vec_short_launch_slots" (launches over a slice of at most that many bytes use tiles of that many ring slots).
 * Kernel selection (test hooks; 1 = the round-2 kernel): "outer_fast", "outer_rows" (row-walk multiply.outer), "conv_rows" (row-walk
 * layout converters), "mat_dmma", "mat_pipe" (persistent producer / consumer step kernel), "mat_onfly_rows" (steps with at most
 * that many rows rank their gathers in the producers), "gram_umma", "sym22", "sym22_min_dim", "sym22_kch", "sym22_debug"
 * (ablation mask), "sym22_rgroup" (tile order), "sym22_batch_tiles" (tiles per launch, at most 32768). */
int st_set_tuning(const char* key, int64_t value);
/* debug: with st_set_tuning("vec_timeline", 1) the vector-contraction kernel stamps %globaltimer at its phase
 * boundaries ([cta][16] stamps, then one finish stamp per warp of the grid); this copies the first n stamps of
 * the last launch to the host.  Profiling aid (tools/vec_timeline.py), not part of the reference-facing surface. */
int st_debug_vec_timeline(unsigned long long* h_out, int64_t n);
/* debug (host only): `count` consecutive representative multi-indices of class `cls` from position `pos` -- the first by a full
 * unrank, the rest by the odometer step the outer kernel uses; returns the number written (rank ints each). */
int64_t st_debug_permcls_successors(int rank, int64_t dim, int32_t cls, int64_t pos, int64_t count, int32_t* h_idx);
/* debug (host only): the sorted multi-index of every coordinate of [begin, end) of the permcls layout (rank ints each, -1 for
 * alignment padding) as the multiply.outer kernel's row walk hands them to its lanes (spans of `span` coordinates, batches of
 * 32); dim <= 255; returns end - begin. */
int64_t st_debug_rowwalk(int rank, int64_t dim, int64_t begin, int64_t end, int64_t span, int32_t* h_idx);
/* ... and the same walk run on the DEVICE by the code the kernels use (d_idx: device buffer; span: a multiple of 32) */
int st_debug_rowwalk_device(int rank, int64_t dim, int64_t begin, int64_t end, int64_t span, int32_t* d_idx, void* stream);
/* ... with the latch of every coordinate (8 ints each: packed values lo / hi, b, m, offset, class, state, cursor values) in d_dbg */
int st_debug_rowwalk_device2(int rank, int64_t dim, int64_t begin, int64_t end, int64_t span, int32_t* d_idx, int32_t* d_dbg, void* stream);
/* debug (host only): the tile words p | q << 16 | r << 32 | s << 48 (index blocks 16 / 16 / 16 / 8) the tiled tensordot kernel
 * runs for the output range [begin, end) of a rank-4 result; returns their number (writes at most `cap`). */
int64_t st_debug_sym22_tiles(int64_t dim, int64_t begin, int64_t end, unsigned long long* h_out, int64_t cap);
/* ... for the union of several ranges (host arrays), as st_tensordot_ranges_f32 runs them */
int64_t st_debug_sym22_tiles_ranges(int64_t dim, int nranges, const int64_t* begins, const int64_t* ends, unsigned long long* h_out, int64_t cap);
/* ... and the terms of its cost model: stats[0] = number of tiles, stats[1] = sum over the tiles of 1 / (l blocks of the tile's k
 * block) */
int st_debug_sym22_tiles_stats(int64_t dim, int nranges, const int64_t* begins, const int64_t* ends, double* stats);
/* number of kernel launches issued by this library since load (bench.py reports it) */
int64_t st_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SYMTENSOR_B200_H */
