// The rows next to the hot path on the packed buffers, so that whole expressions stay on the GPU:
//  * elementwise ufuncs      SymmetricTensor.default_unary_ufunc / default_binary_ufunc   symtensor/base.py:1146-1362
//  * comparisons             isclose / allclose / array_equal                              symtensor/base.py:1521-1684
//  * partial indexing A[i..] the rank-lowering gather                                      symtensor/permcls_symtensor.py:750-781
// All three are one-pass, HBM-bound kernels over the packed coordinates (algorithmic bytes: the operands read once, the
// result written once); the alignment padding of the permcls layout is kept at zero by every kernel that writes a tensor.
#include <algorithm>

#include "st_common.cuh"

namespace st {

template <typename T>
__device__ __forceinline__ T apply_unary(int op, T a) {
  switch (op) {
    case ST_UN_NEGATIVE: return -a;
    case ST_UN_ABS: return a < T(0) ? -a : a;
    case ST_UN_SQRT: return (T)sqrt((double)a);
    case ST_UN_SQUARE: return a * a;
    case ST_UN_EXP: return (T)exp((double)a);
    case ST_UN_LOG: return (T)log((double)a);
    case ST_UN_RECIPROCAL: return T(1) / a;
    default: return a;
  }
}

template <typename T>
__device__ __forceinline__ T apply_binary(int op, T a, T b) {
  switch (op) {
    case ST_BIN_ADD: return a + b;
    case ST_BIN_SUBTRACT: return a - b;
    case ST_BIN_MULTIPLY: return a * b;
    case ST_BIN_DIVIDE: return a / b;
    case ST_BIN_MAXIMUM: return (a != a || b != b) ? (a != a ? a : b) : (a > b ? a : b);  // NaN propagates like np.maximum
    case ST_BIN_MINIMUM: return (a != a || b != b) ? (a != a ? a : b) : (a < b ? a : b);
    case ST_BIN_POWER: return (T)pow((double)a, (double)b);
    default: return a;
  }
}

// true for the alignment padding of the permcls layout (flat layout: never)
__device__ __forceinline__ bool is_padding(const PlanView& P, int layout, int64_t c) {
  if (layout != ST_LAYOUT_PERMCLS) return false;
  const int ci = class_of_coord(P, c);
  return c - P.cls[ci].offset >= P.cls[ci].size;
}

template <typename T>
__global__ void __launch_bounds__(256) unary_kernel(PlanView P, int layout, int op, int64_t n, const T* __restrict__ a, T* __restrict__ out) {
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < n; c += (int64_t)gridDim.x * blockDim.x)
    out[c] = is_padding(P, layout, c) ? T(0) : apply_unary<T>(op, a[c]);
}

// mode 0: a (op) b, 1: a (op) s, 2: s (op) a
template <typename T>
__global__ void __launch_bounds__(256) binary_kernel(PlanView P, int layout, int op, int mode, int64_t n, const T* __restrict__ a, const T* __restrict__ b,
                                                     T s, T* __restrict__ out) {
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < n; c += (int64_t)gridDim.x * blockDim.x) {
    T r = T(0);
    if (!is_padding(P, layout, c)) r = mode == 0 ? apply_binary<T>(op, a[c], b[c]) : mode == 1 ? apply_binary<T>(op, a[c], s) : apply_binary<T>(op, s, a[c]);
    out[c] = r;
  }
}

// numpy.isclose(a, b, rtol, atol, equal_nan) (mode 1) or a == b (mode 0) per component; *all_flag (caller-set to 1) is cleared
// when a component fails; mask (optional) receives 1 / 0 per component (padding: 0).  b_scalar: compare with the scalar s.
template <typename T>
__global__ void __launch_bounds__(256) compare_kernel(PlanView P, int layout, int mode, int64_t n, const T* __restrict__ a, const T* __restrict__ b, int b_scalar,
                                                      double s, double rtol, double atol, int equal_nan, T* __restrict__ mask, int* __restrict__ all_flag) {
  bool bad = false;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < n; c += (int64_t)gridDim.x * blockDim.x) {
    if (is_padding(P, layout, c)) { if (mask) mask[c] = T(0); continue; }
    const double x = (double)a[c], y = b_scalar ? s : (double)b[c];
    bool ok;
    if (mode == 0) ok = x == y;
    else if (x != x || y != y) ok = equal_nan && x != x && y != y;
    else if (isinf(x) || isinf(y)) ok = x == y;
    else ok = fabs(x - y) <= atol + rtol * fabs(y);
    if (mask) mask[c] = ok ? T(1) : T(0);
    bad = bad || !ok;
  }
  if (bad && all_flag) atomicAnd(all_flag, 0);
}

// partial indexing: B[K] = A[K + fixed] for every packed coordinate of the rank-(r - nfixed) tensor B (same layout as A)
template <typename T>
__global__ void __launch_bounds__(256) slice_kernel(PlanView PA, PlanView PB, int layout, int nfixed, const int32_t* __restrict__ fixed, const T* __restrict__ a,
                                                    T* __restrict__ out, int64_t n) {
  const int rb = PB.rank;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < n; c += (int64_t)gridDim.x * blockDim.x) {
    int32_t K[ST_MAX_RANK];
    if (layout == ST_LAYOUT_PERMCLS) {
      if (!permcls_coord_sorted(PB, c, K)) { out[c] = T(0); continue; }
    } else {
      flat_unrank_sorted(PB, c, K);
    }
    for (int q = 0; q < nfixed; ++q) K[rb + q] = fixed[q];
    T v;
    if (layout == ST_LAYOUT_PERMCLS) {
      int32_t vals[ST_MAX_RANK];
      const int ci = classify_index(PA, K, vals);
      v = a[PA.cls[ci].offset + permcls_rank_vals(PA, PA.cls[ci], vals)];
    } else {
      // insertion sort of the merged index, then the flat rank
      for (int q = 1; q < PA.rank; ++q) {
        const int32_t x = K[q];
        int u = q;
        while (u > 0 && K[u - 1] > x) { K[u] = K[u - 1]; --u; }
        K[u] = x;
      }
      v = a[flat_rank_sorted(PA, K)];
    }
    out[c] = v;
  }
}

static int grid_for(int64_t n) {
  const int64_t g = (n + 255) / 256;
  return (int)std::max<int64_t>(1, std::min<int64_t>(g, 148 * 32));
}

static int packed_len(int layout, int rank, int64_t dim, PlanView* P, int64_t* n) {
  if (layout != ST_LAYOUT_PERMCLS && layout != ST_LAYOUT_FLAT) { set_error("unknown layout %d", layout); return ST_ERR_INVALID; }
  int rc = get_device_plan(rank, dim, P);
  if (rc) return rc;
  *n = layout == ST_LAYOUT_PERMCLS ? P->total : P->flat_size;
  return ST_OK;
}

template <typename T>
static int elementwise_unary(int op, int layout, int rank, int64_t dim, const T* a, T* out, cudaStream_t stream) {
  if (op < 0 || op > ST_UN_RECIPROCAL) { set_error("unknown unary op %d", op); return ST_ERR_INVALID; }
  PlanView P;
  int64_t n = 0;
  int rc = packed_len(layout, rank, dim, &P, &n);
  if (rc) return rc;
  if (n == 0) return ST_OK;
  if (!a || !out) { set_error("null pointer"); return ST_ERR_INVALID; }
  unary_kernel<T><<<grid_for(n), 256, 0, stream>>>(P, layout, op, n, a, out);
  count_launch();
  return check_cuda(cudaGetLastError(), "unary_kernel");
}

template <typename T>
static int elementwise_binary(int op, int mode, int layout, int rank, int64_t dim, const T* a, const T* b, double s, T* out, cudaStream_t stream) {
  if (op < 0 || op > ST_BIN_POWER || mode < 0 || mode > 2) { set_error("unknown binary op %d / mode %d", op, mode); return ST_ERR_INVALID; }
  PlanView P;
  int64_t n = 0;
  int rc = packed_len(layout, rank, dim, &P, &n);
  if (rc) return rc;
  if (n == 0) return ST_OK;
  if (!a || !out || (mode == 0 && !b)) { set_error("null pointer"); return ST_ERR_INVALID; }
  binary_kernel<T><<<grid_for(n), 256, 0, stream>>>(P, layout, op, mode, n, a, b, (T)s, out);
  count_launch();
  return check_cuda(cudaGetLastError(), "binary_kernel");
}

template <typename T>
static int compare(int mode, int layout, int rank, int64_t dim, const T* a, const T* b, int b_scalar, double s, double rtol, double atol, int equal_nan,
                   T* mask, int* all_flag, cudaStream_t stream) {
  if (mode < 0 || mode > 1) { set_error("unknown comparison mode %d", mode); return ST_ERR_INVALID; }
  PlanView P;
  int64_t n = 0;
  int rc = packed_len(layout, rank, dim, &P, &n);
  if (rc) return rc;
  if (n == 0) return ST_OK;
  if (!a || (!b_scalar && !b)) { set_error("null pointer"); return ST_ERR_INVALID; }
  compare_kernel<T><<<grid_for(n), 256, 0, stream>>>(P, layout, mode, n, a, b, b_scalar, s, rtol, atol, equal_nan, mask, all_flag);
  count_launch();
  return check_cuda(cudaGetLastError(), "compare_kernel");
}

template <typename T>
static int slice(int layout, int rank, int64_t dim, int nfixed, const int32_t* d_fixed, const T* a, T* out, cudaStream_t stream) {
  if (nfixed < 1 || nfixed > rank) { set_error("%d fixed indices for a rank-%d tensor", nfixed, rank); return ST_ERR_INVALID; }
  PlanView PA, PB;
  int64_t na = 0, nb = 0;
  int rc = packed_len(layout, rank, dim, &PA, &na);
  if (rc) return rc;
  rc = packed_len(layout, rank - nfixed, rank - nfixed ? dim : 1, &PB, &nb);
  if (rc) return rc;
  if (nb == 0) return ST_OK;
  if (!a || !out || !d_fixed) { set_error("null pointer"); return ST_ERR_INVALID; }
  slice_kernel<T><<<grid_for(nb), 256, 0, stream>>>(PA, PB, layout, nfixed, d_fixed, a, out, nb);
  count_launch();
  return check_cuda(cudaGetLastError(), "slice_kernel");
}

}  // namespace st

using namespace st;

extern "C" {
int st_elementwise_unary_f64(int op, int layout, int rank, int64_t dim, const double* d_a, double* d_out, void* stream) {
  return elementwise_unary<double>(op, layout, rank, dim, d_a, d_out, (cudaStream_t)stream);
}
int st_elementwise_unary_f32(int op, int layout, int rank, int64_t dim, const float* d_a, float* d_out, void* stream) {
  return elementwise_unary<float>(op, layout, rank, dim, d_a, d_out, (cudaStream_t)stream);
}
int st_elementwise_binary_f64(int op, int mode, int layout, int rank, int64_t dim, const double* d_a, const double* d_b, double scalar, double* d_out,
                              void* stream) {
  return elementwise_binary<double>(op, mode, layout, rank, dim, d_a, d_b, scalar, d_out, (cudaStream_t)stream);
}
int st_elementwise_binary_f32(int op, int mode, int layout, int rank, int64_t dim, const float* d_a, const float* d_b, double scalar, float* d_out,
                              void* stream) {
  return elementwise_binary<float>(op, mode, layout, rank, dim, d_a, d_b, scalar, d_out, (cudaStream_t)stream);
}
int st_compare_f64(int mode, int layout, int rank, int64_t dim, const double* d_a, const double* d_b, int b_is_scalar, double scalar, double rtol,
                   double atol, int equal_nan, double* d_mask, int* d_all, void* stream) {
  return compare<double>(mode, layout, rank, dim, d_a, d_b, b_is_scalar, scalar, rtol, atol, equal_nan, d_mask, d_all, (cudaStream_t)stream);
}
int st_compare_f32(int mode, int layout, int rank, int64_t dim, const float* d_a, const float* d_b, int b_is_scalar, double scalar, double rtol, double atol,
                   int equal_nan, float* d_mask, int* d_all, void* stream) {
  return compare<float>(mode, layout, rank, dim, d_a, d_b, b_is_scalar, scalar, rtol, atol, equal_nan, d_mask, d_all, (cudaStream_t)stream);
}
int st_slice_f64(int layout, int rank, int64_t dim, int nfixed, const int32_t* d_fixed, const double* d_a, double* d_out, void* stream) {
  return slice<double>(layout, rank, dim, nfixed, d_fixed, d_a, d_out, (cudaStream_t)stream);
}
int st_slice_f32(int layout, int rank, int64_t dim, int nfixed, const int32_t* d_fixed, const float* d_a, float* d_out, void* stream) {
  return slice<float>(layout, rank, dim, nfixed, d_fixed, d_a, d_out, (cudaStream_t)stream);
}
}  // extern "C"
