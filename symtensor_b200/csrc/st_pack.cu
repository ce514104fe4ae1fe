// Dense <-> packed conversion on the device (the step either side of every op).
//
// Replaces, for device-resident tensors, the reference's Python loops
//   PermClsSymmetricTensor._validate_data, ndarray branch   symtensor/permcls_symtensor.py:599-618
//     (utils.is_symmetric :563-578 -- allclose over all axis permutations -- then the gather by σindex_iter;
//      with symmetrize=True utils.symmetrize :507-532 first: the plain mean over all rank! axis permutations)
//   PermClsSymmetricTensor.todense / PermClsTorchSymmetricTensor.todense   permcls_symtensor.py:883-887, torch_symtensor.py:564-568
//   FlatSymmetricTensor.__init__ / todense                  symtensor/flat_symtensor.py:100-110, 251-256
// Both kernels are HBM-bound on the dense side (dim^rank elements read or written once, coalesced); the packed side is
// a gather through the on-device rank / unrank of the index enumerator (st_common.cuh).
#include "st_common.cuh"

namespace st {

// dense (row-major, dim^rank) element offset of a multi-index
ST_HD int64_t dense_offset(int rank, int64_t dim, const int32_t* idx) {
  int64_t o = 0;
  for (int k = 0; k < rank; ++k) o = o * dim + idx[k];
  return o;
}

// lexicographic successor among the distinct permutations of a multiset (std::next_permutation); false: was the last
ST_HD bool next_perm(int32_t* a, int n) {
  int i = n - 2;
  while (i >= 0 && a[i] >= a[i + 1]) --i;
  if (i < 0) return false;
  int j = n - 1;
  while (a[j] <= a[i]) --j;
  int32_t t = a[i]; a[i] = a[j]; a[j] = t;
  for (int l = i + 1, r = n - 1; l < r; ++l, --r) { t = a[l]; a[l] = a[r]; a[r] = t; }
  return true;
}

// packed -> dense: one thread per dense element (coalesced stores); its multi-index is sorted / classified and ranked.
// Run-time-rank form (any rank, both layouts): the index arrays live in local memory.
template <typename T>
__global__ void __launch_bounds__(256) unpack_dense_kernel(PlanView P, int layout, const T* __restrict__ packed, T* __restrict__ dense, int64_t n_dense) {
  const int r = P.rank;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_dense; e += (int64_t)gridDim.x * blockDim.x) {
    int32_t idx[ST_MAX_RANK], vals[ST_MAX_RANK];
    int64_t q = e;
    for (int k = r - 1; k >= 0; --k) {
      const int64_t d = q / P.dim;
      idx[k] = (int32_t)(q - d * P.dim);
      q = d;
    }
    int64_t pos;
    if (layout == ST_LAYOUT_PERMCLS) {
      const int c = classify_index(P, idx, vals);
      pos = P.cls[c].offset + permcls_rank_vals(P, P.cls[c], vals);
    } else {
      for (int k = 1; k < r; ++k) {  // insertion sort
        const int32_t v = idx[k];
        int u = k;
        while (u > 0 && idx[u - 1] > v) { idx[u] = idx[u - 1]; --u; }
        idx[u] = v;
      }
      pos = flat_rank_sorted(P, idx);
    }
    dense[e] = packed[pos];
  }
}

// Compile-time-rank form for the flat layout (ranks 1..8, dim^rank < 2^32): everything in registers.  One thread per
// dense element; the flat rank of the sorted index is a sum of R per-position terms
// F[t][v] = C(dim - 1 + t - v, t + 1) (t-th position from the end holding value v), kept in shared memory, and the sort is
// a fully unrolled odd-even transposition network.  HBM-bound on the dense store (the packed gather hits L2: the packed
// tensor is R! times smaller).
template <typename T, int R>
__global__ void __launch_bounds__(256) unpack_flat_fast_kernel(PlanView P, const T* __restrict__ packed, T* __restrict__ dense, uint32_t n_dense) {
  extern __shared__ uint32_t F_s[];  // [R][dim]
  const uint32_t dim = (uint32_t)P.dim;
  for (uint32_t i = threadIdx.x; i < R * dim; i += blockDim.x) {
    const uint32_t t = i / dim, v = i - t * dim;
    F_s[i] = (uint32_t)binom_at(P.binom, P.rank, (int64_t)dim - 1 + t - v, (int)t + 1);
  }
  __syncthreads();
  const uint32_t last = (uint32_t)(P.flat_size - 1);
  for (uint64_t e0 = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; e0 < n_dense; e0 += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t q = (uint32_t)e0;
    uint32_t s[R];
#pragma unroll
    for (int k = R - 1; k >= 0; --k) {
      const uint32_t d = q / dim;
      s[k] = q - d * dim;
      q = d;
    }
#pragma unroll
    for (int pass = 0; pass < R; ++pass) {
#pragma unroll
      for (int k = pass & 1; k + 1 < R; k += 2) {
        const uint32_t lo = min(s[k], s[k + 1]), hi = max(s[k], s[k + 1]);
        s[k] = lo;
        s[k + 1] = hi;
      }
    }
    uint32_t pos = last;
#pragma unroll
    for (int t = 0; t < R; ++t) pos -= F_s[t * dim + s[R - 1 - t]];
    dense[e0] = packed[pos];
  }
}

// dense -> packed: one thread per packed coordinate of [begin, end).  mode 0: take the representative entry and CHECK
// the symmetry of the dense array over the component's distinct permutations, numpy.allclose semantics
// (|a_p - a_q| <= atol + rtol |a_q| for every ordered pair; one pass: max a <= min (a + tol(a)), min a >= max (a - tol(a)));
// a violation (or a NaN facing a number; a component that is NaN in all its
// permutations is symmetric, equal_nan=True in the reference) raises *flag.  mode 1: symmetrize -- the mean over the distinct permutations, which is the
// reference's mean over all rank! axis permutations (every distinct one appears rank!/gamma times).
template <typename T>
__global__ void __launch_bounds__(256) pack_dense_kernel(PlanView P, int layout, const T* __restrict__ dense, T* __restrict__ packed, int64_t begin,
                                                         int64_t end, int mode, double rtol, double atol, int* __restrict__ flag) {
  const int r = P.rank;
  bool bad = false;
  for (int64_t c = begin + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < end; c += (int64_t)gridDim.x * blockDim.x) {
    int32_t K[ST_MAX_RANK];
    if (layout == ST_LAYOUT_PERMCLS) {
      if (!permcls_coord_sorted(P, c, K)) { packed[c - begin] = T(0); continue; }
    } else {
      flat_unrank_sorted(P, c, K);
    }
    // K is sorted: the first of the distinct permutations
    const double rep = (double)dense[dense_offset(r, P.dim, K)];
    double sum = 0.0, mx = rep, mn = rep, minf = rep + (atol + rtol * fabs(rep)), maxg = rep - (atol + rtol * fabs(rep));
    int64_t cnt = 0, n_nan = 0;
    do {
      const double a = (double)dense[dense_offset(r, P.dim, K)];
      const double tol = atol + rtol * fabs(a);
      sum += a;
      ++cnt;
      n_nan += a != a ? 1 : 0;
      mx = a > mx ? a : mx;
      mn = a < mn ? a : mn;
      minf = a + tol < minf ? a + tol : minf;
      maxg = a - tol > maxg ? a - tol : maxg;
    } while (next_perm(K, r));
    if (mode == 1) {
      packed[c - begin] = (T)(sum / (double)cnt);
    } else {
      packed[c - begin] = (T)rep;
      // NaNs: numpy.allclose(..., equal_nan=True) in utils.is_symmetric accepts a component whose permutations are ALL NaN
      if (n_nan != 0 ? n_nan != cnt : !(mx <= minf && mn >= maxg)) bad = true;
    }
  }
  if (bad && flag) atomicOr(flag, 1);
}

static int grid_for_pack(int64_t n, int threads) {
  int64_t g = (n + threads - 1) / threads;
  const int64_t cap = 148 * 32;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

static int dense_size(int rank, int64_t dim, int64_t* out) {
  int64_t n = 1;
  for (int k = 0; k < rank; ++k) {
    if (dim != 0 && n > INT64_MAX / dim) { set_error("dim^rank does not fit int64"); return ST_ERR_OVERFLOW; }
    n *= dim;
  }
  *out = n;
  return ST_OK;
}

template <typename T>
static int unpack_dense(int layout, int rank, int64_t dim, const T* d_packed, T* d_dense, cudaStream_t stream) {
  if (layout != ST_LAYOUT_PERMCLS && layout != ST_LAYOUT_FLAT) { set_error("unknown layout %d", layout); return ST_ERR_INVALID; }
  PlanView P;
  int rc = get_device_plan(rank, dim, &P);
  if (rc) return rc;
  int64_t n = 0;
  rc = dense_size(rank, dim, &n);
  if (rc) return rc;
  if (n == 0) return ST_OK;
  if (!d_packed || !d_dense) { set_error("null pointer"); return ST_ERR_INVALID; }
  const size_t smem = (size_t)rank * dim * sizeof(uint32_t);
  bool fast = false;
  if (layout == ST_LAYOUT_FLAT && rank >= 1 && rank <= 8 && n < 4294967296LL && smem <= 40 * 1024) {
    const int grid = grid_for_pack(n, 256);
    const uint32_t nn = (uint32_t)n;
    fast = true;
    switch (rank) {
      case 1: unpack_flat_fast_kernel<T, 1><<<grid, 256, smem, stream>>>(P, d_packed, d_dense, nn); break;
      case 2: unpack_flat_fast_kernel<T, 2><<<grid, 256, smem, stream>>>(P, d_packed, d_dense, nn); break;
      case 3: unpack_flat_fast_kernel<T, 3><<<grid, 256, smem, stream>>>(P, d_packed, d_dense, nn); break;
      case 4: unpack_flat_fast_kernel<T, 4><<<grid, 256, smem, stream>>>(P, d_packed, d_dense, nn); break;
      case 5: unpack_flat_fast_kernel<T, 5><<<grid, 256, smem, stream>>>(P, d_packed, d_dense, nn); break;
      case 6: unpack_flat_fast_kernel<T, 6><<<grid, 256, smem, stream>>>(P, d_packed, d_dense, nn); break;
      case 7: unpack_flat_fast_kernel<T, 7><<<grid, 256, smem, stream>>>(P, d_packed, d_dense, nn); break;
      default: unpack_flat_fast_kernel<T, 8><<<grid, 256, smem, stream>>>(P, d_packed, d_dense, nn); break;
    }
  }
  if (!fast) unpack_dense_kernel<T><<<grid_for_pack(n, 256), 256, 0, stream>>>(P, layout, d_packed, d_dense, n);
  count_launch();
  return check_cuda(cudaGetLastError(), "unpack_dense_kernel");
}

template <typename T>
static int pack_dense(int layout, int rank, int64_t dim, const T* d_dense, T* d_packed, int64_t begin, int64_t end, int symmetrize, double rtol,
                      double atol, int* d_flag, cudaStream_t stream) {
  if (layout != ST_LAYOUT_PERMCLS && layout != ST_LAYOUT_FLAT) { set_error("unknown layout %d", layout); return ST_ERR_INVALID; }
  PlanView P;
  int rc = get_device_plan(rank, dim, &P);
  if (rc) return rc;
  const int64_t total = layout == ST_LAYOUT_PERMCLS ? P.total : P.flat_size;
  if (begin < 0 || end < begin || end > total) { set_error("range [%lld, %lld) outside [0, %lld]", (long long)begin, (long long)end, (long long)total); return ST_ERR_INVALID; }
  int64_t n = 0;
  rc = dense_size(rank, dim, &n);
  if (rc) return rc;
  if (end == begin) return ST_OK;
  if (!d_dense || !d_packed || (!symmetrize && !d_flag)) { set_error("null pointer"); return ST_ERR_INVALID; }
  pack_dense_kernel<T><<<grid_for_pack(end - begin, 256), 256, 0, stream>>>(P, layout, d_dense, d_packed, begin, end, symmetrize ? 1 : 0, rtol, atol, d_flag);
  count_launch();
  return check_cuda(cudaGetLastError(), "pack_dense_kernel");
}

}  // namespace st

using namespace st;

extern "C" {

int st_unpack_dense_f64(int layout, int rank, int64_t dim, const double* d_packed, double* d_dense, void* stream) {
  return unpack_dense<double>(layout, rank, dim, d_packed, d_dense, (cudaStream_t)stream);
}
int st_unpack_dense_f32(int layout, int rank, int64_t dim, const float* d_packed, float* d_dense, void* stream) {
  return unpack_dense<float>(layout, rank, dim, d_packed, d_dense, (cudaStream_t)stream);
}
int st_pack_dense_f64(int layout, int rank, int64_t dim, const double* d_dense, double* d_packed, int64_t begin, int64_t end, int symmetrize,
                      double rtol, double atol, int* d_flag, void* stream) {
  return pack_dense<double>(layout, rank, dim, d_dense, d_packed, begin, end, symmetrize, rtol, atol, d_flag, (cudaStream_t)stream);
}
int st_pack_dense_f32(int layout, int rank, int64_t dim, const float* d_dense, float* d_packed, int64_t begin, int64_t end, int symmetrize,
                      double rtol, double atol, int* d_flag, void* stream) {
  return pack_dense<float>(layout, rank, dim, d_dense, d_packed, begin, end, symmetrize, rtol, atol, d_flag, (cudaStream_t)stream);
}

}  // extern "C"
