// GPU index-class enumerator: bulk rank / unrank of multi-indices for the permcls and flat layouts.
//
// Replaces the reference's Python generators and lookup registry, bit-exactly:
//   σindex_iter / _sub_σindex_iter                     symtensor/permcls_symtensor.py:288-347
//   indep_iter_repindex                                symtensor/permcls_symtensor.py:958-960
//   get_index_representative, PosRegistry lookups      symtensor/permcls_symtensor.py:375-381, 422-479
//   flat indep_iter_repindex, index_of_multicombination symtensor/flat_symtensor.py:39-50, 219-220
// One thread per component / index; these kernels are integer-ALU bound helpers (packing, tests), the hot
// contraction kernels embed the same device functions.
#include "st_common.cuh"

namespace st {

__global__ void permcls_unrank_kernel(PlanView P, int cls, int64_t begin, int64_t count, int32_t* __restrict__ out) {
  const ClassDesc C = P.cls[cls];
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    int32_t vals[ST_MAX_RANK];
    permcls_unrank_vals(P, C, begin + i, vals);
    int32_t* o = out + i * P.rank;
    int k = 0;
    for (int v = 0; v < C.nvals; ++v)
      for (int m = 0; m < C.mult[v]; ++m) o[k++] = vals[v];
  }
}

__global__ void permcls_rank_kernel(PlanView P, int64_t n, const int32_t* __restrict__ idx, int32_t* __restrict__ cls_out,
                                    int64_t* __restrict__ pos_out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int32_t in[ST_MAX_RANK], vals[ST_MAX_RANK];
    for (int k = 0; k < P.rank; ++k) in[k] = idx[i * P.rank + k];
    const int c = classify_index(P, in, vals);
    cls_out[i] = c;
    pos_out[i] = c < 0 ? -1 : permcls_rank_vals(P, P.cls[c], vals);
  }
}

__global__ void flat_unrank_kernel(PlanView P, int64_t begin, int64_t count, int32_t* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    int32_t s[ST_MAX_RANK];
    flat_unrank_sorted(P, begin + i, s);
    for (int k = 0; k < P.rank; ++k) out[i * P.rank + k] = s[k];
  }
}

__global__ void flat_rank_kernel(PlanView P, int64_t n, const int32_t* __restrict__ idx, int64_t* __restrict__ pos_out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int32_t s[ST_MAX_RANK];
    bool ok = true;
    for (int k = 0; k < P.rank; ++k) {
      const int32_t v = idx[i * P.rank + k];
      ok = ok && v >= 0 && v < P.dim;
      int u = k;
      while (u > 0 && s[u - 1] > v) { s[u] = s[u - 1]; --u; }
      s[u] = v;
    }
    pos_out[i] = ok ? flat_rank_sorted(P, s) : -1;
  }
}

static int grid_for(int64_t n, int threads) {
  int64_t g = (n + threads - 1) / threads;
  const int64_t cap = 148 * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace st

using namespace st;

extern "C" {

int st_permcls_unrank(int rank, int64_t dim, int32_t cls, int64_t begin, int64_t count, int32_t* d_idx_out, void* stream) {
  PlanView P;
  int rc = get_device_plan(rank, dim, &P);
  if (rc) return rc;
  const HostPlan* hp = get_host_plan(rank, dim);
  if (cls < 0 || cls >= hp->ncls) { set_error("class ordinal %d outside [0, %d)", cls, hp->ncls); return ST_ERR_INVALID; }
  if (begin < 0 || count < 0 || begin + count > hp->h_cls[cls].size) {
    set_error("range [%lld, %lld) outside class of size %lld", (long long)begin, (long long)(begin + count), (long long)hp->h_cls[cls].size);
    return ST_ERR_INVALID;
  }
  if (count == 0 || rank == 0) return ST_OK;
  if (!d_idx_out) { set_error("null output"); return ST_ERR_INVALID; }
  permcls_unrank_kernel<<<grid_for(count, 256), 256, 0, (cudaStream_t)stream>>>(P, cls, begin, count, d_idx_out);
  count_launch();
  return check_cuda(cudaGetLastError(), "permcls_unrank_kernel");
}

int st_permcls_rank(int rank, int64_t dim, int64_t n, const int32_t* d_idx, int32_t* d_cls_out, int64_t* d_pos_out, void* stream) {
  PlanView P;
  int rc = get_device_plan(rank, dim, &P);
  if (rc) return rc;
  if (n < 0) { set_error("negative count"); return ST_ERR_INVALID; }
  if (n == 0) return ST_OK;
  if (!d_idx || !d_cls_out || !d_pos_out) { set_error("null pointer"); return ST_ERR_INVALID; }
  permcls_rank_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(P, n, d_idx, d_cls_out, d_pos_out);
  count_launch();
  return check_cuda(cudaGetLastError(), "permcls_rank_kernel");
}

int st_flat_unrank(int rank, int64_t dim, int64_t begin, int64_t count, int32_t* d_idx_out, void* stream) {
  PlanView P;
  int rc = get_device_plan(rank, dim, &P);
  if (rc) return rc;
  if (begin < 0 || count < 0 || begin + count > P.flat_size) { set_error("range outside [0, size)"); return ST_ERR_INVALID; }
  if (count == 0 || rank == 0) return ST_OK;
  if (!d_idx_out) { set_error("null output"); return ST_ERR_INVALID; }
  flat_unrank_kernel<<<grid_for(count, 256), 256, 0, (cudaStream_t)stream>>>(P, begin, count, d_idx_out);
  count_launch();
  return check_cuda(cudaGetLastError(), "flat_unrank_kernel");
}

int st_flat_rank(int rank, int64_t dim, int64_t n, const int32_t* d_idx, int64_t* d_pos_out, void* stream) {
  PlanView P;
  int rc = get_device_plan(rank, dim, &P);
  if (rc) return rc;
  if (n < 0) { set_error("negative count"); return ST_ERR_INVALID; }
  if (n == 0) return ST_OK;
  if (!d_idx || !d_pos_out) { set_error("null pointer"); return ST_ERR_INVALID; }
  flat_rank_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(P, n, d_idx, d_pos_out);
  count_launch();
  return check_cuda(cudaGetLastError(), "flat_rank_kernel");
}

}  // extern "C"
