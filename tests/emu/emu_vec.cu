// TEST-ONLY CPU emulation of vec_tail_kernel's control flow (symtensor_b200/csrc/st_vec.cu).
//
// Not part of libsymtensor_b200.so and never used by the product: it exists so that the index arithmetic of
// the tail-table kernel (strategy, segment / block walk, table build -- the host+device functions of
// st_vec_core.cuh) can be checked against the oracle in the GPU-less build container.  The CTA / warp / lane
// loops of the kernel are replayed serially.
#include <cstring>
#include <vector>

#include "../../symtensor_b200/csrc/st_common.cuh"
#include "../../symtensor_b200/csrc/st_vec_core.cuh"

namespace st {
bool compute_tail_strategy(const HostPlan* hp, int esize, int nwarps, int nst, std::vector<TailStrategy>& st, int32_t* tbl_cap, size_t* smem_bytes);
extern int g_force_tau;
}

using namespace st;

template <typename T>
static int emu(int rank, int64_t dim, const T* A, int64_t begin, int64_t end, const T* x, int nwarps, int grid, int64_t item_elems,
               int force_tau, double* out, int32_t* taus_out) {
  const HostPlan* hp = get_host_plan(rank, dim);
  if (!hp) return 1;
  const PlanView P = hp->host_view();
  std::vector<TailStrategy> strat;
  int32_t tbl_cap = 0;
  size_t smem = 0;
  g_force_tau = force_tau;
  const bool ok = compute_tail_strategy(hp, (int)sizeof(T), nwarps, 6, strat, &tbl_cap, &smem);
  g_force_tau = 0;
  if (!ok) return 3;
  if (taus_out) for (int c = 0; c < hp->ncls; ++c) taus_out[c] = strat[c].tau;
  std::vector<T> tbl(tbl_cap), xr(dim + 1), xs(dim + 1);
  std::vector<int32_t> blen(dim + 1);
  for (int64_t i = 0; i < dim; ++i) xs[i] = x[i];
  const int64_t n_items = (end - begin + item_elems - 1) / item_elems;
  double grand = 0.0;
  for (int cta = 0; cta < grid; ++cta) {
    int cur_cls = -1;
    int64_t cur_seg = -1;
    int32_t ctlE[ST_MAX_RANK];
    double ctl_wE = 0.0;
    std::vector<double> lane_total((size_t)nwarps * 32, 0.0);
    for (int64_t item = cta; item < n_items; item += grid) {
      const int64_t c0 = begin + item * item_elems;
      const int64_t c1 = (c0 + item_elems < end) ? c0 + item_elems : end;
      int64_t coord = c0;
      while (coord < c1) {
        const int ci = class_of_coord(P, coord);
        const ClassDesc& C = P.cls[ci];
        int64_t pos = coord - C.offset;
        if (pos >= C.size) { coord = P.offsets[ci + 1]; continue; }
        const int64_t pend = (C.size < c1 - C.offset) ? C.size : c1 - C.offset;
        const TailStrategy S = strat[ci];
        const T* Acls = A + (C.offset - begin);
        if (S.tau == 1) {
          const int64_t len = pend - pos;
          int64_t per = (len + nwarps - 1) / nwarps;
          per = (per + 31) / 32 * 32;
          for (int warp = 0; warp < nwarps; ++warp) {
            int64_t w0 = pos + (int64_t)warp * per;
            const int64_t w1 = (w0 + per < pend) ? w0 + per : pend;
            int32_t E[ST_MAX_RANK];
            std::vector<T> priv(dim + 1);
            while (w0 < w1) {
              const int64_t sidx = w0 / S.seg;
              const int64_t sbase = sidx * S.seg;
              const int64_t q1 = (S.seg < w1 - sbase) ? S.seg : w1 - sbase;
              const double wE = unrank_earlier<T>(P, C, sidx, xs.data(), E);
              for (int uu = 0; uu < S.Rt; ++uu) priv[uu] = xrel_pow<T>(xs.data(), E, S.nE, S.mu, uu);
              for (int lane = 0; lane < 32; ++lane)
                lane_total[warp * 32 + lane] += walk_range_direct<T>(P, S, priv.data(), priv.data(), nullptr, wE, Acls + sbase, w0 - sbase, q1, lane);
              w0 = sbase + q1;
            }
          }
          pos = pend;
        } else {
          while (pos < pend) {
            const int64_t sidx = pos / S.seg;
            const int64_t sbase = sidx * S.seg;
            const int64_t q0 = pos - sbase;
            const int64_t q1 = (S.seg < pend - sbase) ? S.seg : pend - sbase;
            if (cur_cls != ci || cur_seg != sidx) {
              ctl_wE = unrank_earlier<T>(P, C, sidx, xs.data(), ctlE);
              cur_cls = ci;
              cur_seg = sidx;
              for (int uu = 0; uu < S.Rt; ++uu) {
                xr[uu] = xrel_pow<T>(xs.data(), ctlE, S.nE, S.mu, uu);
                blen[uu] = (int32_t)binom_at(P.binom, P.rank, S.Rt - 1 - uu, S.tau);
              }
              const int nthreads = nwarps * 32;
              const int64_t per = (S.tbl_n + nthreads - 1) / nthreads;
              for (int t = 0; t < nthreads; ++t) {
                const int64_t q = (int64_t)t * per;
                build_table_slice<T>(P, S, xr.data(), tbl.data(), q, (q + per < S.tbl_n) ? q + per : S.tbl_n);
              }
            }
            const int64_t len = q1 - q0;
            int64_t per = (len + nwarps - 1) / nwarps;
            per = (per + 31) / 32 * 32;
            for (int warp = 0; warp < nwarps; ++warp) {
              const int64_t w0 = q0 + (int64_t)warp * per;
              const int64_t w1 = (w0 + per < q1) ? w0 + per : q1;
              if (w0 < w1)
                for (int lane = 0; lane < 32; ++lane)
                  lane_total[warp * 32 + lane] += walk_range_direct<T>(P, S, tbl.data(), xr.data(), blen.data(), ctl_wE, Acls + sbase, w0, w1, lane);
            }
            pos = sbase + q1;
          }
        }
        coord = C.offset + pos;
      }
    }
    for (double v : lane_total) grand += v;
  }
  *out = grand;
  return 0;
}

extern "C" {
int emu_contract_vec_f64(int rank, int64_t dim, const double* A, int64_t begin, int64_t end, const double* x, int nwarps, int grid,
                         int64_t item_elems, int force_tau, double* out, int32_t* taus_out) {
  return emu<double>(rank, dim, A, begin, end, x, nwarps, grid, item_elems, force_tau, out, taus_out);
}
int emu_contract_vec_f32(int rank, int64_t dim, const float* A, int64_t begin, int64_t end, const float* x, int nwarps, int grid,
                         int64_t item_elems, int force_tau, double* out, int32_t* taus_out) {
  return emu<float>(rank, dim, A, begin, end, x, nwarps, grid, item_elems, force_tau, out, taus_out);
}
}
