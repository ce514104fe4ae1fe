// TEST-ONLY CPU emulation of vec_tail_kernel's control flow (symtensor_b200/csrc/st_vec.cu).
//
// Not part of libsymtensor_b200.so and never used by the product: it exists so that the index arithmetic of
// the tail-table kernel (strategy, class / tile / segment / block walk, table build -- the host+device
// functions of st_vec_core.cuh) can be checked against the oracle in the GPU-less build container.  The CTA /
// warp / lane loops of the kernel are replayed serially with the same tile -> warp map.
#include <cstring>
#include <vector>

#include "../../symtensor_b200/csrc/st_common.cuh"
#include "../../symtensor_b200/csrc/st_vec_core.cuh"

namespace st {
bool compute_tail_strategy(const HostPlan* hp, int esize, int nwarps, std::vector<TailStrategy>& st, int32_t* tbl_cap, int32_t* binom_smem,
                           int32_t* cdesc_smem, size_t* smem_bytes);
extern int g_force_tau;
extern int64_t g_small_class;
}

using namespace st;

template <typename T>
static int emu(int rank, int64_t dim, const T* A, int64_t begin, int64_t end, const T* x, int nwarps, int grid, int64_t tile,
               int force_tau, int use_dir, int64_t small_class, double* out, int32_t* taus_out) {
  const HostPlan* hp = get_host_plan(rank, dim);
  if (!hp) return 1;
  const PlanView P = hp->host_view();
  std::vector<TailStrategy> strat;
  int32_t tbl_cap = 0, binom_smem = 0, cdesc_smem = 0;
  WarpScratch ws;
  size_t smem = 0;
  g_force_tau = force_tau;
  const int64_t small_saved = g_small_class;
  g_small_class = small_class;
  const bool ok = compute_tail_strategy(hp, (int)sizeof(T), nwarps, strat, &tbl_cap, &binom_smem, &cdesc_smem, &smem);
  g_force_tau = 0;
  g_small_class = small_saved;
  if (!ok) return 3;
  if (taus_out) for (int c = 0; c < hp->ncls; ++c) taus_out[c] = strat[c].tau;
  std::vector<T> tbl(tbl_cap), xr(dim + 1), xs(dim + 1), priv(dim + 1);
  std::vector<int32_t> blen(dim + 1);
  T bA[4][16 / sizeof(T)], bB[4][16 / sizeof(T)];
  for (int64_t i = 0; i < dim; ++i) xs[i] = x[i];
  const int64_t W = (int64_t)grid * nwarps;
  double grand = 0.0;
  // what vec_dir_kernel stores for the tile that starts at class position `pos`
  auto make_entry = [&](const ClassDesc& C, int64_t pos) {
    DirEntry e;
    for (int i = 0; i < ST_MAX_RANK; ++i) e.v[i] = 0;
    int32_t vals[ST_MAX_RANK];
    permcls_unrank_vals(P, C, pos, vals);
    for (int i = 0; i < C.nvals; ++i) e.v[i] = (uint16_t)vals[i];
    return e;
  };
  for (int cta = 0; cta < grid; ++cta) {
    int cur_cls = -1;
    int64_t cur_seg = -1;
    int32_t ctlE[ST_MAX_RANK];
    double ctl_wE = 0.0;
    std::vector<double> lane_total((size_t)nwarps * 32, 0.0);
    int64_t tile_base = 0;
    {  // phase 0: small classes, one component per thread of the grid
      int64_t sm_base = 0;
      const int64_t nthreads = W * 32;
      for (int ci = 0; ci < P.ncls; ++ci) {
        if (strat[ci].tau != 0) continue;
        const ClassDesc& C = P.cls[ci];
        const int64_t lo = (begin > C.offset ? begin : C.offset) - C.offset;
        const int64_t hi = (end < C.offset + C.size ? end : C.offset + C.size) - C.offset;
        if (lo >= hi) continue;
        const T* Acls = A + (C.offset - begin);
        for (int t = 0; t < nwarps * 32; ++t) {
          const int64_t tid = (int64_t)cta * nwarps * 32 + t;
          for (int64_t p = lo + (tid - sm_base % nthreads + nthreads) % nthreads; p < hi; p += nthreads) {
            int32_t vals[ST_MAX_RANK];
            permcls_unrank_vals(P, C, p, vals);
            double w = (double)C.gamma;
            for (int k = 0; k < C.nvals; ++k)
              for (int m = 0; m < C.mult[k]; ++m) w *= (double)xs[vals[k]];
            lane_total[t] += (double)Acls[p] * w;
          }
        }
        sm_base += hi - lo;
      }
    }
    auto build_shared = [&](int ci, const ClassDesc& C, const TailStrategy& S, int64_t sidx) -> double {
      ctl_wE = unrank_earlier<T>(P, C, sidx, xs.data(), ctlE, ws);
      cur_cls = ci;
      cur_seg = sidx;
      for (int uu = 0; uu < S.Rt; ++uu) {
        xr[uu] = xrel_pow<T>(xs.data(), ctlE, S.nE, S.mu, uu);
        blen[uu] = (int32_t)binom_at(P.binom, P.rank, S.Rt - 1 - uu, S.tau);
      }
      const int nthreads = nwarps * 32;
      int64_t nA, nB;
      table_scratch(P.binom, P.rank, S.Rt, S.tau, &nA, &nB);
      for (int t = 2; t <= S.tau; ++t) {  // level-by-level build, as build_shared_table in st_vec.cu
        const T* src = t == 2 ? xr.data() : table_level_buffer<T>(tbl.data(), S.tbl_n, nA, S.tau, t - 1);
        T* dst = table_level_buffer<T>(tbl.data(), S.tbl_n, nA, S.tau, t);
        for (int tid = 0; tid < nthreads; ++tid) build_table_level<T>(P.binom, P.rank, S.Rt, t, xr.data(), src, dst, tid, nthreads);
      }
      {  // cross-check against the serial reference form
        std::vector<T> ref(S.tbl_n);
        build_table_slice<T>(P, S, xr.data(), ref.data(), 0, S.tbl_n);
        for (int64_t q = 0; q < S.tbl_n; ++q) {
          const double d = (double)ref[q] - (double)tbl[q];
          if (d > 1e-6 * (double)ref[q] || -d > 1e-6 * (double)ref[q]) return -1.0;
        }
      }
      return 0.0;
    };
    for (int ci = 0; ci < P.ncls; ++ci) {
      if (strat[ci].tau == 0) continue;
      const ClassDesc& C = P.cls[ci];
      const int64_t coff = C.offset, csize = C.size;
      const int64_t lo = (begin > coff ? begin : coff) - coff;
      const int64_t hi = (end < coff + csize ? end : coff + csize) - coff;
      if (lo >= hi) continue;
      const TailStrategy S = strat[ci];
      const T* Acls = A + (coff - begin);
      const int64_t k0 = lo / tile, k1 = (hi + tile - 1) / tile;
      if (S.tau == 1 || S.nE == 0) {
        if (S.tau >= 2) { if (build_shared(ci, C, S, 0) < 0) return 7; }
        else cur_cls = -1;
        for (int warp = 0; warp < nwarps; ++warp) {
          const int64_t gw = (int64_t)cta * nwarps + warp;
          T cA[32][4][16 / sizeof(T)], cB[32][4][16 / sizeof(T)];  // per-lane register batches that live across tiles
          bool preloaded = false;
          for (int64_t j = (gw - tile_base % W + W) % W; k0 + j < k1; j += W) {
            int64_t w0 = (k0 + j) * tile, w1 = w0 + tile;
            const bool have_dir = use_dir && w0 >= lo;
            if (w0 < lo) w0 = lo;
            if (w1 > hi) w1 = hi;
            int32_t E[ST_MAX_RANK], u0[ST_MAX_RANK];
            DirEntry de;
            if (have_dir) de = make_entry(C, w0);
            if (S.tau >= 2) {
              if (have_dir) dir_decode<T>(C, S, de, xs.data(), E, u0);
              const int32_t* ui = have_dir ? u0 : nullptr;
              const int64_t nk = k0 + j + W;
              const int64_t batch_elems = 4 * 32 * (16 / (int64_t)sizeof(T));
              const bool chain = (w1 - w0 == tile) && nk < k1 && nk * tile >= lo && (nk + 1) * tile <= hi && (tile % (2 * batch_elems) == 0);
              for (int lane = 0; lane < 32; ++lane)
                lane_total[warp * 32 + lane] += walk_range<T, 4>(P, S, tbl.data(), xr.data(), blen.data(), ctl_wE, Acls, w0, w1, lane, ui,
                                                                 cA[lane], cB[lane], preloaded, chain ? Acls + nk * tile : nullptr, ws);
              preloaded = chain;
              continue;
            }
            bool first = true;
            while (w0 < w1) {
              const int64_t sidx = S.nE ? w0 / S.seg : 0;
              const int64_t sbase = sidx * S.seg;
              const int64_t q1 = (S.seg < w1 - sbase) ? S.seg : w1 - sbase;
              const bool from_dir = first && have_dir;
              double wE;
              if (from_dir) wE = dir_decode<T>(C, S, de, xs.data(), E, u0);
              else {
                wE = unrank_earlier<T>(P, C, sidx, xs.data(), E, ws);
                if (w0 == sbase) for (int i = 0; i < S.gt; ++i) u0[i] = i;
              }
              const int32_t* ui = (from_dir || w0 == sbase) ? u0 : nullptr;
              first = false;
              for (int uu = 0; uu < S.Rt; ++uu) priv[uu] = xrel_pow<T>(xs.data(), E, S.nE, S.mu, uu);
              for (int lane = 0; lane < 32; ++lane)
                lane_total[warp * 32 + lane] += walk_range<T, 4>(P, S, priv.data(), priv.data(), nullptr, wE, Acls + sbase, w0 - sbase, q1, lane, ui, bA, bB, false, nullptr, ws);
              w0 = sbase + q1;
            }
          }
        }
      } else {
        const int64_t ch0 = k0 / nwarps, ch1 = (k1 + nwarps - 1) / nwarps;
        const int64_t chunk = tile * nwarps;
        for (int64_t jc = ((int64_t)cta - (tile_base / nwarps) % grid + grid) % grid; ch0 + jc < ch1; jc += grid) {
          int64_t pos = (ch0 + jc) * chunk, pend = pos + chunk;
          if (pos < lo) pos = lo;
          if (pend > hi) pend = hi;
          while (pos < pend) {
            const int64_t sidx = pos / S.seg;
            const int64_t sbase = sidx * S.seg;
            const int64_t q0 = pos - sbase;
            const int64_t q1 = (S.seg < pend - sbase) ? S.seg : pend - sbase;
            if (cur_cls != ci || cur_seg != sidx) { if (build_shared(ci, C, S, sidx) < 0) return 7; }
            for (int warp = 0; warp < nwarps; ++warp) {
              const int64_t tk = (ch0 + jc) * nwarps + warp;
              int64_t w0 = tk * tile - sbase, w1 = w0 + tile;
              const bool at_tile = w0 >= q0;
              if (w0 < q0) w0 = q0;
              if (w1 > q1) w1 = q1;
              int32_t E[ST_MAX_RANK], u0[ST_MAX_RANK];
              const int32_t* ui = nullptr;
              if (w0 < w1) {
                if (w0 == 0) { for (int i = 0; i < S.gt; ++i) u0[i] = i; ui = u0; }
                else if (at_tile && use_dir) { dir_decode<T>(C, S, make_entry(C, tk * tile), xs.data(), E, u0); ui = u0; }
              }
              if (w0 < w1)
                for (int lane = 0; lane < 32; ++lane)
                  lane_total[warp * 32 + lane] += walk_range<T, 4>(P, S, tbl.data(), xr.data(), blen.data(), ctl_wE, Acls + sbase, w0, w1, lane, ui, bA, bB, false, nullptr, ws);
            }
            pos = sbase + q1;
          }
        }
      }
      tile_base += k1 - k0;
    }
    for (double v : lane_total) grand += v;
  }
  *out = grand;
  return 0;
}

extern "C" {
int emu_contract_vec_f64(int rank, int64_t dim, const double* A, int64_t begin, int64_t end, const double* x, int nwarps, int grid,
                         int64_t tile_elems, int force_tau, int use_dir, int64_t small_class, double* out, int32_t* taus_out) {
  return emu<double>(rank, dim, A, begin, end, x, nwarps, grid, tile_elems, force_tau, use_dir, small_class, out, taus_out);
}
int emu_contract_vec_f32(int rank, int64_t dim, const float* A, int64_t begin, int64_t end, const float* x, int nwarps, int grid,
                         int64_t tile_elems, int force_tau, int use_dir, int64_t small_class, double* out, int32_t* taus_out) {
  return emu<float>(rank, dim, A, begin, end, x, nwarps, grid, tile_elems, force_tau, use_dir, small_class, out, taus_out);
}
}
