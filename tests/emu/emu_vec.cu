// TEST-ONLY CPU emulation of vec_ring_kernel's control flow (symtensor_b200/csrc/st_vec.cu).
//
// Not part of libsymtensor_b200.so and never used by the product: it exists so that the index arithmetic of
// the ring kernel (strategy, class / chunk / tile / segment / block walk, table build, and the per-warp stream
// cursors of RingSrc -- the host+device code of st_vec_core.cuh) can be checked against the oracle in the
// GPU-less build container.  The program of every (CTA, warp, lane) is replayed serially with the kernel's
// tile -> warp map; a bulk copy is a memcpy at issue time, so a mismatch between what the producer cursor
// copies and what the consumer cursor reads shows up as a wrong result or a counter mismatch.
#include <cstring>
#include <vector>

#include "../../symtensor_b200/csrc/st_common.cuh"
#include "../../symtensor_b200/csrc/st_vec_core.cuh"

namespace st {
bool compute_tail_strategy(const HostPlan* hp, int esize, int nwarps, std::vector<TailStrategy>& st, int32_t* tbl_cap, int32_t* binom_smem,
                           int32_t* cdesc_smem, size_t* smem_bytes, size_t reserve, double block_cost);
extern int g_force_tau;
extern int g_ring_direct;
extern int64_t g_ring_table_max;
extern int64_t g_small_class;
}

using namespace st;

template <typename T>
static int emu(int rank, int64_t dim, const T* A, int64_t begin, int64_t end, const T* x, int nwarps, int grid, int64_t tile, int Bel, int R,
               int force_tau, int use_dir, int64_t small_class, double* out, int32_t* taus_out) {
  const HostPlan* hp = get_host_plan(rank, dim);
  if (!hp) return 1;
  if (tile % Bel) return 2;
  const PlanView P = hp->host_view();
  std::vector<TailStrategy> strat;
  int32_t tbl_cap = 0, binom_smem = 0, cdesc_smem = 0;
  size_t smem = 0;
  // -1: cost model with no table budget (pure column walk); -2: forced pair walk with the default budget (all rows
  // from the suffix table at test sizes); -3: forced pair walk with a 2 KB table (hybrid: both parts in play)
  g_force_tau = force_tau < 0 ? (force_tau == -1 ? 0 : 2) : force_tau;
  const int direct_saved = g_ring_direct;
  const int64_t tmax_saved = g_ring_table_max;
  if (force_tau == -1) g_ring_table_max = 0;
  if (force_tau <= -2) g_ring_direct = 2;
  if (force_tau == -3) g_ring_table_max = 2048;
  const int64_t small_saved = g_small_class;
  g_small_class = small_class;
  const bool ok = compute_tail_strategy(hp, (int)sizeof(T), nwarps, strat, &tbl_cap, &binom_smem, &cdesc_smem, &smem,
                                        (size_t)nwarps * R * Bel * sizeof(T) + 1024, 40.0);
  g_force_tau = 0;
  g_ring_direct = direct_saved;
  g_ring_table_max = tmax_saved;
  g_small_class = small_saved;
  if (!ok) return 3;
  if (taus_out) for (int c = 0; c < hp->ncls; ++c) taus_out[c] = strat[c].tau;
  std::vector<T> xs(dim + 1);
  for (int64_t i = 0; i < dim; ++i) xs[i] = x[i];
  const int G = grid;
  const int64_t W = (int64_t)G * nwarps;
  std::vector<ClsInfo> cls(P.ncls);
  {
    int64_t tb = 0;
    for (int c = 0; c < P.ncls; ++c) {
      cls[c].offset = P.cls[c].offset;
      cls[c].size = P.cls[c].size;
      cls[c].tile_base = tb;
      cls[c].sbase = 0;
      cls[c].S = strat[c];
      tb += (P.cls[c].size + tile - 1) / tile;
    }
  }
  // use_dir bit 0: tile directory; the bits above: g >= 1 -> the DYNAMIC deal (claims served in replay order), groups of g
  // tiles, one round of single tiles, a costly tail of 2 tiles per class
  const int group = use_dir >> 1;
  use_dir &= 1;
  std::vector<ClsRun> run(P.ncls);
  std::vector<int64_t> ntail(P.ncls, 2);
  std::vector<unsigned long long> counters(P.ncls, 0ULL);
  if (group >= 1) make_runs(cls.data(), P.ncls, begin, end, tile, nwarps, G, run.data(), ntail.data(), group, 1, 37);
  else make_runs(cls.data(), P.ncls, begin, end, tile, nwarps, G, run.data());
  if (group >= 1) {  // the deal is a bijection: every tile of the class range belongs to exactly one entry, and the inverse agrees
    for (int c = 0; c < P.ncls; ++c) {
      const ClsRun& r = run[c];
      if (r.mode != 1) continue;
      std::vector<int> seen(r.k1 - r.k0, 0);
      for (int64_t n = 0; n < r.nd; ++n) {
        int32_t tn = 0, tn2 = 0;
        const int64_t t = deal_to_tile(r, n, tn);
        if (t < r.k0 || t + tn > r.k1 || tn < 1) return 10;
        if (tile_to_deal(r, t, tn2) != n || tn2 != tn) return 11;
        for (int i = 0; i < tn; ++i) if (seen[t - r.k0 + i]++) return 12;
      }
      for (int v : seen) if (v != 1) return 13;
      int32_t tn = 0;
      if (deal_to_tile(r, r.nd, tn) < r.k1) return 14;
    }
  }
  // what vec_dir_kernel stores for the tile that starts at class position `pos`
  auto make_entry = [&](const ClassDesc& C, int64_t pos) {
    DirEntry e;
    for (int i = 0; i < ST_MAX_RANK; ++i) e.v[i] = 0;
    int32_t vals[ST_MAX_RANK];
    permcls_unrank_vals(P, C, pos, vals);
    for (int i = 0; i < C.nvals; ++i) e.v[i] = (uint16_t)vals[i];
    return e;
  };
  double grand = 0.0;
  {  // small classes, one component per thread of the grid
    int64_t sm_base = 0;
    const int64_t nthreads = W * 32;
    for (int ci = 0; ci < P.ncls; ++ci) {
      if (strat[ci].tau != 0) continue;
      const ClassDesc& C = P.cls[ci];
      const int64_t lo = (begin > C.offset ? begin : C.offset) - C.offset;
      const int64_t hi = (end < C.offset + C.size ? end : C.offset + C.size) - C.offset;
      if (lo >= hi) continue;
      const T* Acls = A + (C.offset - begin);
      for (int64_t tid = 0; tid < nthreads; ++tid) {
        for (int64_t p = lo + (tid - sm_base % nthreads + nthreads) % nthreads; p < hi; p += nthreads) {
          int32_t vals[ST_MAX_RANK];
          permcls_unrank_vals(P, C, p, vals);
          double w = (double)C.gamma;
          for (int k = 0; k < C.nvals; ++k)
            for (int m = 0; m < C.mult[k]; ++m) w *= (double)xs[vals[k]];
          grand += (double)Acls[p] * w;
        }
      }
      sm_base += hi - lo;
    }
  }
  // the tile directory, as vec_dir_kernel writes it
  std::vector<DirEntry> dir;
  for (int c = 0; c < P.ncls; ++c) {
    const int64_t nt = (P.cls[c].size + tile - 1) / tile;
    for (int64_t t = 0; t < nt; ++t) dir.push_back(make_entry(P.cls[c], t * tile));
  }
  std::vector<TileQ> queue(R + 2);
  std::vector<T> tbl(tbl_cap), xr(dim + 1), priv(dim + 1), ring((size_t)R * Bel);
  std::vector<int32_t> blen(dim + 1);
  std::vector<unsigned long long> counters_at_warp_start;
  for (int cta = 0; cta < G; ++cta)
    for (int warp = 0; warp < nwarps; ++warp)
      for (int lane = 0; lane < 32; ++lane) {
        // one thread's program, start to end (the 32 lanes of a warp see the same claims)
        if (lane == 0) counters_at_warp_start = counters;
        else counters = counters_at_warp_start;
        WarpScratch ws;
        int cur_cls = -1;
        int64_t cur_seg = -1;
        int32_t ctlE[ST_MAX_RANK];
        double ctl_wE = 0.0;
        double total = 0.0;
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
          if (S.direct && S.k0 < S.Rt - 1)
            for (int tid = 0; tid < nthreads; ++tid) build_pair_suffix<T>(S.Rt, S.k0, xr.data(), tbl.data(), tid, nthreads);
          for (int t = 2; t <= S.tau && !S.direct; ++t) {  // level-by-level build, as build_shared_table in st_vec.cu
            const T* src = t == 2 ? xr.data() : table_level_buffer<T>(tbl.data(), S.tbl_n, nA, S.tau, t - 1);
            T* dst = table_level_buffer<T>(tbl.data(), S.tbl_n, nA, S.tau, t);
            for (int tid = 0; tid < nthreads; ++tid) build_table_level<T>(P.binom, P.rank, S.Rt, t, xr.data(), src, dst, tid, nthreads);
          }
          if (lane == 0 && !S.direct) {  // cross-check against the serial reference form
            std::vector<T> ref(S.tbl_n);
            build_table_slice<T>(P, S, xr.data(), ref.data(), 0, S.tbl_n);
            for (int64_t q = 0; q < S.tbl_n; ++q) {
              const double d = (double)ref[q] - (double)tbl[q];
              if (d > 1e-6 * (double)ref[q] || -d > 1e-6 * (double)ref[q]) return -1.0;
            }
          }
          return 0.0;
        };
        RingSrc<T> src;
        src.ring = ring.data();
        src.bar0 = 0;
        src.R = R;
        src.Bel = Bel;
        src.lane = lane;
        src.run = run.data();
        src.cls = cls.data();
        src.A = A;
        src.dir = use_dir ? dir.data() : nullptr;
        src.ctr = group >= 1 ? counters.data() : nullptr;
        src.queue = queue.data();
        src.QD = R + 2;
        src.begin = begin;
        src.tile = tile;
        src.ondemand = 2;
        src.ncls = P.ncls;
        src.NW = nwarps;
        src.G = G;
        src.warp = warp;
        src.cta = cta;
        src.start();
        int priv_cls = -1;
        int64_t priv_seg = -1;
        double priv_wE = 0.0;
        for (int ci = 0; ci < P.ncls; ++ci) {
          const ClsRun rr = run[ci];
          if (!rr.mode) continue;
          const TailStrategy S = strat[ci];
          const ClassDesc& C = P.cls[ci];
          const int64_t coff = C.offset;
          const T* Acls = A + (coff - begin);
          const int64_t lo = rr.lo, hi = rr.hi, k0 = rr.k0, k1 = rr.k1, ch0 = rr.ch0, ch1 = rr.ch1;
          if (rr.mode == 1) {
            const bool shared_tbl = S.tau >= 2 && S.nE == 0;
            if (shared_tbl) { if (build_shared(ci, C, S, 0) < 0) return 7; priv_cls = -1; }
            else cur_cls = -1;
            while (src.head_is(ci)) {
              const TileQ& tq = src.pop();
              const int64_t tk = tq.tk;
              int32_t tn;
              (void)tile_to_deal(rr, tk, tn);
              int64_t w0 = tk * tile, w1 = w0 + tn * tile;
              const bool have_dir = use_dir && w0 >= lo;
              if (w0 < lo) w0 = lo;
              if (w1 > hi) w1 = hi;
              const int64_t tw0 = w0;
              src.open_tile(Acls + w0, (int)(w1 - w0));
              int32_t E[ST_MAX_RANK], u0[ST_MAX_RANK];
              const DirEntry& de = tq.de;
              if (shared_tbl) {
                if (have_dir) for (int i = 0; i < S.gt; ++i) u0[i] = de.v[i];
                total += walk_tile_any<T>(P, S, tbl.data(), xr.data(), blen.data(), ctl_wE, w0, (int)(w1 - w0), 0, lane, have_dir ? u0 : nullptr, src, ws);
              } else {
                const bool gap = S.direct && S.nE != 0 && S.mu == 1;
                bool first = true;
                while (w0 < w1) {
                  const int64_t sidx = S.nE ? w0 / S.seg : 0;
                  const int64_t sbase = sidx * S.seg;
                  const int64_t q1 = (S.seg < w1 - sbase) ? S.seg : w1 - sbase;
                  const bool from_dir = first && have_dir;
                  if (from_dir) {
                    const double wE = dir_decode<T>(C, S, de, xs.data(), E, u0);
                    if (priv_cls != ci || priv_seg != sidx) priv_wE = wE;
                  } else if (w0 == sbase) {
                    for (int i = 0; i < S.gt; ++i) u0[i] = i;
                  }
                  if (priv_cls != ci || priv_seg != sidx) {
                    if (!from_dir) priv_wE = unrank_earlier<T>(P, C, sidx, xs.data(), E, ws);
                    if (!gap) for (int uu = 0; uu < S.Rt; ++uu) priv[uu] = xrel_pow<T>(xs.data(), E, S.nE, S.mu, uu);
                    priv_cls = ci;
                    priv_seg = sidx;
                  }
                  for (int i = 0; i < S.nE; ++i) ws.E[i] = E[i];
                  total += walk_tile_any<T>(P, S, gap ? xs.data() : priv.data(), gap ? xs.data() : priv.data(), nullptr, priv_wE, w0 - sbase, (int)(sbase + q1 - w0), (int)(w0 - tw0), lane,
                                        (from_dir || w0 == sbase) ? u0 : nullptr, src, ws);
                  w0 = sbase + q1;
                  first = false;
                }
              }
              src.close_tile();
            }
          } else {
            const int64_t chunk = tile * nwarps;
            for (int64_t jc = src.jc0(ci); ch0 + jc < ch1; jc += G) {
              const int64_t tk = (ch0 + jc) * nwarps + warp;
              const bool exists = tk >= k0 && tk < k1;
              int64_t tw0 = tk * tile, tw1 = tw0 + tile;
              if (tw0 < lo) tw0 = lo;
              if (tw1 > hi) tw1 = hi;
              const TileQ* tq = nullptr;
              if (exists) {
                tq = &src.pop();
                if (tq->tk != tk || tq->ci != ci) return 9;  // the producer cursor queued another tile
                src.open_tile(Acls + tw0, (int)(tw1 - tw0));
              }
              int64_t pos = (ch0 + jc) * chunk, pend = pos + chunk;
              if (pos < lo) pos = lo;
              if (pend > hi) pend = hi;
              while (pos < pend) {
                const int64_t sidx = pos / S.seg;
                const int64_t sbase = sidx * S.seg;
                const int64_t q0 = pos - sbase;
                const int64_t q1 = (S.seg < pend - sbase) ? S.seg : pend - sbase;
                if (cur_cls != ci || cur_seg != sidx) { if (build_shared(ci, C, S, sidx) < 0) return 7; priv_cls = -1; }
                if (exists) {
                  int64_t w0 = tk * tile - sbase, w1 = w0 + tile;
                  const bool at_tile = w0 >= q0;
                  if (w0 < q0) w0 = q0;
                  if (w1 > q1) w1 = q1;
                  if (w0 < w1) {
                    int32_t E[ST_MAX_RANK], u0[ST_MAX_RANK];
                    const int32_t* ui = nullptr;
                    if (w0 == 0) { for (int i = 0; i < S.gt; ++i) u0[i] = i; ui = u0; }
                    else if (at_tile && use_dir) { dir_decode<T>(C, S, tq->de, xs.data(), E, u0); ui = u0; }
                    total += walk_tile_any<T>(P, S, tbl.data(), xr.data(), blen.data(), ctl_wE, w0, (int)(w1 - w0), (int)(sbase + w0 - tw0), lane, ui, src, ws);
                  }
                }
                pos = sbase + q1;
              }
              if (exists) src.close_tile();
            }
          }
        }
        if (src.pvalid || src.n_issued != src.n_waited) return 8;  // the two cursors of the stream disagree
        grand += total;
      }
  *out = grand;
  return 0;
}

extern "C" {
int emu_contract_vec_f64(int rank, int64_t dim, const double* A, int64_t begin, int64_t end, const double* x, int nwarps, int grid,
                         int64_t tile_elems, int ring_elems, int ring_slots, int force_tau, int use_dir, int64_t small_class, double* out,
                         int32_t* taus_out) {
  return emu<double>(rank, dim, A, begin, end, x, nwarps, grid, tile_elems, ring_elems, ring_slots, force_tau, use_dir, small_class, out, taus_out);
}
int emu_contract_vec_f32(int rank, int64_t dim, const float* A, int64_t begin, int64_t end, const float* x, int nwarps, int grid,
                         int64_t tile_elems, int ring_elems, int ring_slots, int force_tau, int use_dir, int64_t small_class, double* out,
                         int32_t* taus_out) {
  return emu<float>(rank, dim, A, begin, end, x, nwarps, grid, tile_elems, ring_elems, ring_slots, force_tau, use_dir, small_class, out, taus_out);
}
}
