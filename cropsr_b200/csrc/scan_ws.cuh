// Warp-specialised variant of the two-phase scan + score kernel (sm_100a).
//
// Same HBM layout, same count phase / grid barrier / prefix blocks as k_scan_score
// (scan.cuh); what changes is how an SM spends its warps in the emit phase.  There, all
// warps of a CTA alternate between the front end of a tile (hit tests, warp scan, compaction
// into the hit lists: ~300 instructions per warp, latency bound) and the per-candidate body
// (window, Rule Set 1, stores: issue bound), phase-locked by two CTA barriers per tile.
// Here ONE CTA of 1024 threads per SM runs the two halves on different warps, coupled only
// by mbarriers over a ring of staged tiles:
//     warp 31          loader: deals tiles (static share, then tickets), bulk-copies the
//                      tile record and its prefix block into a free stage
//     warps 0 .. 7     front end: warp w compacts the hits of the 2048-position chunk w of
//                      every staged tile into the stage's hit lists (its list offsets come
//                      from the count phase's prefix block, so the warps never talk)
//     warps 8 .. 30    body: claim 32-candidate slices of a finished stage from a shared
//                      counter (no imbalance inside the SM), one thread per candidate
// The lane tables are staged once per SM instead of once per CTA.
#pragma once
#include "scan.cuh"

static constexpr int kWsThreads = 1024;
static constexpr int kWsWarps = kWsThreads / 32;
#ifndef CRP_WS_FE_GROUPS
#define CRP_WS_FE_GROUPS 2
#endif
static constexpr int kWsFeGroups = CRP_WS_FE_GROUPS;          // front-end groups; group g takes the staged tiles n = g (mod groups)
static constexpr int kWsFeWarps = kWarps * kWsFeGroups;       // a group has one warp per 2048-position chunk of a tile
static constexpr int kWsLoaderWarp = kWsWarps - 1;
static constexpr int kWsBodyWarps = kWsWarps - kWsFeWarps - 1;
#ifndef CRP_WS_STAGES
#define CRP_WS_STAGES 8
#endif
static constexpr int kWsStages = CRP_WS_STAGES;               // staged tiles per SM
static constexpr int kWsListCap = 2048;                       // hits per strand of a tile that go through the lists
static constexpr int kWsTeams = kWsWarps / kWarps;            // count phase: teams of 8 warps, one tile each
static constexpr int kWsMaxRange = kWsThreads / kWarps;       // tiles per CTA per wave: one (tile, warp) count per thread
static constexpr size_t kWsRecStride = (kRecBytes + 127) / 128 * 128;
static constexpr size_t kWsListBytes = 2 * (size_t)kWsListCap * sizeof(uint16_t);
static constexpr size_t kWsStageBytes = kWsRecStride + kWsListBytes;
static constexpr size_t kWsEmitBytes = kWsStages * kWsStageBytes;
static constexpr size_t kWsCountBytes = (size_t)kWsTeams * kCountStages * kPamBytes;
static constexpr size_t kWsSmem = kWsEmitBytes > kWsCountBytes ? kWsEmitBytes : kWsCountBytes;
static constexpr int kWsMaxGrid = 256;                        // range prefixes kept in shared memory

struct __align__(16) WsRing {
    unsigned long long pref[kWsStages][kPrefWords];   // bulk-copy destination: prefix block of the staged tile
    unsigned long long rec_full[kWsStages];           // record + prefix block have landed (loader -> front end)
    unsigned long long list_full[kWsStages];          // the hit lists are complete (front end -> body)
    unsigned long long empty[kWsStages];              // every body warp is done with the stage (body -> loader)
    unsigned long long rbase[kWsStages];              // global prefix of the count range of the staged tile
    uint32_t tile[kWsStages];
    uint32_t claim[kWsStages];                        // next unclaimed slice of the stage
    uint32_t done[kWsStages];                         // body warps finished with the stage
    // count phase: one ring of PAM records per team
    unsigned long long cfull[kWsTeams][kCountStages];
    uint32_t ctile[kWsTeams][kCountStages];
    uint32_t cdone[kWsTeams][kCountStages];
};

// one candidate: window out of the staged record, score, the three stores (CROPSR.py:418-434, 285-313)
template <bool kScore, bool kMinus>
__device__ __forceinline__ void ws_emit_candidate(const ScanArgs &a, const double *__restrict__ tab,
                                                  const uint4 *__restrict__ rec, uint32_t ws, uint32_t row,
                                                  uint32_t t_start, uint32_t L) {
    const uint32_t t = t_start - (kMinus ? kWinBiasMinus : kWinBiasPlus) + ws;
    __stcs((kMinus ? a.pos_minus : a.pos_plus) + row, t);
    if (kScore) {
        const Window w = extract_window<kMinus>(rec, ws, t, L);
        const double x = rs1_canonical(tab, w.s0, w.s1, w.valid);
        __stcs((kMinus ? a.packed_minus : a.packed_plus) + row, w.packed);
        __stcs((kMinus ? a.x_minus : a.x_plus) + row, x);
    }
}

// dense tile (more hits than the lists hold): a lane emits the hits of its own word, ranks
// [rank, ...); rows beyond the clipped count are dropped
template <bool kScore, bool kMinus>
__device__ __forceinline__ void ws_emit_word(const ScanArgs &a, const double *__restrict__ tab,
                                             const uint4 *__restrict__ rec, uint32_t m, uint32_t rank, uint32_t count,
                                             uint32_t row0, uint32_t pos0, uint32_t t_start, uint32_t L) {
    while (m) {
        const uint32_t lb = m & (0u - m);
        if (rank < count)
            ws_emit_candidate<kScore, kMinus>(a, tab, rec, pos0 + 31u - (uint32_t)__clz(lb) + (kMinus ? kWinBiasMinus : kWinBiasPlus),
                                              row0 + rank, t_start, L);
        m ^= lb;
        ++rank;
    }
}

#ifdef CRP_WS_WAITSTATS
#define WS_WAIT(slot_, cond_lane_, call_) do { long long t0_ = clock64(); call_; if ((cond_lane_) && g_dbg_times) g_dbg_times[8ull * blockIdx.x + (slot_)] += (unsigned long long)(clock64() - t0_); } while (0)
#else
#define WS_WAIT(slot_, cond_lane_, call_) call_
#endif
template <bool kScore>
__global__ void __launch_bounds__(kWsThreads, 1)
k_scan_ws(const ScanArgs a) {
    extern __shared__ __align__(128) unsigned char s_dyn[];
    __shared__ __align__(16) double s_tab[kScore ? RS1_TABLE_DOUBLES : 1];
    __shared__ WsRing ring;
    __shared__ uint32_t s_cnt[kWsMaxRange][kWarps];
    __shared__ unsigned long long s_scan[kWsWarps];
    __shared__ unsigned long long s_rangepref[kWsMaxGrid];
    __shared__ __align__(8) unsigned long long s_tabbar;

    cg::grid_group grid = cg::this_grid();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int l = a.guide_len;
    const uint32_t G = gridDim.x, cta = blockIdx.x;
    auto stage_rec = [&](int s) { return reinterpret_cast<uint4 *>(s_dyn + (size_t)s * kWsStageBytes); };
    auto stage_list = [&](int s) { return reinterpret_cast<uint16_t *>(s_dyn + (size_t)s * kWsStageBytes + kWsRecStride); };

    dbg_stamp(0);
    if (tid == 0) {
        mbar_init(&s_tabbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (kScore) {
            mbar_expect(&s_tabbar, (uint32_t)kRs1TableBytes);
            bulk_copy(s_tab, a.tables, (uint32_t)kRs1TableBytes, &s_tabbar);
        }
    }
    // (re)arm every mbarrier of the ring; all threads call it
    auto ring_reset = [&](bool first) {
        __syncthreads();
        if (tid == 0) {
            for (int s = 0; s < kWsStages; ++s) {
                if (!first) {
                    mbar_inval(&ring.rec_full[s]);
                    mbar_inval(&ring.list_full[s]);
                    mbar_inval(&ring.empty[s]);
                }
                mbar_init(&ring.rec_full[s], 1);
                mbar_init(&ring.list_full[s], kWarps);
                mbar_init(&ring.empty[s], 1);
                ring.claim[s] = 0;
                ring.done[s] = 0;
            }
            for (int q = 0; q < kWsTeams; ++q)
                for (int s = 0; s < kCountStages; ++s) {
                    if (!first) mbar_inval(&ring.cfull[q][s]);
                    mbar_init(&ring.cfull[q][s], 1);
                    ring.cdone[q][s] = 0;
                }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    };

    unsigned long long wave_base = 0;
    uint32_t wave = 0;
    for (uint32_t w_lo = 0; w_lo < a.n_tiles; w_lo += a.wave_tiles, ++wave) {
        const uint32_t w_hi = min(a.n_tiles, w_lo + a.wave_tiles);
        const uint32_t nt = w_hi - w_lo;
        const uint32_t k = (nt + G - 1) / G;                       // tiles per count range (<= kWsMaxRange)

        // ================================================= count phase: tiles [r_lo, r_lo + n_mine)
        // team q (8 warps) counts the tiles r_lo + q, r_lo + q + kWsTeams, ...; warp wq of the team
        // counts chunk wq of each of them
        const uint32_t r_lo = min(w_hi, w_lo + cta * k), n_mine = min(w_hi, r_lo + k) - r_lo;
        const int q = warp / kWarps, wq = warp % kWarps;
        const uint32_t n_team = n_mine > (uint32_t)q ? (n_mine - (uint32_t)q + kWsTeams - 1) / kWsTeams : 0u;
        auto pam_stage = [&](int s) { return s_dyn + ((size_t)q * kCountStages + s) * kPamBytes; };
        auto produce_count = [&](uint32_t m, int s) {
            if (m < n_team) {
                const uint32_t t = r_lo + (uint32_t)q + m * kWsTeams;
                ring.ctile[q][s] = t;
                mbar_expect(&ring.cfull[q][s], kPamBytes);
                bulk_copy(pam_stage(s), a.pam + (size_t)t * kPamBytes, kPamBytes, &ring.cfull[q][s]);
            } else {
                ring.ctile[q][s] = kNoTile;
                mbar_arrive(&ring.cfull[q][s]);
            }
        };
        if (cta == 0 && tid == 0) a.tickets[wave] = 0;             // read after the grid barrier
        ring_reset(wave == 0);
#ifndef CRP_WS_WAITSTATS
        dbg_stamp(1);
#endif
        if (wq == 0 && lane == 0)
            for (int s = 0; s < kCountStages; ++s) produce_count(s, s);
        for (uint32_t m = 0;; ++m) {
            const int s = m % kCountStages;
            mbar_wait(&ring.cfull[q][s], (m / kCountStages) & 1u);
            if (ring.ctile[q][s] == kNoTile) break;
            const uint4 d = *reinterpret_cast<const uint4 *>(pam_stage(s));
            const TileDesc td = {d.x, d.y, d.z, d.w};
            const Hits h = tile_hits_pam(pam_stage(s), td, l, 64 * wq + lane);
            uint32_t c = (__popc(h.pA) + __popc(h.pB)) | ((__popc(h.mA) + __popc(h.mB)) << 16);
            c = __reduce_add_sync(0xFFFFFFFFu, c);
            if (lane == 0) s_cnt[(uint32_t)q + m * kWsTeams][wq] = c;
            __syncwarp();
            if (lane == 0 && atomicAdd(&ring.cdone[q][s], 1u) == (uint32_t)kWarps - 1u) {
                ring.cdone[q][s] = 0;
                produce_count(m + kCountStages, s);
            }
        }
        __syncthreads();
        {   // exclusive scan over the (tile, warp) counts of the range: thread tid owns tile tid / 8, warp tid % 8
            const uint32_t j = tid / kWarps, cw = tid % kWarps;
            const unsigned long long mine = j < n_mine ? unpack_counts(s_cnt[j][cw]) : 0ull;
            unsigned long long incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += v;
            }
            if (lane == 31) s_scan[warp] = incl;
            __syncthreads();
            unsigned long long before = 0, total = 0;
            for (int w = 0; w < kWsWarps; ++w) {
                const unsigned long long x = s_scan[w];
                if (w < warp) before += x;
                total += x;
            }
            if (j < n_mine) {
                unsigned long long *pf = a.warp_pref + (size_t)(r_lo + j) * kPrefWords;
                pf[cw] = before + incl - mine;
                if (cw == kWarps - 1) pf[kWarps] = before + incl;      // prefix at the end of the tile
                asm volatile("fence.proxy.async.global;" ::: "memory");   // read back by bulk copies after the grid barrier
            }
            if (tid == 0) a.cta_tot[(wave & 1u) * G + cta] = total;
        }
        dbg_stamp(2);
        grid.sync();
        dbg_stamp(3);

        // ================================================= exclusive scan of the range totals
        unsigned long long wave_total;
        {
            const unsigned long long mine = (uint32_t)tid < G ? a.cta_tot[(wave & 1u) * G + tid] : 0ull;
            unsigned long long incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long x = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += x;
            }
            if (lane == 31) s_scan[warp] = incl;
            __syncthreads();
            unsigned long long before = 0, total = 0;
            for (int w = 0; w < kWsMaxGrid / 32; ++w) {       // ranges live in the first kWsMaxGrid threads
                const unsigned long long x = s_scan[w];
                if (w < warp) before += x;
                total += x;
            }
            if ((uint32_t)tid < G) s_rangepref[tid] = wave_base + before + incl - mine;
            wave_total = total;
        }
        const uint32_t ns = (uint32_t)((unsigned long long)nt * a.static_eighths / 8 / G);
        const uint32_t dyn_lo = w_lo + ns * G, n_dyn = w_hi - dyn_lo;
        unsigned int *const ticket = a.tickets + wave;
        ring_reset(false);                                         // also publishes s_rangepref
        dbg_stamp(4);
        // per-segment candidate counts of this wave (as in k_scan_score)
        for (uint32_t sg = cta * kWsThreads + tid; sg < a.n_seg; sg += G * kWsThreads) {
            const uint32_t f = a.seg_first_tile ? a.seg_first_tile[sg] : 0u;
            const uint32_t c = a.seg_tile_count ? a.seg_tile_count[sg] : a.n_tiles;
            const uint32_t lo_t = max(f, w_lo), hi_t = min(f + c, w_hi);
            unsigned long long cnt = 0;
            if (lo_t < hi_t)
                cnt = s_rangepref[(hi_t - 1 - w_lo) / k] + a.warp_pref[(size_t)(hi_t - 1) * kPrefWords + kWarps] -
                      (s_rangepref[(lo_t - w_lo) / k] + a.warp_pref[(size_t)lo_t * kPrefWords]);
            const unsigned long long plus = cnt >> 32, minus = cnt & 0xFFFFFFFFull;
            a.seg_counts[sg] = (wave ? a.seg_counts[sg] : 0ull) + plus;
            a.seg_counts[a.n_seg + sg] = (wave ? a.seg_counts[a.n_seg + sg] : 0ull) + minus;
        }

        // ================================================= emit phase, by role
        if (warp == kWsLoaderWarp) {
            if (lane == 0) {
                // the ticket of a dynamic tile is drawn one tile ahead, so that its round trip to L2
                // overlaps the wait for a free stage
                uint32_t tk_next = ns == 0 ? atomicAdd(ticket, 1u) : 0u;
                for (uint32_t n = 0;; ++n) {
                    const int s = n % kWsStages;
                    uint32_t t;
                    if (n < ns) {
                        t = w_lo + n * G + cta;
                    } else {
                        t = tk_next < n_dyn ? dyn_lo + tk_next : kNoTile;
                    }
                    if (n + 1 >= ns && t != kNoTile) tk_next = atomicAdd(ticket, 1u);
                    if (n >= (uint32_t)kWsStages) WS_WAIT(6, true, mbar_wait(&ring.empty[s], (n / kWsStages - 1u) & 1u));
                    ring.tile[s] = t;
                    if (t == kNoTile) {                            // end of the sequence: one end marker per front-end group
                        mbar_arrive(&ring.rec_full[s]);
                        for (uint32_t e = 1; e < (uint32_t)kWsFeGroups; ++e) {
                            const int se = (n + e) % kWsStages;
                            if (n + e >= (uint32_t)kWsStages) mbar_wait(&ring.empty[se], ((n + e) / kWsStages - 1u) & 1u);
                            ring.tile[se] = kNoTile;
                            mbar_arrive(&ring.rec_full[se]);
                        }
                        break;
                    }
                    ring.rbase[s] = s_rangepref[(t - w_lo) / k];
                    mbar_expect(&ring.rec_full[s], kRecBytes + kPrefWords * 8);
                    bulk_copy(stage_rec(s), a.records + (size_t)t * kRecWords, kRecBytes, &ring.rec_full[s]);
                    bulk_copy(ring.pref[s], a.warp_pref + (size_t)t * kPrefWords, kPrefWords * 8, &ring.rec_full[s]);
                }
            }
        } else if (warp < kWsFeWarps) {
            const int chunk = warp % kWarps;
            for (uint32_t n = warp / kWarps;; n += kWsFeGroups) {
                const int s = n % kWsStages;
                WS_WAIT(1, warp == 0 && lane == 0, mbar_wait(&ring.rec_full[s], (n / kWsStages) & 1u));
                if (ring.tile[s] == kNoTile) {
                    if (lane == 0) mbar_arrive(&ring.list_full[s]);
                    break;
                }
                const uint4 *rec = stage_rec(s);
                const unsigned long long tile_pref = ring.pref[s][0];
                const unsigned long long tot = ring.pref[s][kWarps] - tile_pref;
                const uint32_t np = (uint32_t)(tot >> 32), nm = (uint32_t)tot;
                if (np <= (uint32_t)kWsListCap && nm <= (uint32_t)kWsListCap) {
                    const uint4 d = rec[0];
                    const TileDesc td = {d.x, d.y, d.z, d.w};
                    const unsigned long long off = ring.pref[s][chunk] - tile_pref;   // hits of the tile before my chunk
                    const uint32_t wordA = 64 * chunk + lane;
                    const Hits h = tile_hits(rec, td, l, wordA);
                    // warp scan of the per-word counts: the A words of the chunk precede its B words
                    const uint32_t cA = __popc(h.pA) | (__popc(h.mA) << 16), cB = __popc(h.pB) | (__popc(h.mB) << 16);
                    uint32_t iA = cA, iB = cB;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t vA = __shfl_up_sync(0xFFFFFFFFu, iA, o), vB = __shfl_up_sync(0xFFFFFFFFu, iB, o);
                        if (lane >= o) {
                            iA += vA;
                            iB += vB;
                        }
                    }
                    const uint32_t totA = __shfl_sync(0xFFFFFFFFu, iA, 31);
                    const uint32_t xA = iA - cA, xB = totA + iB - cB;
                    const uint32_t op = (uint32_t)(off >> 32), om = (uint32_t)off;
                    uint16_t *const list_p = stage_list(s), *const list_m = list_p + kWsListCap;
                    list_hits(list_p + op + (xA & 0xFFFFu), h.pA, 32u * wordA + kWinBiasPlus);
                    list_hits(list_p + op + (xB & 0xFFFFu), h.pB, 32u * (wordA + 32) + kWinBiasPlus);
                    list_hits(list_m + om + (xA >> 16), h.mA, 32u * wordA + kWinBiasMinus);
                    list_hits(list_m + om + (xB >> 16), h.mB, 32u * (wordA + 32) + kWinBiasMinus);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&ring.list_full[s]);
            }
        } else {
            if (kScore && wave == 0) mbar_wait(&s_tabbar, 0);
            const uint32_t cap = (uint32_t)a.capacity;
            for (uint32_t n = 0;; ++n) {
                const int s = n % kWsStages;
                WS_WAIT(7, warp == kWsFeWarps && lane == 0, mbar_wait(&ring.list_full[s], (n / kWsStages) & 1u));
                mbar_wait(&ring.rec_full[s], (n / kWsStages) & 1u);   // completed long ago: orders the bulk-copied bytes for this warp too
                if (ring.tile[s] == kNoTile) break;
                const uint4 *rec = stage_rec(s);
                const uint4 d = rec[0];
                const uint32_t t_start = d.x, L = d.y;
                const unsigned long long tile_pref = ring.pref[s][0];
                const unsigned long long base = ring.rbase[s] + tile_pref;
                const unsigned long long tot = ring.pref[s][kWarps] - tile_pref;
                const uint32_t np = (uint32_t)(tot >> 32), nm = (uint32_t)tot;
                const uint32_t base_p = (uint32_t)(base >> 32), base_m = (uint32_t)base;
                // rows of this tile that fit the streams
                const uint32_t cp = base_p < cap ? min(np, cap - base_p) : 0u, cm = base_m < cap ? min(nm, cap - base_m) : 0u;
                if (np <= (uint32_t)kWsListCap && nm <= (uint32_t)kWsListCap) {
                    const uint16_t *const list_p = stage_list(s), *const list_m = list_p + kWsListCap;
                    const uint32_t sp = (cp + 31u) / 32u, sl = sp + (cm + 31u) / 32u;
                    for (;;) {
                        uint32_t c = 0;
                        if (lane == 0) c = atomicAdd(&ring.claim[s], 1u);
                        c = __shfl_sync(0xFFFFFFFFu, c, 0);
                        if (c >= sl) break;
                        if (c < sp) {
                            const uint32_t i = 32u * c + lane;
                            if (i < cp) ws_emit_candidate<kScore, false>(a, s_tab, rec, list_p[i], base_p + i, t_start, L);
                        } else {
                            const uint32_t i = 32u * (c - sp) + lane;
                            if (i < cm) ws_emit_candidate<kScore, true>(a, s_tab, rec, list_m[i], base_m + i, t_start, L);
                        }
                    }
                } else {                                           // dense tile: a slice is one warp chunk, straight from the record
                    const TileDesc td = {d.x, d.y, d.z, d.w};
                    for (;;) {
                        uint32_t c = 0;
                        if (lane == 0) c = atomicAdd(&ring.claim[s], 1u);
                        c = __shfl_sync(0xFFFFFFFFu, c, 0);
                        if (c >= (uint32_t)kWarps) break;
                        const unsigned long long off = ring.pref[s][c] - tile_pref;
                        const uint32_t wordA = 64 * c + lane;
                        const Hits h = tile_hits(rec, td, l, wordA);
                        const uint32_t cA = __popc(h.pA) | (__popc(h.mA) << 16), cB = __popc(h.pB) | (__popc(h.mB) << 16);
                        uint32_t iA = cA, iB = cB;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const uint32_t vA = __shfl_up_sync(0xFFFFFFFFu, iA, o), vB = __shfl_up_sync(0xFFFFFFFFu, iB, o);
                            if (lane >= o) {
                                iA += vA;
                                iB += vB;
                            }
                        }
                        const uint32_t totA = __shfl_sync(0xFFFFFFFFu, iA, 31);
                        const uint32_t xA = iA - cA, xB = totA + iB - cB;
                        const uint32_t op = (uint32_t)(off >> 32), om = (uint32_t)off;
                        ws_emit_word<kScore, false>(a, s_tab, rec, h.pA, op + (xA & 0xFFFFu), cp, base_p, 32u * wordA, t_start, L);
                        ws_emit_word<kScore, false>(a, s_tab, rec, h.pB, op + (xB & 0xFFFFu), cp, base_p, 32u * (wordA + 32), t_start, L);
                        ws_emit_word<kScore, true>(a, s_tab, rec, h.mA, om + (xA >> 16), cm, base_m, 32u * wordA, t_start, L);
                        ws_emit_word<kScore, true>(a, s_tab, rec, h.mB, om + (xB >> 16), cm, base_m, 32u * (wordA + 32), t_start, L);
                    }
                }
                __syncwarp();
                if (lane == 0 && atomicAdd(&ring.done[s], 1u) == (uint32_t)kWsBodyWarps - 1u) {
                    ring.done[s] = 0;                              // last body warp out: the stage is free again
                    ring.claim[s] = 0;
                    mbar_arrive(&ring.empty[s]);
                }
            }
        }
        __syncthreads();
        dbg_stamp(5);
        wave_base += wave_total;
    }
}
