// k_scan_lists: the scan + score kernel with the hit COMPACTION moved into the count phase.
//
// k_scan_score (scan.cuh) finds every PAM hit twice: the count phase counts them (one warp per tile,
// bandwidth bound, integer pipes idle), and the emit phase finds them again -- eight warps per tile
// recompute the hit masks from the full record, rank them with a warp scan and pop them into the
// CTA's hit lists -- before one thread per candidate can start.  That front end is a third of the
// emit phase's instructions, on the pipe that bounds the kernel (the integer ALU).
//
// Here the warp that counts a tile also BUILDS its hit lists, while it has the PAM planes in shared
// memory anyway: a second sweep over the 16 rows of 32 words ranks the hits (one warp scan per row
// instead of two per 64-word chunk and warp) and pops them into a per-warp list buffer, which a bulk
// store (TMA, shared -> global) writes to the tile's 4 KB slot of a list array in HBM.  ~1,100 warp
// instructions per tile instead of ~2,400, executed where the ALU has nothing else to do.  The emit
// phase then receives the lists with the record and the prefix block (one more bulk copy per tile),
// every thread takes its <= 4 + 4 entries into registers, and after ONE barrier -- which frees the
// list buffer for the next tile's copy -- the candidate bodies run.  A tile with more than kListCap
// hits on a strand (poly-G / poly-C) is flagged in its prefix block and goes through the old in-tile
// front end in windows of kListCap ranks.
//
// Shared memory is one dynamic block with two layouts:
//   count phase   [PAM ring: kCountStages2 slots][per-warp list buffers: 8 x 4 KB]
//   emit phase    [2 tile records][hit lists 4 KB][range prefixes][Rule-Set-1 lane tables]
// (the lane tables are fetched when the CTA's count phase is through, behind the grid barrier).
#pragma once
#include "scan.cuh"

static constexpr int kCountStages2 = 5;
static constexpr uint32_t kListBytes = 2u * kListCap * (uint32_t)sizeof(uint16_t);      // '+' list, then '-' list
static constexpr uint32_t kPamRingBytes = ((uint32_t)kCountStages2 * kPamBytes + 127u) / 128u * 128u;
static constexpr uint32_t kCountSmem2 = kPamRingBytes + (uint32_t)kWarps * kListBytes;
// emit layout offsets
static constexpr uint32_t kOffLists = (uint32_t)kStages * kRecBytes;
static constexpr uint32_t kOffRange = kOffLists + kListBytes;
__host__ __device__ inline uint32_t off_tables2(uint32_t grid) { return (kOffRange + grid * 8u + 15u) / 16u * 16u; }
__host__ __device__ inline uint32_t smem_bytes2(uint32_t grid) {
    const uint32_t emit = off_tables2(grid) + (uint32_t)kRs1TableBytes;
    return emit > kCountSmem2 ? emit : kCountSmem2;
}
static_assert(kCountStages2 <= kCountStages, "Ring has kCountStages mbarriers");

struct ListArgs {
    ScanArgs s;
    unsigned char *lists;            // [n_tiles][kListBytes]: '+' positions (u16, biased like the in-tile lists), then '-'
};

__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes)
                 : "memory");
}

// One warp, one staged PAM record: the hit lists of the tile into `wl` ('+' at wl, '-' at wl + kListCap).
// Lane l owns the words 32 i + l of row i; ranks run row by row, lane by lane, bit by bit: ascending positions.
__device__ __forceinline__ void warp_list_tile(const unsigned char *__restrict__ rec, int l, int lane, uint16_t *__restrict__ wl) {
    const uint4 d = *reinterpret_cast<const uint4 *>(rec);
    const uint2 *w = reinterpret_cast<const uint2 *>(rec + 16);
    const int32_t t0 = (int32_t)d.x, L = (int32_t)d.y;
    const int32_t last_owned = t0 + (int32_t)d.z - 1;
    const int32_t hi_p = min(L - 3, last_owned);
    const int32_t hi_m = min(L - l + 7, hi_p);
    const bool edge = t0 < l + 5 || t0 + kTile - 1 > hi_m;         // warp-uniform
    uint32_t run = 0;                                              // (plus | minus << 16) hits of the rows before
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
        const int word = 32 * i + lane;
        const uint2 a = w[word], an = w[word + 1];
        uint32_t p = __funnelshift_r(a.x, an.x, 1) & __funnelshift_r(a.x, an.x, 2);
        uint32_t m = a.y & __funnelshift_r(a.y, an.y, 1);
        if (edge) {
            const int32_t tw = t0 + 32 * word;
            p &= range_mask(tw, l + 5, hi_p);
            m &= range_mask(tw, 2, hi_m);
        }
        if (!__any_sync(0xFFFFFFFFu, (p | m) != 0u)) continue;     // a soft-masked / N row: nothing to rank
        const uint32_t c = __popc(p) | (__popc(m) << 16);
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += v;
        }
        const uint32_t excl = run + incl - c;
        run += __shfl_sync(0xFFFFFFFFu, incl, 31);
        list_hits(wl + (excl & 0xFFFFu), p, 32u * (uint32_t)word + kWinBiasPlus);
        list_hits(wl + kListCap + (excl >> 16), m, 32u * (uint32_t)word + kWinBiasMinus);
    }
}

// one thread per listed hit, list entries already in registers (e[j] = entry slot + 256 j)
template <bool kScore, bool kMinus>
__device__ __forceinline__ void emit_strand_reg(const ScanArgs &a, const double *__restrict__ tab, const uint4 *__restrict__ rec,
                                                const uint32_t (&e)[4], uint32_t count, uint32_t row0, uint32_t t_start,
                                                uint32_t L, uint32_t slot) {
    const uint32_t cap = (uint32_t)a.capacity;
    if (row0 >= cap) return;
    if (count > cap - row0) count = cap - row0;
    uint32_t *const pos = kMinus ? a.pos_minus : a.pos_plus;
    unsigned long long *const packed = kMinus ? a.packed_minus : a.packed_plus;
    double *const xs = kMinus ? a.x_minus : a.x_plus;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t i = slot + (uint32_t)kThreads * j;
        if (i >= count) break;
        const uint32_t ws = e[j], t = t_start - (kMinus ? kWinBiasMinus : kWinBiasPlus) + ws;
        const uint32_t row = row0 + i;
        CRP_CHECK(a, ws >= (kMinus ? kWinBiasMinus : kWinBiasPlus) && 2u + (ws >> 5) < (uint32_t)kRecWords, 2);
        CRP_CHECK(a, row < cap && t < L, 3);
        __stcs(pos + row, t);
        if (kScore) {
            const Window w = extract_window<kMinus>(rec, ws, t, L);
            const double x = rs1_canonical(tab, w.s0, w.s1, w.valid);
            __stcs(packed + row, w.packed);
            __stcs(xs + row, x);
        }
    }
}

template <bool kScore>
__global__ void __launch_bounds__(kThreads, CRP_CTAS_PER_SM)
k_scan_lists(const ListArgs la) {
    const ScanArgs &a = la.s;
    extern __shared__ __align__(128) unsigned char s_dyn[];
    const uint32_t G = gridDim.x, cta = blockIdx.x;
    auto stage = [&](int s) { return reinterpret_cast<uint4 *>(s_dyn + (size_t)s * kRecBytes); };
    uint16_t *const s_list = reinterpret_cast<uint16_t *>(s_dyn + kOffLists);
    unsigned long long *const s_rangepref = reinterpret_cast<unsigned long long *>(s_dyn + kOffRange);
    const double *const s_tab = reinterpret_cast<const double *>(s_dyn + off_tables2(G));
    __shared__ Ring ring;
    __shared__ uint32_t s_cnt[kMaxRange][kWarps];
    __shared__ unsigned long long s_scan[kWarps];
    __shared__ uint32_t s_dense[kMaxRange];
    __shared__ __align__(8) unsigned long long s_tabbar;

    cg::grid_group grid = cg::this_grid();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int l = a.guide_len;

    dbg_stamp(0);
    if (tid == 0) {
        mbar_init(&s_tabbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    auto record = [&](uint32_t tile) { return a.records + (size_t)tile * kRecWords; };
    const uint32_t nt = a.n_tiles;
    const uint32_t k = (nt + G - 1) / G;                           // tiles per count range

    // ================================================= count + list phase: tiles [r_lo, r_lo + n_mine)
    const uint32_t r_lo = min(nt, cta * k), n_mine = min(nt, r_lo + k) - r_lo;
    auto pam_stage = [&](int s) { return s_dyn + (size_t)s * kPamBytes; };
    uint16_t *const my_list = reinterpret_cast<uint16_t *>(s_dyn + kPamRingBytes + (size_t)warp * kListBytes);
    auto produce_count = [&](uint32_t n) {            // n < n_mine
        const int s = n % kCountStages2;
        *reinterpret_cast<volatile uint32_t *>(&ring.tile[s]) = n;
        mbar_expect(&ring.full[s], kPamBytes);
        bulk_copy(pam_stage(s), a.pam + (size_t)(r_lo + n) * kPamBytes, kPamBytes, &ring.full[s]);
    };
    if (cta == 0 && tid == 0) {                                    // read after the grid barrier
        a.tickets[0] = 0;
        a.tickets[1] = 0;
    }
    ring_reset(ring, true, true);
    dbg_stamp(1);
    if (tid == 0)
        for (uint32_t n = 0; n < (uint32_t)kCountStages2 && n < n_mine; ++n) produce_count(n);
    unsigned long long range_run = 0;
    bool store_in_flight = false;                                  // lane 0: a bulk store of my list buffer may still be reading it
    for (uint32_t b_lo = 0;; b_lo += kMaxRange) {
        const uint32_t b_n = min(n_mine - min(n_mine, b_lo), (uint32_t)kMaxRange);
        for (uint32_t n = b_lo + warp; n < b_lo + b_n; n += kWarps) {
            const int s = n % kCountStages2;
            while (*reinterpret_cast<volatile uint32_t *>(&ring.tile[s]) != n) __nanosleep(32);
            mbar_wait(&ring.full[s], (n / kCountStages2) & 1u);
            uint32_t *cnt = s_cnt[n - b_lo];
            warp_count_tile(pam_stage(s), l, lane, cnt);
            __syncwarp();
            uint32_t tot = 0;
#pragma unroll
            for (int c = 0; c < kWarps; ++c) tot += cnt[c];
            const uint32_t np = tot & 0xFFFFu, nm = tot >> 16;     // a tile holds <= 16,384 hits per strand: no carry between the halves
            const bool dense = np > (uint32_t)kListCap || nm > (uint32_t)kListCap;
            if (lane == 0) s_dense[n - b_lo] = dense ? 1u : 0u;
            if (!dense && tot) {
                if (lane == 0 && store_in_flight) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
                warp_list_tile(pam_stage(s), l, lane, my_list);
                __syncwarp();
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    unsigned char *slot = la.lists + (size_t)(r_lo + n) * kListBytes;
                    if (np) bulk_store(slot, my_list, (np * 2u + 15u) & ~15u);
                    if (nm) bulk_store(slot + kListCap * sizeof(uint16_t), my_list + kListCap, (nm * 2u + 15u) & ~15u);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    store_in_flight = true;
                }
            }
            __syncwarp();
            if (lane == 0 && n + kCountStages2 < n_mine) produce_count(n + kCountStages2);
        }
        __syncthreads();
        {   // exclusive scan over the (tile, warp) counts of the batch: thread tid owns tile tid / 8, chunk tid % 8
            const uint32_t j = tid / kWarps, wq = tid % kWarps;
            const unsigned long long mine = j < b_n ? unpack_counts(s_cnt[j][wq]) : 0ull;
            unsigned long long incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += v;
            }
            if (lane == 31) s_scan[warp] = incl;
            __syncthreads();
            unsigned long long before = range_run, total = 0;
#pragma unroll
            for (int q = 0; q < kWarps; ++q) {
                const unsigned long long x = s_scan[q];
                if (q < warp) before += x;
                total += x;
            }
            if (j < b_n) {
                unsigned long long *pf = a.warp_pref + (size_t)(r_lo + b_lo + j) * kPrefWords;
                pf[wq] = before + incl - mine;
                if (wq == kWarps - 1) {
                    pf[kWarps] = before + incl;                    // prefix at the end of the tile
                    pf[kWarps + 1] = s_dense[j];                   // 1: no lists in HBM, the emit phase compacts this tile itself
                }
                asm volatile("fence.proxy.async.global;" ::: "memory");
            }
            range_run += total;
            __syncthreads();
        }
        if (b_lo + kMaxRange >= n_mine) break;
    }
    if (lane == 0 && store_in_flight) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // my lists are in HBM
    __threadfence();
    __syncthreads();                                               // every warp is through with the count layout of shared memory
    if (tid == 0) {
        a.cta_tot[cta] = range_run;
        if (kScore) {                                              // the lane tables land behind the grid barrier
            mbar_expect(&s_tabbar, (uint32_t)kRs1TableBytes);
            bulk_copy(const_cast<double *>(s_tab), a.tables, (uint32_t)kRs1TableBytes, &s_tabbar);
        }
    }
    // first emit tile of this CTA (static share): record fetched across the grid barrier
    const uint32_t ns = (uint32_t)((unsigned long long)nt * a.static_eighths / 8 / G);
    const bool pre = ns > 0;
    const uint32_t t_pre = cta;
    if (pre && tid == 0) {
        mbar_expect(&ring.pre, kRecBytes + kPrefWords * 8 + kListBytes);
        bulk_copy(stage(0), record(t_pre), kRecBytes, &ring.pre);
    }
    dbg_stamp(2);
    grid.sync();
    dbg_stamp(3);
    if (pre && tid == 0) {
        bulk_copy(ring.pref[0], a.warp_pref + (size_t)t_pre * kPrefWords, kPrefWords * 8, &ring.pre);
        bulk_copy(s_list, la.lists + (size_t)t_pre * kListBytes, kListBytes, &ring.pre);
    }

    // ================================================= exclusive scan of the range totals
    {
        unsigned long long v[4] = {0, 0, 0, 0}, mine = 0;     // thread owns ranges 4*tid .. 4*tid+3
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t i = 4u * tid + q;
            if (i < G) v[q] = a.cta_tot[i];
            mine += v[q];
        }
        unsigned long long incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long x = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += x;
        }
        if (lane == 31) s_scan[warp] = incl;
        __syncthreads();
        unsigned long long before = 0;
#pragma unroll
        for (int q = 0; q < kWarps; ++q) {
            const unsigned long long x = s_scan[q];
            if (q < warp) before += x;
        }
        unsigned long long run = before + incl - mine;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t i = 4u * tid + q;
            if (i < G) s_rangepref[i] = run;
            run += v[q];
        }
    }

    // ================================================= emit phase
    const uint32_t dyn_lo = ns * G, n_dyn = nt - dyn_lo;
    unsigned int *const ticket = a.tickets;
    // one thread: stage tile number n of this CTA into slot s -- record and prefix block now, the lists when
    // the list buffer is free (stage_lists)
    auto produce_emit = [&](uint32_t n, int s) {
        uint32_t t;
        if (n < ns) {
            t = n * G + cta;
        } else {
            const uint32_t q = atomicAdd(ticket, 1u);
            t = q < n_dyn ? dyn_lo + q : kNoTile;
        }
        ring.tile[s] = t;
        if (t != kNoTile) {
            ring.rbase[s] = s_rangepref[t / k];
            mbar_expect(&ring.full[s], kRecBytes + kPrefWords * 8 + kListBytes);
            bulk_copy(stage(s), record(t), kRecBytes, &ring.full[s]);
            bulk_copy(ring.pref[s], a.warp_pref + (size_t)t * kPrefWords, kPrefWords * 8, &ring.full[s]);
        } else {
            mbar_arrive(&ring.full[s]);
        }
    };
    auto stage_lists = [&](int s) {                  // one thread, the list buffer is free
        const uint32_t t = ring.tile[s];
        if (t != kNoTile) bulk_copy(s_list, la.lists + (size_t)t * kListBytes, kListBytes, &ring.full[s]);
    };
    ring_reset(ring, false, false);                            // also publishes s_rangepref
    dbg_stamp(4);
    if (kScore) mbar_wait(&s_tabbar, 0);
    // per-segment candidate counts (and their exchange in a sharded scan), as in k_scan_score
    {
        const bool xchg = a.world > 1;
        for (uint32_t sg = cta * kThreads + tid; sg < a.seg_stride; sg += G * kThreads) {
            unsigned long long plus = 0, minus = 0;
            if (sg < a.n_seg) {
                const uint32_t f = a.seg_first_tile ? a.seg_first_tile[sg] : 0u;
                const uint32_t c = a.seg_tile_count ? a.seg_tile_count[sg] : a.n_tiles;
                if (c) {
                    const unsigned long long cnt = s_rangepref[(f + c - 1) / k] + a.warp_pref[(size_t)(f + c - 1) * kPrefWords + kWarps] -
                                                   (s_rangepref[f / k] + a.warp_pref[(size_t)f * kPrefWords]);
                    plus = cnt >> 32;
                    minus = cnt & 0xFFFFFFFFull;
                }
            }
            a.seg_counts[sg] = plus;
            a.seg_counts[a.seg_stride + sg] = minus;
            if (a.seg_counts_host) {
                a.seg_counts_host[sg] = plus;
                a.seg_counts_host[a.seg_stride + sg] = minus;
            }
            if (xchg) {
                const size_t at = (size_t)a.rank * 2 * a.seg_stride + sg;
                for (uint32_t p = 0; p < a.world; ++p) {
                    a.peer_gather[p][at] = plus;
                    a.peer_gather[p][at + a.seg_stride] = minus;
                }
            }
        }
        if (xchg && cta * kThreads < a.seg_stride) {
            __threadfence_system();
            __syncthreads();
            if (tid == 0) {
                const uint32_t writers = min(G, (a.seg_stride + kThreads - 1) / kThreads);
                __threadfence_system();
                if (atomicAdd(a.tickets + 1, 1u) == writers - 1) xchg_publish(a);
            }
        }
    }
    if (tid == 0) {
        if (pre) {
            mbar_arrive(&ring.full[0]);                        // tile 0 came through ring.pre: skip that phase of slot 0
        } else {
            produce_emit(0, 0);
            stage_lists(0);
        }
    }
    uint16_t *const list_p = s_list, *const list_m = s_list + kListCap;
    bool lists_owed = false;                                   // thread 0: the lists of the tile staged in the other slot are not requested yet
    for (uint32_t n = 0;; ++n) {
        const int s = n % kStages;
        if (tid == 0) {
            if (lists_owed) {                                  // (after a dense tile: its windows used the list buffer to the end)
                stage_lists(s);
                lists_owed = false;
            }
            produce_emit(n + 1, s ^ 1);
            lists_owed = true;
        }
        const bool first_pre = pre && n == 0;
        if (first_pre) mbar_wait(&ring.pre, 0u);
        else mbar_wait(&ring.full[s], (n / kStages) & 1u);
        if (!first_pre && ring.tile[s] == kNoTile) break;
        const uint4 *rec = stage(s);
        const uint4 d = rec[0];
        const TileDesc td = {d.x, d.y, d.z, d.w};
        const unsigned long long tile_pref = ring.pref[s][0];
        const unsigned long long base = (first_pre ? s_rangepref[t_pre / k] : ring.rbase[s]) + tile_pref;
        const unsigned long long tot = ring.pref[s][kWarps] - tile_pref;
        const bool dense = ring.pref[s][kWarps + 1] != 0ull;
        const uint32_t np = (uint32_t)(tot >> 32), nm = (uint32_t)tot;
        const uint32_t base_p = (uint32_t)(base >> 32), base_m = (uint32_t)base;
        if (!dense) {
            // my entries: '+' slot tid, '-' slot tid ^ 128 (the two strands start at opposite ends of the CTA)
            const uint32_t slot_p = (uint32_t)tid, slot_m = (uint32_t)tid ^ (kThreads / 2);
            uint32_t ep[4], em[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t ip = slot_p + (uint32_t)kThreads * j, im = slot_m + (uint32_t)kThreads * j;
                ep[j] = ip < np ? list_p[ip] : 0u;
                em[j] = im < nm ? list_m[im] : 0u;
            }
#ifdef CRP_CHECKED
            for (int j = 0; j < 4; ++j) {
                const uint32_t ip = slot_p + (uint32_t)kThreads * j, im = slot_m + (uint32_t)kThreads * j;
                CRP_CHECK(a, !(ip < np && ip > 0) || list_p[ip - 1] < list_p[ip], 4);
                CRP_CHECK(a, !(im < nm && im > 0) || list_m[im - 1] < list_m[im], 4);
            }
            CRP_CHECK(a, np <= (uint32_t)kListCap && nm <= (uint32_t)kListCap, 14);
#endif
            __syncthreads();                                   // the list buffer is free: the next tile's lists may land
            if (tid == 0 && lists_owed) {
                stage_lists(s ^ 1);
                lists_owed = false;
            }
            emit_strand_reg<kScore, false>(a, s_tab, rec, ep, np, base_p, td.t_start, td.L, slot_p);
            emit_strand_reg<kScore, true>(a, s_tab, rec, em, nm, base_m, td.t_start, td.L, slot_m);
        } else {
            // dense tile: compact it here, in windows of kListCap ranks (the front end of k_scan_score)
            const unsigned long long off = ring.pref[s][warp] - tile_pref;
            const uint32_t wordA = 64 * warp + lane;
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
            const uint32_t epA = op + (xA & 0xFFFFu), emA = om + (xA >> 16), epB = op + (xB & 0xFFFFu), emB = om + (xB >> 16);
            for (uint32_t lo = 0; lo < np || lo < nm; lo += kListCap) {
                const uint32_t cp = np > lo ? min(np - lo, (uint32_t)kListCap) : 0u;
                const uint32_t cm = nm > lo ? min(nm - lo, (uint32_t)kListCap) : 0u;
                if (lo) __syncthreads();
                list_hits_window(list_p, h.pA, epA, 32u * wordA + kWinBiasPlus, lo);
                list_hits_window(list_p, h.pB, epB, 32u * (wordA + 32) + kWinBiasPlus, lo);
                list_hits_window(list_m, h.mA, emA, 32u * wordA + kWinBiasMinus, lo);
                list_hits_window(list_m, h.mB, emB, 32u * (wordA + 32) + kWinBiasMinus, lo);
                __syncthreads();
                emit_strand<kScore, false>(a, s_tab, rec, list_p, cp, base_p + lo, td.t_start, td.L, tid);
                emit_strand<kScore, true>(a, s_tab, rec, list_m, cm, base_m + lo, td.t_start, td.L, tid);
            }
            // (the next tile's lists are requested at the top of the next iteration, behind the barrier below)
        }
        __syncthreads();                                       // slot s is free again
    }
    dbg_stamp(5);
    if (a.world > 1 && cta == 0) xchg_wait(a);
}
