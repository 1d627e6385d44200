// k_scan_sp: the single-pass scan + score kernel (sm_100a).
//
// One persistent CTA of 1024 threads per SM, split into 4 independent TEAMS of 256 threads
// (8 warps).  The Rule-Set-1 lane tables are staged ONCE per SM and shared by the teams;
// everything else (ring of staged tile records, hit lists, barriers) is per team, and teams
// only ever synchronise among themselves (named barriers, bar.sync id, 256).
//
// Tiles are handed out in genome order through a ticket counter.  Per tile, a team
//   1. waits for the bulk (TMA) copy of the self-contained tile record (issued one tile ahead),
//   2. tests the PAMs of both strands with bit operations on the planes (every thread owns
//      two 32-position words) and scans the per-word hit counts inside each warp,
//   3. publishes the tile's aggregate (hits per strand) in global memory, then warp 0 looks
//      back over the aggregates / inclusive prefixes of the preceding tiles (decoupled
//      look-back, 128 predecessors per round trip, all loads in flight at once) while the
//      other warps pop their hits into the team's compacted hit lists,
//   4. publishes the tile's inclusive prefix, and
//   5. one thread per candidate extracts the 30-base window from the staged words, evaluates
//      Rule Set 1 in the canonical OpenBLAS lane order and stores (pos, packed 30-mer, x)
//      coalesced at  exclusive prefix + rank  of the two ordered strand streams.
// A tile is therefore read from HBM exactly once and nothing is counted twice; there is no
// grid-wide barrier.  A team waits only for tiles with smaller tickets, which are held by
// teams that are already running, so the kernel needs no co-residency guarantee.
//
// Scan state (per genome, zeroed once when it is allocated, never re-zeroed):
//   agg[tile]   u64  epoch << 56 | plus hits << 28 | minus hits
//   incl[tile]  2 x u64  epoch << 56 | inclusive prefix (plus, minus); < 2^32 by construction
//   ctl[0]      ticket counter, ctl[1] finished-tile counter; the host passes their values
//               at launch (every launch adds a known amount), so they are never reset.
// The epoch (1..255, bumped per launch) makes words of earlier launches read as "not yet".
#pragma once
#include "scan.cuh"

static constexpr int kTeams = 4;
static constexpr int kSpThreads = kTeams * kThreads;       // 1024
static constexpr int kLookQ = 4;                            // look-back: 32 * kLookQ predecessors per round
static constexpr size_t kTeamStageBytes = ((size_t)kStages * kRecBytes + 127) / 128 * 128;
static constexpr size_t kTeamBytes = kTeamStageBytes + 2 * (size_t)kListCap * sizeof(uint16_t);

struct SpArgs {
    const uint4 *records;
    uint32_t n_tiles;
    int guide_len;
    uint32_t flags;
    uint32_t epoch;                  // 1..255
    uint32_t ticket_base, done_base; // values of ctl[0] / ctl[1] when the kernel starts
    const double *tables;
    uint64_t capacity;               // entries per strand stream
    uint32_t *pos_plus, *pos_minus;
    unsigned long long *packed_plus, *packed_minus;
    double *x_plus, *x_minus;
    unsigned long long *agg;         // [n_tiles]
    ulonglong2 *incl;                // [n_tiles]
    unsigned int *ctl;               // [2]
    unsigned long long *seg_counts;  // [2 * n_seg] out: plus[0..n_seg) then minus[0..n_seg)
    const uint32_t *seg_first_tile, *seg_tile_count;   // [n_seg], or NULL: one segment = all tiles
    uint32_t n_seg;
};

struct __align__(16) TeamCtl {
    unsigned long long full[kStages];   // mbarriers: bulk copy of the slot has landed
    unsigned long long base[2];         // exclusive prefix of the current tile: plus, minus
    uint32_t tile[kStages];             // staged tile, or kNoTile: no more tiles
    uint32_t wcnt[kWarps];              // hits of each warp chunk: plus | minus << 16
};

__device__ __forceinline__ void team_sync(int team) {
    asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "n"(kThreads) : "memory");
}
__device__ __forceinline__ void st_relaxed(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ ulonglong2 ld_relaxed2(const ulonglong2 *p) {
    ulonglong2 v;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}

static constexpr unsigned long long kValMask = (1ull << 56) - 1ull;

// Exclusive prefix (plus, minus) of `tile`: sum of the aggregates of the nearest preceding
// tiles down to the first one whose inclusive prefix is already published.  One warp.
__device__ __forceinline__ void look_back(const SpArgs &a, uint32_t tile, int lane, unsigned long long &base_p,
                                          unsigned long long &base_m) {
    const unsigned long long ep = (unsigned long long)a.epoch << 56;
    unsigned long long sp = 0, sm = 0;
    for (int64_t j0 = (int64_t)tile - 1;; j0 -= 32 * kLookQ) {
        unsigned long long ag[kLookQ];
        ulonglong2 in[kLookQ];
#pragma unroll
        for (int q = 0; q < kLookQ; ++q) {       // every load of the round is in flight before the first use
            const int64_t j = j0 - (32 * q + lane);
            if (j >= 0) {
                ag[q] = ld_relaxed(a.agg + j);
                in[q] = ld_relaxed2(a.incl + j);
            } else {                             // before the first tile: nothing, and known
                ag[q] = ep;
                in[q] = make_ulonglong2(ep, ep);
            }
        }
        bool found = false;                      // warp-uniform
        auto step = [&](unsigned long long &agq, ulonglong2 &inq, int q) {
            if (found) return;
            const int64_t j = j0 - (32 * q + lane);
            while (!__all_sync(0xFFFFFFFFu, (agq >> 56) == a.epoch)) {
                if ((agq >> 56) != a.epoch) {    // the tile has a ticket but has not counted yet
                    __nanosleep(64);
                    agq = ld_relaxed(a.agg + j);
                    inq = ld_relaxed2(a.incl + j);
                }
            }
            const bool has = (inq.x >> 56) == a.epoch && (inq.y >> 56) == a.epoch;
            const uint32_t bal = __ballot_sync(0xFFFFFFFFu, has);
            const int k = bal ? __ffs(bal) - 1 : 32;         // nearest tile with an inclusive prefix
            const uint32_t ap = (uint32_t)(agq >> 28) & 0x0FFFFFFFu, am = (uint32_t)agq & 0x0FFFFFFFu;
            sp += __reduce_add_sync(0xFFFFFFFFu, lane < k ? ap : 0u);
            sm += __reduce_add_sync(0xFFFFFFFFu, lane < k ? am : 0u);
            if (bal) {
                sp += (unsigned long long)__shfl_sync(0xFFFFFFFFu, (uint32_t)inq.x, k);
                sm += (unsigned long long)__shfl_sync(0xFFFFFFFFu, (uint32_t)inq.y, k);
                found = true;
            }
        };
        static_assert(kLookQ == 4, "unrolled by hand");
        step(ag[0], in[0], 0);
        step(ag[1], in[1], 1);
        step(ag[2], in[2], 2);
        step(ag[3], in[3], 3);
        if (found) {
            base_p = sp;
            base_m = sm;
            return;
        }
    }
}

// one thread per listed hit of one strand of a tile: window, score, coalesced stores
template <bool kScore, bool kMinus>
__device__ __forceinline__ void emit_strand_sp(const SpArgs &a, const double *__restrict__ tab,
                                               const uint4 *__restrict__ rec, const uint16_t *__restrict__ list,
                                               uint32_t count, uint64_t out0, uint32_t t_start, uint32_t L, uint32_t slot) {
    if (out0 >= a.capacity) return;
    if (count > a.capacity - out0) count = (uint32_t)(a.capacity - out0);
    uint32_t *const pos = (kMinus ? a.pos_minus : a.pos_plus) + out0;
    unsigned long long *const packed = (kMinus ? a.packed_minus : a.packed_plus) + out0;
    double *const xs = (kMinus ? a.x_minus : a.x_plus) + out0;
    for (uint32_t i = slot; i < count; i += kThreads) {
        const uint32_t pl = list[i], t = t_start + pl;
        __stcs(pos + i, t);
        if (kScore) {
            const Window w = extract_window<kMinus>(rec, pl + (kMinus ? kWinBiasMinus : kWinBiasPlus), t, L);
            const double x = rs1_canonical(tab, w.s0, w.s1, w.valid);
            __stcs(packed + i, w.packed);
            __stcs(xs + i, x);
        }
    }
}

template <bool kScore>
__global__ void __launch_bounds__(kSpThreads, 1)
k_scan_sp(const SpArgs a) {
    // dynamic shared memory: [lane tables][team 0: stages, hit lists][team 1] ...
    extern __shared__ __align__(128) unsigned char s_dyn[];
    constexpr size_t kTabBytes = kScore ? (kRs1TableBytes + 127) / 128 * 128 : 0;
    __shared__ TeamCtl s_ctl[kTeams];
    __shared__ __align__(8) unsigned long long s_tabbar;

    const int tid = threadIdx.x, team = tid / kThreads, ttid = tid % kThreads, lane = tid & 31, warp = ttid >> 5;
    const int l = a.guide_len;
    const double *const s_tab = reinterpret_cast<const double *>(s_dyn);
    unsigned char *const s_team = s_dyn + kTabBytes + (size_t)team * kTeamBytes;
    auto stage = [&](int s) { return reinterpret_cast<uint4 *>(s_team + (size_t)s * kRecBytes); };
    uint16_t *const list_p = reinterpret_cast<uint16_t *>(s_team + kTeamStageBytes), *const list_m = list_p + kListCap;
    TeamCtl &ctl = s_ctl[team];

    auto record = [&](uint32_t tile) { return a.records + (size_t)tile * kRecWords; };
    auto produce = [&](int s) {        // one thread of the team: next ticket, stage its tile into slot s
        const uint32_t q = atomicAdd(a.ctl, 1u) - a.ticket_base;
        const uint32_t t = q < a.n_tiles ? q : kNoTile;
        ctl.tile[s] = t;
        if (t != kNoTile) {
            mbar_expect(&ctl.full[s], kRecBytes);
            bulk_copy(stage(s), record(t), kRecBytes, &ctl.full[s]);
        } else {
            mbar_arrive(&ctl.full[s]);
        }
    };

    if (tid == 0) mbar_init(&s_tabbar, 1);
    if (ttid == 0)
        for (int s = 0; s < kStages; ++s) mbar_init(&ctl.full[s], 1);
    if (ttid == 0 || tid == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    if (kScore && tid == 0) {          // the lane tables arrive while the first tiles are tested
        mbar_expect(&s_tabbar, (uint32_t)kRs1TableBytes);
        bulk_copy(s_dyn, a.tables, (uint32_t)kRs1TableBytes, &s_tabbar);
    }
    if (ttid == 0) produce(0);

    const unsigned long long ep = (unsigned long long)a.epoch << 56;
    bool tables_ready = !kScore;
    for (uint32_t n = 0;; ++n) {
        const int s = n % kStages;
        // next tile of this team: its copy lands while this tile is emitted (slot s^1 was
        // released by the barrier that ended tile n - 1)
        if (ttid == 0) produce(s ^ 1);
        mbar_wait(&ctl.full[s], (n / kStages) & 1u);
        const uint32_t tile = ctl.tile[s];
        if (tile == kNoTile) break;
        const uint4 *rec = stage(s);
        const uint4 d = rec[0];
        const TileDesc td = {d.x, d.y, d.z, d.w};
        const uint32_t wordA = 64 * warp + lane;
        const Hits h = tile_hits(rec, td, l, wordA);
        // ---- warp scan of the per-word counts: the A words of the chunk precede its B words
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
        if (lane == 31) ctl.wcnt[warp] = totA + iB;
        team_sync(team);                                           // (1) the warp counts are there
        uint32_t off = 0, tot = 0;                                  // hits before my warp chunk / of the tile, packed
#pragma unroll
        for (int q = 0; q < kWarps; ++q) {
            const uint32_t c = ctl.wcnt[q];
            if (q < warp) off += c;
            tot += c;
        }
        const uint32_t np = tot & 0xFFFFu, nm = tot >> 16;
        if (warp == 0) {
            if (lane == 0 && !(a.flags & 0x10000000u)) st_relaxed(a.agg + tile, ep | ((unsigned long long)np << 28) | nm);
        }
        // rank (inside the tile, per strand) of the first hit of my words
        const uint32_t fA = off + iA - cA, fB = off + totA + iB - cB;
        const bool sparse = np <= (uint32_t)kListCap && nm <= (uint32_t)kListCap;
        if (sparse) {
            list_hits(list_p + (fA & 0xFFFFu), h.pA, 32u * wordA);
            list_hits(list_p + (fB & 0xFFFFu), h.pB, 32u * (wordA + 32));
            list_hits(list_m + (fA >> 16), h.mA, 32u * wordA);
            list_hits(list_m + (fB >> 16), h.mB, 32u * (wordA + 32));
        }
        if (warp == 0) {
            unsigned long long bp = 0, bm = 0;
            if (a.flags & 0x10000000u) {        // debug: prefixes left behind by the previous launch, no look-back
                if (tile) {
                    const ulonglong2 v = ld_relaxed2(a.incl + (tile - 1));
                    bp = v.x & kValMask;
                    bm = v.y & kValMask;
                }
            } else {
                look_back(a, tile, lane, bp, bm);
                if (lane == 0) {
                    ulonglong2 v = make_ulonglong2(ep | (bp + np), ep | (bm + nm));
                    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(a.incl + tile), "l"(v.x), "l"(v.y)
                                 : "memory");
                }
            }
            if (lane == 0) {
                ctl.base[0] = bp;
                ctl.base[1] = bm;
            }
        }
        if (kScore && !tables_ready) {
            mbar_wait(&s_tabbar, 0);
            tables_ready = true;
        }
        team_sync(team);                                           // (2) the hit lists and the tile's prefix are there
        const uint64_t base_p = ctl.base[0], base_m = ctl.base[1];
        if (a.flags & 0x20000000u) {
        } else if (sparse) {
            emit_strand_sp<kScore, false>(a, s_tab, rec, list_p, np, base_p, td.t_start, td.L, ttid);
            emit_strand_sp<kScore, true>(a, s_tab, rec, list_m, nm, base_m, td.t_start, td.L, ttid ^ (kThreads / 2));
        } else {                                                   // dense tile: windows of kListCap ranks
            for (uint32_t lo = 0; lo < np || lo < nm; lo += kListCap) {
                const uint32_t cp = np > lo ? min(np - lo, (uint32_t)kListCap) : 0u;
                const uint32_t cm = nm > lo ? min(nm - lo, (uint32_t)kListCap) : 0u;
                if (lo) team_sync(team);
                list_hits_window(list_p, h.pA, fA & 0xFFFFu, 32u * wordA, lo);
                list_hits_window(list_p, h.pB, fB & 0xFFFFu, 32u * (wordA + 32), lo);
                list_hits_window(list_m, h.mA, fA >> 16, 32u * wordA, lo);
                list_hits_window(list_m, h.mB, fB >> 16, 32u * (wordA + 32), lo);
                team_sync(team);
                emit_strand_sp<kScore, false>(a, s_tab, rec, list_p, cp, base_p + lo, td.t_start, td.L, ttid);
                emit_strand_sp<kScore, true>(a, s_tab, rec, list_m, cm, base_m + lo, td.t_start, td.L, ttid);
            }
        }
        // ---- the team that finishes the last tile derives the per-segment counts from the
        // inclusive prefixes (all published by then)
        if (warp == 0) {
            uint32_t last = 0;
            if (lane == 0) {
                __threadfence();
                last = atomicAdd(a.ctl + 1, 1u) - a.done_base == a.n_tiles - 1u;
            }
            last = __shfl_sync(0xFFFFFFFFu, last, 0);
            if (last) {
                __threadfence();
                for (uint32_t sg = lane; sg < a.n_seg; sg += 32) {
                    const uint32_t f = a.seg_first_tile ? a.seg_first_tile[sg] : 0u;
                    const uint32_t c = a.seg_tile_count ? a.seg_tile_count[sg] : a.n_tiles;
                    unsigned long long cp = 0, cm = 0;
                    if (c) {
                        const ulonglong2 hi = ld_relaxed2(a.incl + (f + c - 1));
                        cp = hi.x & kValMask;
                        cm = hi.y & kValMask;
                        if (f) {
                            const ulonglong2 lo = ld_relaxed2(a.incl + (f - 1));
                            cp -= lo.x & kValMask;
                            cm -= lo.y & kValMask;
                        }
                    }
                    a.seg_counts[sg] = cp;
                    a.seg_counts[a.n_seg + sg] = cm;
                }
            }
        }
        team_sync(team);                                           // (3) slot s and the lists are free again
    }
}
