// k_scan_pw: the scan + score kernel with the count phase folded into the emit phase
// ("progressive waves").
//
// k_scan_score (scan.cuh) counts ALL tiles, crosses a grid barrier, scans the range totals and only
// then emits its first candidate: 12-14 % of a launch at every genome size passes before the first
// store, bandwidth bound while the ALU pipes idle.  Here the tiles are cut into waves that double in
// size -- wave w holds 2^w tiles per CTA: tiles [G (2^w - 1), G (2^(w+1) - 1)) -- and the count of
// wave w + 1 is done WHILE wave w is emitted: every CTA, after the first tile it emits in wave w,
// counts its contiguous range of wave w + 1 straight from the PAM records in global memory (no
// staging: by then the CTAs of an SM are out of step, so the load latency of one hides behind the
// candidate bodies of the others), writes the prefix blocks of those tiles and publishes the range
// total.  The last CTA to publish scans the G totals of the wave into global memory and raises the
// wave's flag; nobody waits for it in practice, because a wave is counted a whole wave ahead.  Only
// wave 0 -- one tile per CTA -- is counted up front.  There is no grid barrier; the launch stays
// cooperative because CTAs wait for flags that other CTAs raise (co-residency), and every wait has a
// timeout that turns a lost flag into an error instead of a hung device.
//
// Everything that touches a candidate (hit tests, compaction, windows, Rule Set 1, stores) is the
// code of scan.cuh, unchanged; so are the output streams, the prefix blocks and the segment counts.
#pragma once
#include "scan.cuh"

struct PwCtl {                       // one per wave; zero before a launch, zeroed again by the last CTA of the launch
    unsigned int published;          // CTAs that have published their range total of this wave
    unsigned int ready;              // 1 once range_base[wave][*] is written
    unsigned int ticket;             // dispenser of the wave's dynamic tiles
    unsigned int pad;
};

struct PwArgs {
    ScanArgs s;                      // records, pam, streams, warp_pref, seg counts, exchange: as in k_scan_score
    PwCtl *ctl;                      // [n_waves + 1]; ctl[n_waves].published counts the CTAs that are done (self-cleaning)
    unsigned long long *wave_tot;    // [n_waves][G] range totals
    unsigned long long *range_base;  // [n_waves][G] global exclusive prefix of every range (counts are plus << 32 | minus)
    unsigned long long *grand;       // [n_waves] inclusive total through the wave
    uint32_t n_waves;
    unsigned long long timeout_ns;
    unsigned int *error;             // mapped host word: set if a wave's flag did not come up in time
};

// PAM hits of one tile counted by one warp straight from global memory: the count loop of
// warp_count_tile on ld.global.nc data (lane l owns the words 32 i + l; word + 1 is the neighbour's
// word of the same iteration, or lane 0's of the next one).
__device__ __forceinline__ void warp_count_tile_global(const unsigned char *__restrict__ rec, int l, int lane,
                                                       uint32_t *__restrict__ cnt) {
    const uint4 d = __ldg(reinterpret_cast<const uint4 *>(rec));
    const uint2 *w = reinterpret_cast<const uint2 *>(rec + 16);
    const int32_t t0 = (int32_t)d.x, L = (int32_t)d.y;
    const int32_t last_owned = t0 + (int32_t)d.z - 1;
    const int32_t hi_p = min(L - 3, last_owned);
    const int32_t hi_m = min(L - l + 7, hi_p);
    const bool edge = t0 < l + 5 || t0 + kTile - 1 > hi_m;         // warp-uniform
    uint2 v[17];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __ldg(w + 32 * i + lane);
    v[16] = __ldg(w + 512);                                        // right halo word (the same for every lane)
#pragma unroll
    for (int c = 0; c < kWarps; ++c) {
        uint32_t n = 0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int i = 2 * c + h;
            const uint2 a = v[i];
            // word + 1: lane + 1 of this iteration; for lane 31 lane 0 of the next iteration (v[16] after the last)
            uint2 an;
            an.x = __shfl_down_sync(0xFFFFFFFFu, a.x, 1);
            an.y = __shfl_down_sync(0xFFFFFFFFu, a.y, 1);
            const uint32_t nx = __shfl_sync(0xFFFFFFFFu, v[i + 1].x, 0), ny = __shfl_sync(0xFFFFFFFFu, v[i + 1].y, 0);
            if (lane == 31) {
                an.x = nx;
                an.y = ny;
            }
            uint32_t p = __funnelshift_r(a.x, an.x, 1) & __funnelshift_r(a.x, an.x, 2);
            uint32_t m = a.y & __funnelshift_r(a.y, an.y, 1);
            if (edge) {
                const int32_t tw = t0 + 32 * (32 * i + lane);
                p &= range_mask(tw, l + 5, hi_p);
                m &= range_mask(tw, 2, hi_m);
            }
            n += __popc(p) | (__popc(m) << 16);
        }
        n = __reduce_add_sync(0xFFFFFFFFu, n);
        if (lane == 0) cnt[c] = n;
    }
}

// first tile / tiles per CTA of wave w
__device__ __forceinline__ uint32_t pw_first(uint32_t G, uint32_t w) { return G * ((1u << w) - 1u); }
__device__ __forceinline__ uint32_t pw_wave_of(uint32_t G, uint32_t t) { return 31u - (uint32_t)__clz(t / G + 1u); }

// spin until *flag != 0 (one thread); false after the timeout
__device__ __forceinline__ bool pw_wait_flag(volatile unsigned int *flag, unsigned long long timeout_ns) {
    if (*flag) return true;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (!*flag) {
        __nanosleep(100);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > timeout_ns) return false;
    }
    return true;
}

// Count duty of one wave: this CTA's range of wave w -- tiles [first(w) + cta r, + r) with r = 2^w, clipped to
// the genome -- counted from the PAM records in global memory, prefix blocks written, range total published;
// the last CTA to publish scans the wave's totals into range_base and raises the wave's flag.  All threads.
__device__ __noinline__ void pw_count_duty(const PwArgs &pa, uint32_t w, uint32_t (*s_cnt)[kWarps], unsigned long long *s_scan,
                                           uint32_t *s_flag_p) {
    const ScanArgs &a = pa.s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int l = a.guide_len;
    const uint32_t G = gridDim.x, cta = blockIdx.x, nt = a.n_tiles;
    uint32_t &s_flag = *s_flag_p;
    const uint32_t r = 1u << w, w_lo = pw_first(G, w);
    const uint32_t r_lo = min(nt, w_lo + cta * r), n_mine = min(nt, r_lo + r) - r_lo;
    unsigned long long range_run = 0;
    for (uint32_t b_lo = 0;; b_lo += kMaxRange) {
        const uint32_t b_n = min(n_mine - min(n_mine, b_lo), (uint32_t)kMaxRange);
        for (uint32_t n = b_lo + warp; n < b_lo + b_n; n += kWarps)
            warp_count_tile_global(a.pam + (size_t)(r_lo + n) * kPamBytes, l, lane, s_cnt[n - b_lo]);
        __syncthreads();
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
            if (wq == kWarps - 1) pf[kWarps] = before + incl;
            asm volatile("fence.proxy.async.global;" ::: "memory");   // read back by bulk copies
        }
        range_run += total;
        __syncthreads();
        if (b_lo + kMaxRange >= n_mine) break;
    }
    // publish; the last publisher scans the wave's totals for everybody
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        pa.wave_tot[(size_t)w * G + cta] = range_run;
        __threadfence();
        s_flag = atomicAdd(&pa.ctl[w].published, 1u) == G - 1 ? 1u : 0u;
    }
    __syncthreads();
    if (s_flag) {
        __threadfence();
        unsigned long long v[4] = {0, 0, 0, 0}, mine = 0;     // thread owns ranges 4*tid .. 4*tid+3
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t i = 4u * tid + q;
            if (i < G) v[q] = *reinterpret_cast<volatile unsigned long long *>(pa.wave_tot + (size_t)w * G + i);
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
        unsigned long long before = w ? *reinterpret_cast<volatile unsigned long long *>(pa.grand + w - 1) : 0ull, total = 0;
#pragma unroll
        for (int q = 0; q < kWarps; ++q) {
            const unsigned long long x = s_scan[q];
            if (q < warp) before += x;
            total += x;
        }
        unsigned long long run = before + incl - mine;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t i = 4u * tid + q;
            if (i < G) pa.range_base[(size_t)w * G + i] = run;
            run += v[q];
        }
        if (tid == kThreads - 1) pa.grand[w] = run;            // inclusive total through this wave
        __threadfence();
        __syncthreads();
        if (tid == 0) *reinterpret_cast<volatile unsigned int *>(&pa.ctl[w].ready) = 1u;
    }
    __syncthreads();
}

template <bool kScore>
__global__ void __launch_bounds__(kThreads, CRP_CTAS_PER_SM)
k_scan_pw(const PwArgs pa) {
    const ScanArgs &a = pa.s;
    // dynamic shared memory: [stages][hit lists]
    extern __shared__ __align__(128) unsigned char s_dyn[];
    auto stage = [&](int s) { return reinterpret_cast<uint4 *>(s_dyn + (size_t)s * kRecBytes); };
    uint16_t *const s_list = reinterpret_cast<uint16_t *>(s_dyn + kStages * kRecBytes);
    __shared__ __align__(16) double s_tab[kScore ? RS1_TABLE_DOUBLES : 1];
    __shared__ Ring ring;
    __shared__ uint32_t s_cnt[kMaxRange][kWarps];
    __shared__ unsigned long long s_scan[kWarps];
    __shared__ uint32_t s_flag;                                    // CTA-wide verdicts of thread 0 (last publisher? flag seen?)
    __shared__ uint32_t s_pend;                                    // bit s: slot s is staged without its prefix block yet

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int l = a.guide_len;
    const uint32_t G = gridDim.x, cta = blockIdx.x;
    const uint32_t nt = a.n_tiles, n_waves = pa.n_waves;

    dbg_stamp(0);
    __shared__ __align__(8) unsigned long long s_tabbar;
    if (tid == 0) {
        mbar_init(&s_tabbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (kScore) {
            mbar_expect(&s_tabbar, (uint32_t)kRs1TableBytes);
            bulk_copy(s_tab, a.tables, (uint32_t)kRs1TableBytes, &s_tabbar);
        }
    }
    auto record = [&](uint32_t tile) { return a.records + (size_t)tile * kRecWords; };
    auto fail = [&](unsigned code) {
        if (pa.error) *reinterpret_cast<volatile unsigned int *>(pa.error) = code;
        if (a.fault) atomicCAS(a.fault, 0u, code | ((unsigned)__LINE__ << 8));
    };

    auto count_duty = [&](uint32_t w) { pw_count_duty(pa, w, s_cnt, s_scan, &s_flag); };

    // global exclusive prefix of the range that holds tile t (L2 loads: the arrays are written by other CTAs during this launch)
    auto tile_base = [&](uint32_t t) -> unsigned long long {
        const uint32_t w = pw_wave_of(G, t);
        return __ldcg(pa.range_base + (size_t)w * G + ((t - pw_first(G, w)) >> w));
    };
    unsigned int *const xchg_ctr = &pa.ctl[n_waves].ticket;        // CTAs that have sent their segment counts (self-cleaning)
    // per-segment candidate counts (and their exchange in a sharded scan): every CTA does its share once the
    // range bases of the last wave exist
    auto segment_counts = [&]() {
        const bool xchg = a.world > 1;
        for (uint32_t sg = cta * kThreads + tid; sg < a.seg_stride; sg += G * kThreads) {
            unsigned long long plus = 0, minus = 0;
            if (sg < a.n_seg) {
                const uint32_t f = a.seg_first_tile ? a.seg_first_tile[sg] : 0u;
                const uint32_t c = a.seg_tile_count ? a.seg_tile_count[sg] : a.n_tiles;
                if (c) {
                    const unsigned long long cnt = tile_base(f + c - 1) + __ldcg(a.warp_pref + (size_t)(f + c - 1) * kPrefWords + kWarps) -
                                                   (tile_base(f) + __ldcg(a.warp_pref + (size_t)f * kPrefWords));
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
        if (xchg && cta * kThreads < a.seg_stride) {               // this CTA wrote counts
            __threadfence_system();
            __syncthreads();
            if (tid == 0) {
                const uint32_t writers = min(G, (a.seg_stride + kThreads - 1) / kThreads);
                __threadfence_system();
                if (atomicAdd(xchg_ctr, 1u) == writers - 1) xchg_publish(a);
            }
        }
    };

    // ---------------------------------------------------------------- the producer (thread 0)
    // Walks this CTA's tile sequence: in every wave first its static tiles (first(w) + n G + cta), then
    // tickets of the wave's dispenser, then on to the next wave.  A tile of a wave whose range bases are
    // not known yet is staged "pending": the record is fetched (it depends on nothing), the prefix block
    // and the range base follow when the consumer has seen the wave's flag.
    uint32_t pw_w = 0, pw_n = 0;                                   // producer position (meaningful in thread 0)
    uint32_t ready_w = 0xFFFFFFFFu;                                // last wave whose flag this CTA has seen (uniform)
    auto static_share = [&](uint32_t w) -> uint32_t {
        const uint32_t ns = (uint32_t)(((unsigned long long)1u << w) * a.static_eighths / 8);
        return ns ? ns : 1u;
    };
    auto stage_rest = [&](int s) {                                  // prefix block + range base of a staged tile
        const uint32_t t = ring.tile[s], w = ring.done[s];          // ring.done[] holds the wave of the staged tile here
        asm volatile("fence.proxy.async.global;" ::: "memory");
        ring.rbase[s] = tile_base(t);
        bulk_copy(ring.pref[s], a.warp_pref + (size_t)t * kPrefWords, kPrefWords * 8, &ring.full[s]);
        (void)w;
    };
    auto produce = [&](int s) {
        uint32_t t = kNoTile;
        while (pw_w < n_waves) {
            const uint32_t w_lo = pw_first(G, pw_w), w_hi = min(nt, pw_first(G, pw_w + 1));
            const uint32_t ns = static_share(pw_w);
            const uint32_t dyn_lo = min(w_hi, w_lo + ns * G), n_dyn = w_hi - dyn_lo;
            if (pw_n < ns) {
                const uint32_t c = w_lo + pw_n * G + cta;
                if (c < dyn_lo) t = c;
            } else {
                const uint32_t q = atomicAdd(&pa.ctl[pw_w].ticket, 1u);
                if (q < n_dyn) t = dyn_lo + q;
            }
            if (t != kNoTile) {
                ++pw_n;
                break;
            }
            ++pw_w;
            pw_n = 0;
        }
        ring.tile[s] = t;
        if (t == kNoTile) {
            mbar_arrive(&ring.full[s]);
            return;
        }
        ring.done[s] = pw_w;
        mbar_expect(&ring.full[s], kRecBytes + kPrefWords * 8);
        bulk_copy(stage(s), record(t), kRecBytes, &ring.full[s]);
        const bool known = ready_w != 0xFFFFFFFFu && pw_w <= ready_w;
        if (known) {
            s_pend &= ~(1u << s);
            stage_rest(s);
        } else {
            s_pend |= 1u << s;
        }
    };

    ring_reset(ring, true, false);
    if (tid == 0) s_pend = 0u;
    uint32_t duty_w = 0;
    count_duty(duty_w++);
    dbg_stamp(2);
    if (tid == 0) produce(0);
    if (kScore) mbar_wait(&s_tabbar, 0);
    __syncthreads();

    uint16_t *const list_p = s_list, *const list_m = s_list + kListCap;
    auto see_wave = [&](uint32_t w) -> bool {                       // all threads: the range bases of wave w are in global memory
        if (tid == 0) s_flag = pw_wait_flag(&pa.ctl[w].ready, pa.timeout_ns) ? 1u : 0u;
        __syncthreads();
        const bool ok = s_flag != 0;
        __syncthreads();
        if (!ok) {
            if (tid == 0) fail(21);
            return false;
        }
        __threadfence();
        ready_w = w;
        if (w == 0) dbg_stamp(3);
        if (w + 1 == n_waves) segment_counts();
        return true;
    };
    for (uint32_t n_iter = 0;; ++n_iter) {
        const int s = n_iter % kStages;
        if (tid == 0) produce(s ^ 1);
        const uint32_t t_now = ring.tile[s];
        if (t_now == kNoTile) {
            mbar_wait(&ring.full[s], (n_iter / kStages) & 1u);
            // the marker just produced into the other slot is the end of the sequence as well: consume its phase
            mbar_wait(&ring.full[s ^ 1], ((n_iter + 1) / kStages) & 1u);
            break;
        }
        const uint32_t w = ring.done[s];
        while (duty_w <= w) count_duty(duty_w++);                   // nobody's flag may wait for a duty of mine that I postponed
        if (ready_w == 0xFFFFFFFFu || w > ready_w) {
            if (!see_wave(w)) return;
        }
        if (tid == 0 && (s_pend >> s & 1u)) {
            stage_rest(s);
            s_pend &= ~(1u << s);
        }
        mbar_wait(&ring.full[s], (n_iter / kStages) & 1u);
        const uint4 *rec = stage(s);
        const uint4 d = rec[0];
        const TileDesc td = {d.x, d.y, d.z, d.w};
        const unsigned long long tile_pref = ring.pref[s][0];
        const unsigned long long base = ring.rbase[s] + tile_pref;
        const unsigned long long off = ring.pref[s][warp] - tile_pref;
        const unsigned long long tot = ring.pref[s][kWarps] - tile_pref;
        const uint32_t np = (uint32_t)(tot >> 32), nm = (uint32_t)tot;
        const uint32_t base_p = (uint32_t)(base >> 32), base_m = (uint32_t)base;
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
#ifdef CRP_CHECKED
        {
            const uint32_t mine = totA + __shfl_sync(0xFFFFFFFFu, iB, 31);
            const unsigned long long nxt = ring.pref[s][warp + 1] - ring.pref[s][warp];
            CRP_CHECK(a, (uint32_t)(nxt >> 32) == (mine & 0xFFFFu) && (uint32_t)nxt == (mine >> 16), 7);
            CRP_CHECK(a, t_now < nt && pw_wave_of(G, t_now) == w, 8);
            CRP_CHECK(a, (unsigned long long)base_p + np <= 0xFFFFFFFFull && (unsigned long long)base_m + nm <= 0xFFFFFFFFull, 9);
        }
#endif
        if (np <= (uint32_t)kListCap && nm <= (uint32_t)kListCap) {
            list_hits(list_p + epA, h.pA, 32u * wordA + kWinBiasPlus);
            list_hits(list_p + epB, h.pB, 32u * (wordA + 32) + kWinBiasPlus);
            list_hits(list_m + emA, h.mA, 32u * wordA + kWinBiasMinus);
            list_hits(list_m + emB, h.mB, 32u * (wordA + 32) + kWinBiasMinus);
            __syncthreads();
            emit_strand<kScore, false>(a, s_tab, rec, list_p, np, base_p, td.t_start, td.L, tid);
            emit_strand<kScore, true>(a, s_tab, rec, list_m, nm, base_m, td.t_start, td.L, tid ^ (kThreads / 2));
        } else {                                               // dense tile: windows of kListCap ranks
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
        }
        __syncthreads();                                       // slot s and the lists are free again
        // the count duty of the wave after this tile's, a whole wave before anybody needs it
        while (duty_w < n_waves && duty_w <= w + 1) count_duty(duty_w++);
    }
    __syncthreads();
    // a CTA that ran out of tiles early still owes its duties, and its share of the segment counts
    while (duty_w < n_waves) count_duty(duty_w++);
    if (ready_w == 0xFFFFFFFFu || ready_w + 1 < n_waves) {
        if (!see_wave(n_waves - 1)) return;
    }
    dbg_stamp(5);
    // ---- self-cleaning control words: the last CTA through zeroes them for the next launch
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(&pa.ctl[n_waves].published, 1u) == G - 1) {
            for (uint32_t w = 0; w <= n_waves; ++w) {
                pa.ctl[w].published = 0;
                pa.ctl[w].ready = 0;
                pa.ctl[w].ticket = 0;
            }
            __threadfence();
        }
    }
    if (a.world > 1 && cta == 0) xchg_wait(a);
}
