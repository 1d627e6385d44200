// Device side of libcropsr_b200: pack, scan + score, segment counts, rescore (sm_100a).
//
// HBM layout of a packed genome shard: an array of TILE RECORDS, one per 16384 token
// positions of a segment, each record self-contained so that ONE bulk (TMA) copy stages
// everything a CTA needs for the tile:
//     word 0            descriptor {t_start, token length L, owned positions n, segment}
//     word 1            halo: the 32 positions before the tile
//     words 2 .. 513    the tile: 512 words of 32 positions
//     word 514          halo: the 32 positions after the tile
// A word is a uint4 of four 32-position bit planes {code low bit, code high bit (A0 T1 C2
// G3, CROPSR.py:300-302), lower-case, other byte}: 0.5 byte per base.
//
// k_scan_score is one cooperative, persistent launch.  Per wave of tiles:
//   count phase   every CTA counts the PAM hits of a contiguous range of tiles (bandwidth
//                 bound: bulk copies + a few bit ops per word) and publishes the range total
//                 and the range-local exclusive prefix of every tile;
//   grid barrier, every CTA scans the range totals into shared memory;
//   emit phase    tiles are handed out dynamically; the global output offset of a tile is
//                 range prefix + tile prefix (one load), so there is no ordering between
//                 CTAs and no spinning.  Hits are compacted through shared memory and one
//                 thread per candidate extracts the 30-base window from the staged words,
//                 scores it (fp64, canonical OpenBLAS lane order) and stores
//                 (pos, packed 30-mer, x) coalesced into the two ordered strand streams.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "rs1.cuh"

namespace cg = cooperative_groups;

// ------------------------------------------------------------------ geometry
static constexpr int kThreads = 256;
static constexpr int kWarps = kThreads / 32;
static constexpr int kTileWords = 2 * kThreads;            // every thread owns word tid and word tid + 256
static constexpr int kTile = kTileWords * 32;              // 16384 positions
static constexpr int kRecWords = kTileWords + 3;           // descriptor + halo + tile + halo
static constexpr uint32_t kRecBytes = kRecWords * 16;      // 8240, one bulk copy
static constexpr int kListCap = 1024;                      // hits per strand compacted per batch
static constexpr int kMaxRange = 32;                       // tiles per CTA per wave in the count phase
static constexpr uint32_t kAlign = 128;                    // positions; segment placement granularity

static_assert(kRecBytes % 16 == 0, "bulk copies move multiples of 16 bytes");

// record word 0
struct TileDesc {
    uint32_t t_start;   // token-relative position of the tile's first position
    uint32_t L;         // token length
    uint32_t n;         // positions of this tile owned by the segment (<= kTile)
    uint32_t segment;   // index of the segment in the genome
};

// what k_pack needs to build one record
struct PackDesc {
    uint64_t ascii_off;  // offset in the ASCII staging buffer of token position stage_begin
    uint32_t stage_begin, stage_end;   // token positions present in the staging buffer
    TileDesc td;
};

// ------------------------------------------------------------------ k_pack
// byte -> nibble: bit0 code low, bit1 code high (A0 T1 C2 G3), bit2 lower-case, bit3 other.
// 'U' and 'Z' are "other" bytes that still score (reference replace chains,
// CROPSR.py:120,128,458): they carry the code of T resp. G.
__host__ __device__ inline uint32_t classify(uint32_t c) {
    uint32_t up = c & 0xDFu;
    uint32_t r = 8u;
    if (up == 'A') r = 0u;
    else if (up == 'T') r = 1u;
    else if (up == 'C') r = 2u;
    else if (up == 'G') r = 3u;
    if (r < 8u) return r | ((c & 0x20u) >> 3);
    if (c == 'U') return 8u | 1u;
    if (c == 'Z') return 8u | 3u;
    return 8u;
}

// One thread per record word.  Positions outside the staged part of the token become
// "other" bytes (they never match a PAM and never score).
__global__ void __launch_bounds__(256)
k_pack(const uint8_t *__restrict__ ascii, const PackDesc *__restrict__ descs, uint64_t n_items,
       uint4 *__restrict__ records) {
    __shared__ uint8_t lut[256];
    lut[threadIdx.x] = (uint8_t)classify(threadIdx.x);
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t it = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; it < n_items; it += stride) {
        const uint64_t tile = it / kRecWords;
        const uint32_t k = (uint32_t)(it - tile * kRecWords);
        const PackDesc &pd = descs[tile];
        if (k == 0) {
            records[it] = make_uint4(pd.td.t_start, pd.td.L, pd.td.n, pd.td.segment);
            continue;
        }
        const int64_t p0 = (int64_t)pd.td.t_start + ((int64_t)k - 2) * 32;   // token position of bit 0
        const int64_t lo = pd.stage_begin, hi = pd.stage_end;
        uint32_t o0 = 0, o1 = 0, ol = 0, oo = 0;
        if (p0 >= lo && p0 + 32 <= hi) {
            const uint4 *src = reinterpret_cast<const uint4 *>(ascii + pd.ascii_off + (uint64_t)(p0 - lo));
            const uint4 a = __ldg(src), b = __ldg(src + 1);
            const uint32_t v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t nib = lut[(v[i] >> (8 * q)) & 0xFFu];
                    const int bit = 4 * i + q;
                    o0 |= (nib & 1u) << bit;
                    o1 |= ((nib >> 1) & 1u) << bit;
                    ol |= ((nib >> 2) & 1u) << bit;
                    oo |= ((nib >> 3) & 1u) << bit;
                }
            }
        } else {
            for (int bit = 0; bit < 32; ++bit) {
                const int64_t p = p0 + bit;
                uint32_t nib = 8u;
                if (p >= lo && p < hi) nib = lut[ascii[pd.ascii_off + (uint64_t)(p - lo)]];
                o0 |= (nib & 1u) << bit;
                o1 |= ((nib >> 1) & 1u) << bit;
                ol |= ((nib >> 2) & 1u) << bit;
                oo |= ((nib >> 3) & 1u) << bit;
            }
        }
        records[it] = make_uint4(o0, o1, ol, oo);
    }
}

// ------------------------------------------------------------------ small device helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// One bulk copy global -> shared, completion counted in bytes on the mbarrier (TMA).
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
    const uint32_t b = smem_u32(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(b)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    const uint32_t b = smem_u32(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(b),
        "r"(parity)
        : "memory");
}

// bits b of a 32-position word starting at token position t0 with lo <= t0+b <= hi
__device__ __forceinline__ uint32_t range_mask(int32_t t0, int32_t lo, int32_t hi) {
    int32_t a = lo - t0, b = hi - t0;
    if (a < 0) a = 0;
    if (b > 31) b = 31;
    if (a > b) return 0u;
    return (0xFFFFFFFFu >> (31 - b)) & (0xFFFFFFFFu << a);
}

struct Hits {
    uint32_t pA, mA, pB, mB;   // '+' / '-' hit masks of word tid (A) and word tid + 256 (B)
};

// PAM tests and bounds of one staged tile for this thread's two words.
//   '+': (?=.GG) at t  <=>  tok[t+1]==tok[t+2]=='G'            (CROPSR.py:415)
//   '-': (?=CC.) at t  <=>  tok[t]==tok[t+1]=='C' and t+2 < L  (CROPSR.py:426)
//   bounds (CROPSR.py:419 / :430): '+' t >= l+5;  '-' 2 <= t <= L-l+7
//   plus ownership: t inside the n positions of the tile that the segment owns.
// All positions fit int32: L < 2^31 - 2^15 (crp_genome_add_segment), 1 <= l <= 10^6.
__device__ __forceinline__ Hits tile_hits(const uint4 *__restrict__ rec, const TileDesc td, int l, int tid) {
    const uint4 a = rec[2 + tid], an = rec[3 + tid], b = rec[2 + kThreads + tid], bn = rec[3 + kThreads + tid];
    const uint32_t uA = ~(a.z | a.w), uAn = ~(an.z | an.w), uB = ~(b.z | b.w), uBn = ~(bn.z | bn.w);   // upper-case ACGT
    const uint32_t gA = a.x & a.y & uA, gAn = an.x & an.y & uAn, gB = b.x & b.y & uB, gBn = bn.x & bn.y & uBn;
    const uint32_t cA = ~a.x & a.y & uA, cAn = ~an.x & an.y & uAn, cB = ~b.x & b.y & uB, cBn = ~bn.x & bn.y & uBn;
    Hits h;
    h.pA = __funnelshift_r(gA, gAn, 1) & __funnelshift_r(gA, gAn, 2);
    h.pB = __funnelshift_r(gB, gBn, 1) & __funnelshift_r(gB, gBn, 2);
    h.mA = cA & __funnelshift_r(cA, cAn, 1);
    h.mB = cB & __funnelshift_r(cB, cBn, 1);
    const int32_t t0 = (int32_t)td.t_start, L = (int32_t)td.L;
    const int32_t last_owned = t0 + (int32_t)td.n - 1;
    const int32_t hi_p = min(L - 3, last_owned);
    const int32_t hi_m = min(L - l + 7, hi_p);
    if (t0 < l + 5 || t0 + kTile - 1 > hi_m) {   // edge tiles only
        const int32_t tA = t0 + 32 * tid, tB = tA + 32 * kThreads;
        h.pA &= range_mask(tA, l + 5, hi_p);
        h.pB &= range_mask(tB, l + 5, hi_p);
        h.mA &= range_mask(tA, 2, hi_m);
        h.mB &= range_mask(tB, 2, hi_m);
    }
    return h;
}

struct ScanArgs {
    const uint4 *records;            // n_tiles records of kRecWords words
    uint32_t n_tiles;
    uint32_t wave_tiles;             // tiles per wave (<= gridDim.x * kMaxRange)
    int guide_len;
    uint32_t flags;
    const double *tables;            // RS1 lane tables (RS1_TABLE_DOUBLES doubles)
    uint64_t capacity;               // entries per strand stream
    uint32_t *pos_plus, *pos_minus;
    unsigned long long *packed_plus, *packed_minus;
    double *x_plus, *x_minus;
    // scan state, zeroed before the launch.  Counts are (plus << 32) | minus.
    unsigned long long *tile_pref;   // [n_tiles] exclusive prefix of a tile inside its count range
    unsigned long long *tile_incl;   // [n_tiles] inclusive global prefix (written by the emit phase)
    unsigned long long *cta_tot;     // [2][gridDim.x] range totals, double-buffered by wave parity
    unsigned int *tickets;           // [n_waves] emit-phase tile dispensers
};

struct Window {
    uint32_t s0, s1, valid;              // planar codes / scoring mask of the 30-mer, output order
    unsigned long long packed;
};

// 30-base window of one hit out of the staged record.
// '+': tok[t-25, t+5) read backwards (output base q = tok[t+4-q]), upper-case bases complemented;
// '-': tok[t-2, t+28) read forwards.   pl = position inside the tile.
template <bool kMinus>
__device__ __forceinline__ Window extract_window(const uint4 *__restrict__ rec, uint32_t pl, uint32_t t, uint32_t L) {
    const uint32_t ws = pl + (kMinus ? 30u : 7u);           // relative to the first halo position
    const uint32_t wi = 1u + (ws >> 5), sh = ws & 31u;
    const uint4 lo = rec[wi], hi = rec[wi + 1];
    const uint32_t p0 = __funnelshift_r(lo.x, hi.x, sh), p1 = __funnelshift_r(lo.y, hi.y, sh);
    const uint32_t lw = __funnelshift_r(lo.z, hi.z, sh), ot = __funnelshift_r(lo.w, hi.w, sh);
    const uint32_t special = ot & p0;       // 'U' / 'Z': "other" bytes that still score
    const uint32_t valid = ~ot | special;
    Window w;
    if (kMinus) {
        w.s0 = (p0 ^ special) & 0x3FFFFFFFu;                 // U scores as A, Z as C
        w.s1 = p1 & 0x3FFFFFFFu;
        w.valid = valid & 0x3FFFFFFFu;
    } else {
        w.s0 = __brev(p0 ^ ~(lw | ot)) >> 2;                 // A<->T, C<->G of upper-case bases flips the low code bit
        w.s1 = __brev(p1) >> 2;
        w.valid = __brev(valid) >> 2;
    }
    uint32_t hi32 = w.s1;
    if (w.valid != 0x3FFFFFFFu) hi32 |= (uint32_t)(CRP_PACKED_UNSCORED >> 32);
    uint32_t lo32 = w.s0;
    if ((lw | ot) & 0x3FFFFFFFu) lo32 |= (uint32_t)CRP_PACKED_IRREGULAR;
    if (t + (kMinus ? 28u : 5u) > L) lo32 |= (uint32_t)CRP_PACKED_TRUNCATED;
    w.packed = ((unsigned long long)hi32 << 32) | lo32;
    return w;
}

__device__ __forceinline__ unsigned long long unpack_counts(uint32_t c) {   // (plus | minus << 16) -> plus << 32 | minus
    return ((unsigned long long)(c & 0xFFFFu) << 32) | (c >> 16);
}

// hits of one word into the compacted list, highest position first.  end = one past the
// slot of the word's last hit.
__device__ __forceinline__ void list_hits(uint16_t *__restrict__ end, uint32_t m, uint32_t pos0) {
    while (m) {
        const uint32_t b = 31u - (uint32_t)__clz(m);
        m ^= 1u << b;
        *--end = (uint16_t)(pos0 + b);
    }
}
// same, for a tile with more than kListCap hits on a strand: only ranks [lo, lo + kListCap)
__device__ __forceinline__ void list_hits_window(uint16_t *__restrict__ list, uint32_t m, uint32_t rank_end,
                                                 uint32_t pos0, uint32_t lo) {
    uint32_t r = rank_end;
    while (m) {
        const uint32_t b = 31u - (uint32_t)__clz(m);
        m ^= 1u << b;
        --r;
        if (r - lo < (uint32_t)kListCap) list[r - lo] = (uint16_t)(pos0 + b);
    }
}

// one thread per listed hit of one strand: window, score, coalesced stores
template <bool kScore, bool kMinus>
__device__ __forceinline__ void emit_strand(const ScanArgs &a, const double *__restrict__ tab,
                                            const uint4 *__restrict__ rec, const uint16_t *__restrict__ list,
                                            uint32_t count, uint64_t out0, uint32_t t_start, uint32_t L, uint32_t slot) {
    uint32_t *const pos = (kMinus ? a.pos_minus : a.pos_plus) + out0;
    unsigned long long *const packed = (kMinus ? a.packed_minus : a.packed_plus) + out0;
    double *const xs = (kMinus ? a.x_minus : a.x_plus) + out0;
    if (out0 >= a.capacity) return;
    if (count > a.capacity - out0) count = (uint32_t)(a.capacity - out0);
    for (uint32_t i = slot; i < count; i += kThreads) {
        const uint32_t pl = list[i], t = t_start + pl;
        __stcs(pos + i, t);
        if (kScore) {
            const Window w = extract_window<kMinus>(rec, pl, t, L);
            double x = rs1_canonical(tab, w.s0, w.s1, w.valid);
            if (a.flags & CRP_SCAN_LOGISTIC) x = 1.0 / (1.0 + exp(x));
            __stcs(packed + i, w.packed);
            __stcs(xs + i, x);
        }
    }
}

template <bool kScore>
__global__ void __launch_bounds__(kThreads, CRP_CTAS_PER_SM)
k_scan_score(const ScanArgs a) {
    // dynamic shared memory: [stage 0][stage 1][hit lists][range prefixes]
    extern __shared__ __align__(128) unsigned char s_dyn[];
    auto stage = [&](int b) { return reinterpret_cast<uint4 *>(s_dyn + (size_t)b * kRecBytes); };
    uint16_t *const s_list = reinterpret_cast<uint16_t *>(s_dyn + 2 * kRecBytes);
    unsigned long long *const s_rangepref =
        reinterpret_cast<unsigned long long *>(s_dyn + 2 * kRecBytes + 2 * kListCap * sizeof(uint16_t));
    __shared__ __align__(16) double s_tab[kScore ? RS1_TABLE_DOUBLES : 1];   // static: LDS takes the table offset as an immediate
    __shared__ __align__(8) unsigned long long s_bar[2];
    __shared__ uint32_t s_cnt[kMaxRange][kWarps];
    __shared__ uint2 s_wt[kWarps];
    __shared__ unsigned long long s_scan[kWarps];
    __shared__ unsigned long long s_base[2];     // output offsets (plus << 32 | minus) of the tile staged in buffer b
    __shared__ uint32_t s_next;

    cg::grid_group grid = cg::this_grid();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int l = a.guide_len;
    const uint32_t G = gridDim.x, cta = blockIdx.x;
    uint32_t phase = 0;     // bit b: parity of the next completion of stage b's mbarrier

    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (kScore)
        for (int i = tid; i < RS1_TABLE_DOUBLES; i += kThreads) s_tab[i] = a.tables[i];
    __syncthreads();

    auto record = [&](uint32_t tile) { return a.records + (size_t)tile * kRecWords; };
    auto wait_stage = [&](int b) {
        mbar_wait(&s_bar[b], (phase >> b) & 1u);
        phase ^= 1u << b;
    };

    unsigned long long wave_base = 0;
    uint32_t wave = 0;
    for (uint32_t w_lo = 0; w_lo < a.n_tiles; w_lo += a.wave_tiles, ++wave) {
        const uint32_t w_hi = min(a.n_tiles, w_lo + a.wave_tiles);
        const uint32_t k = (w_hi - w_lo + G - 1) / G;              // tiles per count range (<= kMaxRange)

        // ================================================= count phase: tiles [r_lo, r_hi)
        const uint32_t r_lo = min(w_hi, w_lo + cta * k), r_hi = min(w_hi, r_lo + k);
        const uint32_t n_mine = r_hi - r_lo;
        if (tid == 0) {
            if (n_mine > 0) bulk_load(stage(0), record(r_lo), kRecBytes, &s_bar[0]);
            if (n_mine > 1) bulk_load(stage(1), record(r_lo + 1), kRecBytes, &s_bar[1]);
        }
        for (uint32_t j = 0; j < n_mine; ++j) {
            const int b = j & 1;
            wait_stage(b);
            const uint4 d = stage(b)[0];
            const TileDesc td = {d.x, d.y, d.z, d.w};
            const Hits h = tile_hits(stage(b), td, l, tid);
            uint32_t c = (__popc(h.pA) + __popc(h.pB)) | ((__popc(h.mA) + __popc(h.mB)) << 16);
            c = __reduce_add_sync(0xFFFFFFFFu, c);
            if (lane == 0) s_cnt[j][warp] = c;
            __syncthreads();                                   // stage b is free again
            if (tid == 0 && j + 2 < n_mine) bulk_load(stage(b), record(r_lo + j + 2), kRecBytes, &s_bar[b]);
        }
        __syncthreads();
        if (warp == 0) {
            unsigned long long tot = 0;
            if ((uint32_t)lane < n_mine) {
#pragma unroll
                for (int q = 0; q < kWarps; ++q) tot += unpack_counts(s_cnt[lane][q]);
            }
            unsigned long long incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += v;
            }
            if ((uint32_t)lane < n_mine) a.tile_pref[r_lo + lane] = incl - tot;
            if (lane == 31) a.cta_tot[(wave & 1u) * G + cta] = incl;
        }
        grid.sync();

        // ================================================= exclusive scan of the range totals
        unsigned long long wave_total;
        {
            unsigned long long v[4] = {0, 0, 0, 0}, mine = 0;     // thread owns ranges 4*tid .. 4*tid+3
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t i = 4u * tid + q;
                if (i < G) v[q] = a.cta_tot[(wave & 1u) * G + i];
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
            unsigned long long before = 0, total = 0;
#pragma unroll
            for (int q = 0; q < kWarps; ++q) {
                const unsigned long long x = s_scan[q];
                if (q < warp) before += x;
                total += x;
            }
            unsigned long long run = wave_base + before + incl - mine;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t i = 4u * tid + q;
                if (i < G) s_rangepref[i] = run;
                run += v[q];
            }
            wave_total = total;
            __syncthreads();
        }

        // ================================================= emit phase: dynamic tiles of the wave
        unsigned int *const ticket = a.tickets + wave;
        if (tid == 0) {
            const uint32_t t0 = w_lo + atomicAdd(ticket, 1u);
            s_next = t0;
            if (t0 < w_hi) {
                bulk_load(stage(0), record(t0), kRecBytes, &s_bar[0]);
                s_base[0] = s_rangepref[(t0 - w_lo) / k] + a.tile_pref[t0];
            }
        }
        __syncthreads();
        uint32_t tile = s_next;
        for (uint32_t it = 0; tile < w_hi; ++it) {
            const int b = it & 1;
            // the next tile of this CTA: ticket now, its bulk copy lands while this tile is emitted
            uint32_t claimed = 0;
            unsigned long long claimed_pref = 0;
            if (tid == 0) claimed = w_lo + atomicAdd(ticket, 1u);
            wait_stage(b);
            const uint4 *rec = stage(b);
            const uint4 d = rec[0];
            const TileDesc td = {d.x, d.y, d.z, d.w};
            const Hits h = tile_hits(rec, td, l, tid);
            // ---- block scan of the per-word counts: all A words precede all B words
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
            if (lane == 31) s_wt[warp] = make_uint2(iA, iB);
            __syncthreads();
            if (tid == 0) {                                    // the ticket has arrived by now: start the prefetch
                s_next = claimed;
                if (claimed < w_hi) {
                    bulk_load(stage(b ^ 1), record(claimed), kRecBytes, &s_bar[b ^ 1]);
                    claimed_pref = __ldg(a.tile_pref + claimed);   // consumed after the list is built
                }
            }
            const uint2 wt = s_wt[lane & (kWarps - 1)];
            const uint32_t totA = __reduce_add_sync(0xFFFFFFFFu, lane < kWarps ? wt.x : 0u);
            const uint32_t totB = __reduce_add_sync(0xFFFFFFFFu, lane < kWarps ? wt.y : 0u);
            const uint32_t preA = __reduce_add_sync(0xFFFFFFFFu, lane < warp ? wt.x : 0u);
            const uint32_t preB = __reduce_add_sync(0xFFFFFFFFu, lane < warp ? wt.y : 0u);
            const uint32_t np = (totA & 0xFFFFu) + (totB & 0xFFFFu), nm = (totA >> 16) + (totB >> 16);
            // rank (inside the tile, per strand) one past the last hit of my words
            const uint32_t endA = preA + iA, endB = totA + preB + iB;
            const uint32_t epA = endA & 0xFFFFu, emA = endA >> 16, epB = endB & 0xFFFFu, emB = endB >> 16;
            const unsigned long long base = s_base[b];
            const uint64_t base_p = base >> 32, base_m = base & 0xFFFFFFFFull;
            if (tid == 0) a.tile_incl[tile] = base + (((unsigned long long)np << 32) | nm);

            uint16_t *const list_p = s_list, *const list_m = s_list + kListCap;
            if (np <= (uint32_t)kListCap && nm <= (uint32_t)kListCap) {
                list_hits(list_p + epA, h.pA, 32u * tid);
                list_hits(list_p + epB, h.pB, 32u * (kThreads + tid));
                list_hits(list_m + emA, h.mA, 32u * tid);
                list_hits(list_m + emB, h.mB, 32u * (kThreads + tid));
                if (tid == 0 && claimed < w_hi) s_base[b ^ 1] = s_rangepref[(claimed - w_lo) / k] + claimed_pref;
                __syncthreads();
                emit_strand<kScore, false>(a, s_tab, rec, list_p, np, base_p, td.t_start, td.L, tid);
                emit_strand<kScore, true>(a, s_tab, rec, list_m, nm, base_m, td.t_start, td.L, tid ^ (kThreads / 2));
            } else {                                           // pathological density: batches of kListCap ranks
                if (tid == 0 && claimed < w_hi) s_base[b ^ 1] = s_rangepref[(claimed - w_lo) / k] + claimed_pref;
                for (uint32_t lo = 0; lo < np || lo < nm; lo += kListCap) {
                    const uint32_t cp = np > lo ? min(np - lo, (uint32_t)kListCap) : 0u;
                    const uint32_t cm = nm > lo ? min(nm - lo, (uint32_t)kListCap) : 0u;
                    if (lo) __syncthreads();                   // previous batch fully consumed
                    list_hits_window(list_p, h.pA, epA, 32u * tid, lo);
                    list_hits_window(list_p, h.pB, epB, 32u * (kThreads + tid), lo);
                    list_hits_window(list_m, h.mA, emA, 32u * tid, lo);
                    list_hits_window(list_m, h.mB, emB, 32u * (kThreads + tid), lo);
                    __syncthreads();
                    emit_strand<kScore, false>(a, s_tab, rec, list_p, cp, base_p + lo, td.t_start, td.L, tid);
                    emit_strand<kScore, true>(a, s_tab, rec, list_m, cm, base_m + lo, td.t_start, td.L, tid);
                }
            }
            __syncthreads();                                   // stage b, the lists, s_next and s_base are settled
            tile = s_next;
        }
        wave_base += wave_total;
    }
}

// per-segment counts from the inclusive tile prefixes
__global__ void k_segment_counts(const unsigned long long *__restrict__ tile_incl,
                                 const uint32_t *__restrict__ seg_first_tile,
                                 const uint32_t *__restrict__ seg_tile_count, uint32_t n_seg,
                                 unsigned long long *__restrict__ counts /* [2*n_seg] */) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seg) return;
    const uint32_t f = seg_first_tile[s], c = seg_tile_count[s];
    unsigned long long end = 0, begin = 0;
    if (f > 0) begin = tile_incl[f - 1];
    end = c > 0 ? tile_incl[f + c - 1] : begin;
    counts[s] = (end >> 32) - (begin >> 32);
    counts[n_seg + s] = (end & 0xFFFFFFFFull) - (begin & 0xFFFFFFFFull);
}

struct RescoreItem {
    uint32_t tile;     // record holding token position t
    uint32_t pl;       // position of t inside the tile
    uint32_t strand;   // '+' or '-'
    uint32_t cls;
};

__global__ void k_rescore(const uint4 *__restrict__ records, const RescoreItem *__restrict__ items, uint64_t n,
                          double *__restrict__ x_out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const RescoreItem it = items[i];
    const uint4 *rec = records + (size_t)it.tile * kRecWords;
    const Window w = it.strand == '-' ? extract_window<true>(rec, it.pl, 0u, 0xFFFFFFFFu)
                                      : extract_window<false>(rec, it.pl, 0u, 0xFFFFFFFFu);
    x_out[i] = rs1_dense(w.s0, w.s1, w.valid, (int)(it.cls & 15u), (int)(it.cls >> 4));
}
