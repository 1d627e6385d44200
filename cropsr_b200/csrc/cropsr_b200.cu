// libcropsr_b200: CROPSR Cas9 gRNA candidate scan + Rule-Set-1 score on B200 (sm_100a).
//
// Kernels (all HBM-bound integer / fp64 work -- no tensor cores on this path):
//   k_pack        ASCII token bytes -> 4 bit-planes (code low bit, code high bit,
//                 lower-case, other-byte), 0.5 byte per base resident in HBM.
//   k_scan_score  one pass over the planes: PAM tests (+: .GG, -: CC.) as 32-wide
//                 bit ops, block scan, decoupled look-back across tiles for the
//                 ordered global offsets, then one thread per candidate extracts
//                 the 30-base window from shared memory, scores it (fp64, canonical
//                 OpenBLAS lane order) and stores (pos, packed 30-mer, x) coalesced.
//   k_rescore     dense re-evaluation of selected candidates in any BLAS lane class.
//   k_segment_counts  per-segment candidate counts from the tile prefix array.
//
// Reference semantics implemented here: /root/reference/CROPSR.py:413-434 (scan,
// bounds, windows, transforms) and :285-313 (rs1_score); see DESIGN.md.

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>
#include <new>
#include <vector>

#include "../../include/cropsr_b200.h"
#include "rs1_weights.inc"

#define CRP_ABI_VERSION 1

// ------------------------------------------------------------------ geometry
static constexpr int kWarps = 8;                            // worker warps per CTA
static constexpr int kCtaThreads = (kWarps + 1) * 32;       // + the service (look-back) warp
static constexpr int kWarpWords = 64;                       // plane words per warp-tile (2 per lane)
static constexpr int kWarpPos = kWarpWords * 32;            // 2048 positions
static constexpr int kTile = kWarps * kWarpPos;             // positions per CTA tile
static constexpr int kListCap = 128;                        // hits per strand compacted per round
static constexpr size_t kSmemTableBytes = (size_t)RS1_TABLE_DOUBLES * sizeof(double);
static constexpr size_t kSmemBytes = kSmemTableBytes + 2 * (size_t)kWarps * (kWarpWords + 2) * sizeof(uint4) +
                                     (size_t)kWarps * 2 * kListCap * sizeof(uint16_t);
static constexpr uint32_t kAlign = 128;            // positions; segment placement granularity

struct TileDesc {
    uint32_t gword;     // plane word index of the tile's first position
    uint32_t t_start;   // token-relative position of the tile's first position
    uint32_t L;         // token length
    uint32_t n;         // positions of this tile owned by the segment (<= kTile)
};

// status word of the decoupled look-back: [63:62] flag, [61:31] plus count, [30:0] minus count
static constexpr unsigned long long kFlagAgg = 1ull << 62;
static constexpr unsigned long long kFlagIncl = 2ull << 62;
static constexpr unsigned long long kValMask = (1ull << 62) - 1;
static constexpr unsigned long long kMinusMask = (1ull << 31) - 1;

// ------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
static int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
#define CUDA_TRY(expr)                                                                      \
    do {                                                                                    \
        cudaError_t e_ = (expr);                                                            \
        if (e_ != cudaSuccess)                                                              \
            return fail(CRP_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                \
    } while (0)

// ------------------------------------------------------------------ context
struct Context {
    bool ready = false;
    int device = -1;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    uint64_t launches = 0;
    double *d_tables = nullptr;        // RS1 lane tables in device memory
};
static Context g_ctx;

// ------------------------------------------------------------------ k_pack
// byte -> nibble: bit0 code low, bit1 code high (A0 T1 C2 G3), bit2 lower-case, bit3 other.
// 'U' and 'Z' are "other" bytes that still score (reference replace chains,
// CROPSR.py:120,128,458): they carry the code of T resp. G.
__device__ __forceinline__ uint32_t classify(uint32_t c) {
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

__global__ void __launch_bounds__(256)
k_pack(const uint4 *__restrict__ ascii, uint64_t n_words, uint32_t *__restrict__ p0,
       uint32_t *__restrict__ p1, uint32_t *__restrict__ lower, uint32_t *__restrict__ other) {
    __shared__ uint8_t lut[256];
    lut[threadIdx.x] = (uint8_t)classify(threadIdx.x);
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += stride) {
        uint4 a = __ldg(ascii + 2 * w);
        uint4 b = __ldg(ascii + 2 * w + 1);
        uint32_t v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t o0 = 0, o1 = 0, ol = 0, oo = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t nib = lut[(v[i] >> (8 * k)) & 0xFFu];
                int bit = 4 * i + k;
                o0 |= (nib & 1u) << bit;
                o1 |= ((nib >> 1) & 1u) << bit;
                ol |= ((nib >> 2) & 1u) << bit;
                oo |= ((nib >> 3) & 1u) << bit;
            }
        }
        p0[w] = o0;
        p1[w] = o1;
        lower[w] = ol;
        other[w] = oo;
    }
}

// ------------------------------------------------------------------ RS1 scoring
// Canonical lane order (rows handled by OpenBLAS' 4-row dgemv_t kernel): one sequential
// accumulator per column-mod-4 lane, i.e. per base class for the first-order term and per
// SECOND base for the dinucleotide term, columns in ascending order; lanes combined
// (p0+p2)+(p1+p3) = (A+C)+(T+G).  A lane's value is a function of which of its entries
// match, so the leading entries of every lane come from a table of exact sequential fp64
// sums (built on the host at crp_init, staged in shared memory; rs1_weights.inc).
// s0/s1: planar code bits of the scored 30-mer (bit q = base q), valid: bases that score.
__device__ __forceinline__ double rs1_canonical(const double *__restrict__ T, uint32_t s0, uint32_t s1,
                                                uint32_t valid) {
    const uint32_t mA = ~s1 & ~s0 & valid, mT = ~s1 & s0 & valid, mC = s1 & ~s0 & valid, mG = s1 & s0 & valid;
    RS1_LANE_SUMS(T, mA, mT, mC, mG)
    const double first = __dadd_rn(__dadd_rn(fA, fC), __dadd_rn(fT, fG));
    const double second = __dadd_rn(__dadd_rn(dA, dC), __dadd_rn(dT, dG));
    // (score_first + score_second + intersect + low_gc) * -1, CROPSR.py:312
    return -__dadd_rn(__dadd_rn(__dadd_rn(first, second), RS1_INTERCEPT), RS1_LOW_GC);
}

// Host side: exact sequential sums of every valid subset of each lane's table entries.
static int build_rs1_tables(std::vector<double> &tab, char *err, size_t errlen) {
    tab.assign(RS1_TABLE_DOUBLES, 0.0);
    std::vector<char> used(RS1_TABLE_DOUBLES, 0);
    for (const Rs1Lane &ln : kRs1Lanes) {
        for (uint32_t sub = 0; sub < (1u << ln.n_table); ++sub) {
            bool ok = true;                       // two entries at one position are mutually exclusive
            for (int i = 0; i < ln.n_table && ok; ++i)
                for (int j = i + 1; j < ln.n_table; ++j)
                    if ((sub >> i & 1) && (sub >> j & 1) && ln.entries[i].pos == ln.entries[j].pos) ok = false;
            if (!ok) continue;
            uint32_t h = 0;
            volatile double sum = 0.0;            // one IEEE add per entry, in ascending column order
            for (int i = 0; i < ln.n_table; ++i) {
                if (!(sub >> i & 1)) continue;
                for (int g = 0; g < ln.n_groups; ++g)
                    if (ln.groups[g].first_base == ln.entries[i].first_base) h += (1u << ln.entries[i].pos) * ln.groups[g].magic;
                sum = sum + ln.entries[i].weight;
            }
            const uint32_t idx = ln.offset + (h >> (32 - ln.bits));
            if (idx >= RS1_TABLE_DOUBLES || used[idx]) {
                snprintf(err, errlen, "rs1 table hash of lane %s is not injective", ln.name);
                return -1;
            }
            used[idx] = 1;
            tab[idx] = sum;
        }
    }
    return 0;
}

__constant__ double c_w1[120] = RS1_DENSE_FIRST;
__constant__ double c_w2[464] = RS1_DENSE_SECOND;

// Dense emulation of one row of np.matmul(matrix, weights) for a given lane class.
// ind(j) is the 0/1 matrix entry of column j.
template <typename Ind>
__device__ double blas_row(const double *w, int d, int cls, Ind ind) {
    if (cls == CRP_CLASS_CANONICAL) {
        double p[4] = {0.0, 0.0, 0.0, 0.0};
        for (int j = 0; j < d; ++j)
            if (ind(j)) p[j & 3] = __dadd_rn(p[j & 3], w[j]);
        return __dadd_rn(__dadd_rn(p[0], p[2]), __dadd_rn(p[1], p[3]));
    }
    if (cls == CRP_CLASS_PAIR) {
        double q[2] = {0.0, 0.0};
        for (int j = 0; j < d; ++j)
            if (ind(j)) q[j & 1] = __dadd_rn(q[j & 1], w[j]);
        return __dadd_rn(q[0], q[1]);
    }
    // CRP_CLASS_SINGLE: OpenBLAS ddot (AVX-512): 4 accumulators x 8 lanes over
    // the 32-column blocks, folded to 4 lanes, one 16-column pass, lane-wise
    // ((a0+a1)+a2)+a3, (l0+l2)+(l1+l3), then a sequential tail.
    double acc[4][8];
    for (int a = 0; a < 4; ++a)
        for (int l = 0; l < 8; ++l) acc[a][l] = 0.0;
    const int n32 = d & ~31;
    for (int j = 0; j < n32; ++j)
        if (ind(j)) {
            int a = (j & 31) >> 3, l = j & 7;
            acc[a][l] = __dadd_rn(acc[a][l], w[j]);
        }
    double f[4][4];
    for (int a = 0; a < 4; ++a)
        for (int i = 0; i < 4; ++i) f[a][i] = __dadd_rn(acc[a][i], acc[a][i + 4]);
    int pos = n32;
    if (d & 16) {
        for (int a = 0; a < 4; ++a)
            for (int i = 0; i < 4; ++i) {
                int j = pos + 4 * a + i;
                if (ind(j)) f[a][i] = __dadd_rn(f[a][i], w[j]);
            }
        pos += 16;
    }
    double t[4];
    for (int i = 0; i < 4; ++i)
        t[i] = __dadd_rn(__dadd_rn(__dadd_rn(f[0][i], f[1][i]), f[2][i]), f[3][i]);
    double dot = __dadd_rn(__dadd_rn(t[0], t[2]), __dadd_rn(t[1], t[3]));
    for (int j = pos; j < d; ++j)
        if (ind(j)) dot = __dadd_rn(dot, w[j]);
    return dot;
}

__device__ double rs1_dense(uint32_t s0, uint32_t s1, uint32_t valid, int cls1, int cls2) {
    auto code = [&](int p) -> int { return (int)(((s1 >> p) & 1u) << 1 | ((s0 >> p) & 1u)); };
    auto ok = [&](int p) -> bool { return (valid >> p) & 1u; };
    double first = blas_row(c_w1, 120, cls1, [&](int j) { int p = j >> 2; return ok(p) && code(p) == (j & 3); });
    double second = blas_row(c_w2, 464, cls2, [&](int j) {
        int p = j >> 4;
        return ok(p) && ok(p + 1) && code(p) == ((j >> 2) & 3) && code(p + 1) == (j & 3);
    });
    return -__dadd_rn(__dadd_rn(__dadd_rn(first, second), RS1_INTERCEPT), RS1_LOW_GC);
}

// ------------------------------------------------------------------ plane algebra
struct Derived {
    uint32_t s0p;    // scored low code bit, '+' strand (complement of upper-case bases)
    uint32_t s0m;    // scored low code bit, '-' strand
    uint32_t s1;     // scored high code bit (both strands)
    uint32_t valid;  // base contributes to the score
    uint32_t irr;    // byte is not an upper-case ACGT
    uint32_t gup;    // upper-case G
    uint32_t cup;    // upper-case C
};

__device__ __forceinline__ Derived derive(uint32_t p0, uint32_t p1, uint32_t lo, uint32_t ot) {
    Derived d;
    const uint32_t upper = ~lo & ~ot;       // upper-case ACGT
    const uint32_t special = ot & p0;       // 'U' or 'Z'
    d.s0p = p0 ^ upper;                     // A<->T, C<->G flips the low bit
    d.s0m = p0 ^ special;                   // '-' strand: U scores as A, Z as C
    d.s1 = p1;
    d.valid = ~ot | special;
    d.irr = lo | ot;
    d.gup = p0 & p1 & upper;
    d.cup = ~p0 & p1 & upper;
    return d;
}

// bits b of a 32-position word starting at token position t0 with lo <= t0+b <= hi
__device__ __forceinline__ uint32_t range_mask(int64_t t0, int64_t lo, int64_t hi) {
    int64_t a = lo - t0, b = hi - t0;
    if (a < 0) a = 0;
    if (b > 31) b = 31;
    if (a > b) return 0u;
    return (0xFFFFFFFFu >> (31 - (int)b)) & (0xFFFFFFFFu << (int)a);
}

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// optional timeline instrumentation (tools/tile_timeline.py): 8 x u64 per tile, or NULL
__device__ unsigned long long *g_dbg_times = nullptr;
__device__ __forceinline__ void dbg_stamp(uint32_t tile, int slot) {
    unsigned long long *p = g_dbg_times;
    if (p) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p[8ull * tile + slot] = t;
    }
}

struct ScanArgs {
    const uint32_t *p0, *p1, *lower, *other;
    const TileDesc *tiles;
    uint32_t n_tiles;
    int guide_len;
    uint32_t flags;
    unsigned long long *status;      // [n_tiles], zeroed before launch
    unsigned int *ticket;            // tile dispenser, zeroed before launch
    const double *tables;            // RS1 lane tables (RS1_TABLE_DOUBLES doubles)
    uint64_t capacity;               // entries per strand stream
    uint32_t *pos_plus, *pos_minus;
    unsigned long long *packed_plus, *packed_minus;
    double *x_plus, *x_minus;
};

struct Hit {
    uint32_t s0, s1, valid;              // planar codes / scoring mask of the 30-mer, output order
    unsigned long long packed;
};

// 30-base window of one hit out of the warp's staged plane words.
// '+': tok[t-25, t+5) read backwards (output base q = tok[t+4-q]), upper-case bases complemented;
// '-': tok[t-2, t+28) read forwards.
template <bool kMinus>
__device__ __forceinline__ Hit extract_window(const uint4 *raw, uint32_t pl, uint32_t t, uint32_t L) {
    const uint32_t ws = pl + 32u - (kMinus ? 2u : 25u);
    const uint32_t wi = ws >> 5, sh = ws & 31u;
    const uint4 lo = raw[wi], hi = raw[wi + 1];
    const uint32_t p0 = __funnelshift_r(lo.x, hi.x, sh), p1 = __funnelshift_r(lo.y, hi.y, sh);
    const uint32_t lw = __funnelshift_r(lo.z, hi.z, sh), ot = __funnelshift_r(lo.w, hi.w, sh);
    const uint32_t upper = ~lw & ~ot;       // upper-case ACGT
    const uint32_t special = ot & p0;       // 'U' / 'Z': "other" bytes that still score
    const uint32_t valid = ~ot | special;
    Hit h;
    if (kMinus) {
        h.s0 = (p0 ^ special) & 0x3FFFFFFFu;        // U scores as A, Z as C
        h.s1 = p1 & 0x3FFFFFFFu;
        h.valid = valid & 0x3FFFFFFFu;
    } else {
        h.s0 = __brev(p0 ^ upper) >> 2;             // A<->T, C<->G flips the low code bit
        h.s1 = __brev(p1) >> 2;
        h.valid = __brev(valid) >> 2;
    }
    h.packed = (unsigned long long)h.s0 | ((unsigned long long)h.s1 << 32);
    if ((lw | ot) & 0x3FFFFFFFu) h.packed |= CRP_PACKED_IRREGULAR;
    if ((uint64_t)t + (kMinus ? 28u : 5u) > L) h.packed |= CRP_PACKED_TRUNCATED;
    if (h.valid != 0x3FFFFFFFu) h.packed |= CRP_PACKED_UNSCORED;
    return h;
}

// Per-warp state of a tile whose hits are known but not yet scored.
struct Pending {
    TileDesc td;
    uint32_t tile;
    uint32_t hit[2][2];     // [strand][word] hit masks of this lane's two words
    uint32_t excl;          // packed (plus | minus << 16) rank of this lane's first hit inside the warp-tile
    uint32_t wtot;          // packed hit totals of the warp-tile
    uint32_t cta_excl;      // packed hits of the CTA tile that precede this warp-tile
};

// Every spin in the kernel goes through here: back off, and trap instead of hanging the
// GPU if a wait ever exceeds ~1 s (a logic error or a grid that is not co-resident).
__device__ __forceinline__ void spin_pause(uint32_t &spins) {
    __nanosleep(32);
    if (++spins > (1u << 23)) __trap();
}

// named barrier 3 = the worker warps among themselves (the service warp never joins, so a
// look-back in flight never blocks the workers); hand-offs to and from the service warp go
// through sequence-numbered shared-memory slots.
template <int kId>
__device__ __forceinline__ void bar_sync(int n) {
    asm volatile("barrier.sync.aligned %0, %1;" ::"n"(kId), "r"(n) : "memory");
}
template <int kId>
__device__ __forceinline__ void bar_arrive(int n) {
    asm volatile("barrier.arrive.aligned %0, %1;" ::"n"(kId), "r"(n) : "memory");
}

// Persistent, software-pipelined kernel.  CTA tile = kWarps warp-tiles of kWarpPos
// positions; worker warp w owns warp-tile w; the extra (last) warp publishes the CTA's
// counts and runs the decoupled look-back (128-tile window).  Phase 1 (stage planes,
// PAM tests, counts) of tile i+1 runs BEFORE phase 2 (windows, scores, stores) of tile i,
// so the look-back of a tile has a whole tile period to complete.  Workers meet on one
// named barrier per tile and never wait for the service warp except for a prefix that is
// a tile period old; shared slots are double-buffered by tile parity.
template <bool kScore>
__global__ void __launch_bounds__(kCtaThreads, 3)
k_scan_score(const ScanArgs a) {
    // dynamic shared memory: [lane tables][staged plane words, 2 buffers][hit lists]
    extern __shared__ __align__(16) unsigned char s_dyn[];
    double *s_tab = reinterpret_cast<double *>(s_dyn);
    typedef uint4 RawBuf[kWarps][kWarpWords + 2];          // {p0, p1, lower, other} per word, 1 halo word each side
    RawBuf *s_raw = reinterpret_cast<RawBuf *>(s_dyn + kSmemTableBytes);
    typedef uint16_t ListBuf[2][kListCap];                 // warp-tile-local hit positions, '+' then '-'
    ListBuf *s_list = reinterpret_cast<ListBuf *>(s_dyn + kSmemTableBytes + 2 * sizeof(RawBuf));
    __shared__ uint32_t s_tot[2][kWarps];
    __shared__ unsigned long long s_prefix[2];
    __shared__ volatile uint32_t s_prefix_seq[2];
    __shared__ volatile uint32_t s_tot_seq[2];
    __shared__ volatile uint32_t s_tile[2];          // tile id of iteration it (slot it & 1)
    __shared__ uint32_t s_tile_seq[2];               // it + 1 once s_tile holds iteration it's tile
    __shared__ uint32_t s_arrive[2];                 // worker warps that reached iteration it

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int l = a.guide_len;
    const bool service = warp == kWarps;
    uint32_t spins = 0;
    if (threadIdx.x < 2) s_prefix_seq[threadIdx.x] = s_tot_seq[threadIdx.x] = s_tile_seq[threadIdx.x] = s_arrive[threadIdx.x] = 0;
    if (kScore)
        for (int i = threadIdx.x; i < RS1_TABLE_DOUBLES; i += kCtaThreads) s_tab[i] = a.tables[i];
    __syncthreads();

    // ---------------- tile id of this CTA's iteration `it`.  Tiles are handed out in the order
    // CTAs become READY for them: the global ticket is taken by the last worker warp to
    // finish its previous scoring phase, so a tile's counts are published a fixed ~2 us after
    // its ticket and the look-back of a later tile never waits for a CTA that is busy scoring.
    auto get_tile = [&](uint32_t it) -> uint32_t {
        const int par = it & 1;
        uint32_t tile = 0;
        if (lane == 0) {
            volatile uint32_t *seq = (volatile uint32_t *)&s_tile_seq[par];
            if (!service && atomicAdd(&s_arrive[par], 1u) == (uint32_t)kWarps - 1u) {
                s_arrive[par] = 0;                       // next use: iteration it + 2
                s_tile[par] = atomicAdd(a.ticket, 1u);
                __threadfence_block();
                *seq = it + 1;
            } else {
                while (*seq != it + 1) spin_pause(spins);
            }
            __threadfence_block();
            tile = s_tile[par];
        }
        return __shfl_sync(0xFFFFFFFFu, tile, 0);
    };

    // ---------------- phase 1 of tile number `it` of this CTA
    auto phase1 = [&](uint32_t it, uint32_t tile, Pending &pd) {
        const int par = it & 1;
        pd.tile = tile;
        pd.td = a.tiles[tile];
        const TileDesc &td = pd.td;
        if (service) {
            if (lane == 0) dbg_stamp(tile, 1);
            while (s_tot_seq[par] != it + 1) spin_pause(spins);      // worker totals of this tile are ready
            __threadfence_block();
            if (lane == 0) dbg_stamp(tile, 2);
            const uint32_t v = lane < kWarps ? s_tot[par][lane] : 0u;
            const uint32_t tot = __reduce_add_sync(0xFFFFFFFFu, v);
            const unsigned long long mine = ((unsigned long long)(tot & 0xFFFFu) << 31) | (tot >> 16);
            unsigned long long prefix = 0;
            uint32_t dbg_windows = 0, dbg_polls = 0;
            if (tile > 0) {
                int64_t j = (int64_t)tile - 1;
                for (;;) {
                    ++dbg_windows;
                    // lane reads 4 consecutive predecessors, nearest first (4 loads in flight)
                    unsigned long long sv[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int64_t idx = j - 4 * lane - q;
                        sv[q] = idx >= 0 ? ld_status(a.status + idx) : kFlagIncl;
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int64_t idx = j - 4 * lane - q;
                        while ((sv[q] >> 62) == 0) {
                            spin_pause(spins);
                            ++dbg_polls;
                            sv[q] = ld_status(a.status + idx);
                        }
                    }
                    unsigned long long acc = 0;
                    bool found = false;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (!found) acc += sv[q] & kValMask;
                        found = found || (sv[q] >> 62) == 2;
                    }
                    const uint32_t fm = __ballot_sync(0xFFFFFFFFu, found);
                    const int first = fm ? __ffs(fm) - 1 : 31;
                    unsigned long long c = lane <= first ? acc : 0ull;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
                    prefix += c;
                    if (fm) break;
                    j -= 128;
                }
            }
            // atomicMax: the inclusive word (flag 2) always wins over the counts word (flag 1)
            if (lane == 0) atomicMax(a.status + tile, kFlagIncl | (prefix + mine));
            dbg_polls = __reduce_max_sync(0xFFFFFFFFu, dbg_polls);
            if (lane == 0) {
                dbg_stamp(tile, 3);
                if (g_dbg_times) g_dbg_times[8ull * tile + 7] = ((unsigned long long)dbg_windows << 32) | dbg_polls;
                s_prefix[par] = prefix;
                __threadfence_block();
                s_prefix_seq[par] = it + 1;
            }
            return;
        }
        // ---- worker: stage this warp-tile, find its hits
        uint4 *raw = s_raw[par][warp];
        const int64_t n_w = (int64_t)td.n - (int64_t)warp * kWarpPos;      // owned positions of this warp-tile
        pd.hit[0][0] = pd.hit[0][1] = pd.hit[1][0] = pd.hit[1][1] = 0u;
        if (n_w > 0) {
            const uint64_t w0 = (uint64_t)td.gword + (uint64_t)warp * kWarpWords + 2 * lane;
            const uint2 q0 = __ldg((const uint2 *)(a.p0 + w0)), q1 = __ldg((const uint2 *)(a.p1 + w0));
            const uint2 ql = __ldg((const uint2 *)(a.lower + w0)), qo = __ldg((const uint2 *)(a.other + w0));
            uint4 edge = make_uint4(0u, 0u, 0u, 0u);
            if (lane == 0 || lane == 31) {
                const uint64_t we = lane == 0 ? w0 - 1 : w0 + 2;
                edge = make_uint4(__ldg(a.p0 + we), __ldg(a.p1 + we), __ldg(a.lower + we), __ldg(a.other + we));
                raw[lane == 0 ? 0 : kWarpWords + 1] = edge;
            }
            raw[1 + 2 * lane] = make_uint4(q0.x, q1.x, ql.x, qo.x);
            raw[2 + 2 * lane] = make_uint4(q0.y, q1.y, ql.y, qo.y);
            // upper-case G / C masks of my two words and of the word after them
            const uint32_t uA = ~ql.x & ~qo.x, uB = ~ql.y & ~qo.y;
            const uint32_t gA = q0.x & q1.x & uA, gB = q0.y & q1.y & uB;
            const uint32_t cA = ~q0.x & q1.x & uA, cB = ~q0.y & q1.y & uB;
            uint32_t gN = __shfl_down_sync(0xFFFFFFFFu, gA, 1), cN = __shfl_down_sync(0xFFFFFFFFu, cA, 1);
            if (lane == 31) {
                const uint32_t uN = ~edge.z & ~edge.w;
                gN = edge.x & edge.y & uN;
                cN = ~edge.x & edge.y & uN;
            }
            // '+': (?=.GG) at t <=> tok[t+1]==tok[t+2]=='G'   (CROPSR.py:415)
            pd.hit[0][0] = __funnelshift_r(gA, gB, 1) & __funnelshift_r(gA, gB, 2);
            pd.hit[0][1] = __funnelshift_r(gB, gN, 1) & __funnelshift_r(gB, gN, 2);
            // '-': (?=CC.) at t <=> tok[t]==tok[t+1]=='C' and t+2 < L   (CROPSR.py:426)
            pd.hit[1][0] = cA & __funnelshift_r(cA, cB, 1);
            pd.hit[1][1] = cB & __funnelshift_r(cB, cN, 1);
            // bounds tests of CROPSR.py:419 / :430:  '+' t >= l+5,  '-' 2 <= t <= L-l+7;
            // plus ownership (t inside this segment's tile) and t+2 < L.  Interior warp-tiles skip this.
            const int64_t t_w = (int64_t)td.t_start + (int64_t)warp * kWarpPos;
            const int64_t L = td.L;
            const int64_t last_owned = (int64_t)td.t_start + td.n - 1;
            const int64_t hi_p = L - 3 < last_owned ? L - 3 : last_owned;
            const int64_t hi_m = L - l + 7 < hi_p ? L - l + 7 : hi_p;
            if (t_w < l + 5 || t_w + kWarpPos - 1 > hi_m) {
                const int64_t t0 = t_w + 64 * lane;
                pd.hit[0][0] &= range_mask(t0, l + 5, hi_p);
                pd.hit[0][1] &= range_mask(t0 + 32, l + 5, hi_p);
                pd.hit[1][0] &= range_mask(t0, 2, hi_m);
                pd.hit[1][1] &= range_mask(t0 + 32, 2, hi_m);
            }
        }
        const uint32_t cnt = (__popc(pd.hit[0][0]) + __popc(pd.hit[0][1])) |
                             ((__popc(pd.hit[1][0]) + __popc(pd.hit[1][1])) << 16);
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += v;
        }
        pd.wtot = __shfl_sync(0xFFFFFFFFu, incl, 31);
        pd.excl = incl - cnt;
        if (lane == 31) s_tot[par][warp] = incl;
        bar_sync<3>(kWarps * 32);                 // workers only: never blocked by a look-back in flight
        const uint32_t tv = lane < kWarps ? s_tot[par][lane] : 0u;
        pd.cta_excl = __reduce_add_sync(0xFFFFFFFFu, lane < warp ? tv : 0u);
        if (warp == 0) {
            // publish this tile's counts at once (the service warp upgrades them to an
            // inclusive prefix later) and hand the totals to the service warp
            const uint32_t tot = __reduce_add_sync(0xFFFFFFFFu, tv);
            if (lane == 0) {
                atomicMax(a.status + tile, kFlagAgg | ((unsigned long long)(tot & 0xFFFFu) << 31) | (tot >> 16));
                dbg_stamp(tile, 0);
                __threadfence_block();
                s_tot_seq[par] = it + 1;
            }
        }
    };

    // ---------------- phase 2 (workers): compact hits, one lane per hit: window, score, store
    auto phase2 = [&](uint32_t it, const Pending &pd) {
        const int par = it & 1;
        const TileDesc &td = pd.td;
        const uint32_t np = pd.wtot & 0xFFFFu, nm = pd.wtot >> 16;
        const uint4 *raw = s_raw[par][warp];
        uint16_t(*list)[kListCap] = s_list[warp];
        const uint32_t ex_p = pd.excl & 0xFFFFu, ex_m = pd.excl >> 16;
        const uint32_t t_w = td.t_start + (uint32_t)warp * kWarpPos;
        // Global rows of this CTA tile (the look-back finished long ago in steady state).
        // Every worker warp waits here, hits or not: it is also the flow control that keeps
        // the workers from reusing this parity's shared slots before the service warp has
        // consumed them.
        if (warp == 0 && lane == 0) dbg_stamp(pd.tile, 4);
        while (s_prefix_seq[par] != it + 1) spin_pause(spins);
        if (warp == 0 && lane == 0) dbg_stamp(pd.tile, 5);
        if ((np | nm) == 0u) return;
        __threadfence_block();
        const unsigned long long pre = s_prefix[par];
        const uint64_t base_p = (pre >> 31) + (pd.cta_excl & 0xFFFFu);
        const uint64_t base_m = (pre & kMinusMask) + (pd.cta_excl >> 16);
        for (uint32_t base = 0; base < np || base < nm; base += kListCap) {
            // ---- my hits whose rank falls in [base, base + kListCap) go to the warp lists
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                uint32_t r = s == 0 ? ex_p : ex_m;
#pragma unroll
                for (int w = 0; w < 2; ++w) {
                    uint32_t m = pd.hit[s][w];
                    while (m) {
                        const int b = __ffs(m) - 1;
                        m &= m - 1;
                        if (r - base < (uint32_t)kListCap) list[s][r - base] = (uint16_t)(64 * lane + 32 * w + b);
                        ++r;
                    }
                }
            }
            __syncwarp();
            const uint32_t cp = np > base ? (np - base < (uint32_t)kListCap ? np - base : kListCap) : 0u;
            const uint32_t cm = nm > base ? (nm - base < (uint32_t)kListCap ? nm - base : kListCap) : 0u;
            for (uint32_t k = lane; k < cp; k += 32) {
                const uint32_t pl = list[0][k], t = t_w + pl;
                const uint64_t o = base_p + base + k;
                if (o < a.capacity) {
                    a.pos_plus[o] = t;
                    if (kScore) {
                        const Hit h = extract_window<false>(raw, pl, t, td.L);
                        double x = rs1_canonical(s_tab, h.s0, h.s1, h.valid);
                        if (a.flags & CRP_SCAN_LOGISTIC) x = 1.0 / (1.0 + exp(x));
                        a.packed_plus[o] = h.packed;
                        a.x_plus[o] = x;
                    }
                }
            }
            for (uint32_t k = lane; k < cm; k += 32) {
                const uint32_t pl = list[1][k], t = t_w + pl;
                const uint64_t o = base_m + base + k;
                if (o < a.capacity) {
                    a.pos_minus[o] = t;
                    if (kScore) {
                        const Hit h = extract_window<true>(raw, pl, t, td.L);
                        double x = rs1_canonical(s_tab, h.s0, h.s1, h.valid);
                        if (a.flags & CRP_SCAN_LOGISTIC) x = 1.0 / (1.0 + exp(x));
                        a.packed_minus[o] = h.packed;
                        a.x_minus[o] = x;
                    }
                }
            }
            __syncwarp();
        }
        if (warp == 0 && lane == 0) dbg_stamp(pd.tile, 6);
    };

    Pending cur, nxt;
    uint32_t tile = get_tile(0);
    if (tile >= a.n_tiles) return;
    phase1(0, tile, cur);
    for (uint32_t it = 0;; ++it) {
        tile = get_tile(it + 1);
        const bool has_next = tile < a.n_tiles;
        if (has_next) phase1(it + 1, tile, nxt);
        if (!service) phase2(it, cur);
        if (!has_next) break;
        cur = nxt;
    }
}

// per-segment counts from the inclusive tile prefixes left in `status`
__global__ void k_segment_counts(const unsigned long long *__restrict__ status,
                                 const uint32_t *__restrict__ seg_first_tile,
                                 const uint32_t *__restrict__ seg_tile_count, uint32_t n_seg,
                                 unsigned long long *__restrict__ counts /* [2*n_seg] */) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seg) return;
    const uint32_t f = seg_first_tile[s], c = seg_tile_count[s];
    unsigned long long end = 0, begin = 0;
    if (c > 0) end = status[f + c - 1] & kValMask;
    else if (f > 0) end = status[f - 1] & kValMask;
    if (f > 0) begin = status[f - 1] & kValMask;
    counts[s] = (end >> 31) - (begin >> 31);
    counts[n_seg + s] = (end & kMinusMask) - (begin & kMinusMask);
}

struct RescoreItem {
    uint64_t gpos;     // plane position of token position t
    uint32_t strand;   // '+' or '-'
    uint32_t cls;
};

__device__ __forceinline__ uint32_t window32(const uint32_t *plane, uint64_t start) {
    const uint64_t w = start >> 5;
    return __funnelshift_r(plane[w], plane[w + 1], (uint32_t)(start & 31u));
}

__global__ void k_rescore(const uint32_t *__restrict__ p0, const uint32_t *__restrict__ p1,
                          const uint32_t *__restrict__ lower, const uint32_t *__restrict__ other,
                          const RescoreItem *__restrict__ items, uint64_t n, double *__restrict__ x_out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const RescoreItem it = items[i];
    const bool plus = it.strand == '+';
    const uint64_t start = plus ? it.gpos - 25 : it.gpos - 2;
    // derive() is bitwise, so it commutes with the window extraction
    const Derived d = derive(window32(p0, start), window32(p1, start), window32(lower, start),
                             window32(other, start));
    uint32_t s0, s1, va;
    if (plus) {
        s0 = __brev(d.s0p) >> 2;
        s1 = __brev(d.s1) >> 2;
        va = __brev(d.valid) >> 2;
    } else {
        s0 = d.s0m & 0x3FFFFFFFu;
        s1 = d.s1 & 0x3FFFFFFFu;
        va = d.valid & 0x3FFFFFFFu;
    }
    x_out[i] = rs1_dense(s0, s1, va, (int)(it.cls & 15u), (int)(it.cls >> 4));
}

// ------------------------------------------------------------------ host objects
struct Segment {
    uint32_t token_id;
    const uint8_t *token;
    uint64_t token_len, begin, end;
    uint64_t stage_begin, stage_end;   // token positions copied to the device
    uint64_t gpos0;                    // plane position of token position stage_begin
    uint32_t first_tile, n_tiles;
};

struct crp_genome {
    std::vector<Segment> segs;
    bool committed = false;
    uint64_t n_positions = 0;          // owned positions
    uint64_t g_total = 0;              // plane positions (multiple of 128)
    uint32_t *planes = nullptr;        // 4 planes, each g_total/32 + 8 words
    uint64_t plane_words = 0;
    TileDesc *d_tiles = nullptr;
    uint32_t n_tiles = 0;
    uint32_t *d_seg_first = nullptr, *d_seg_count = nullptr;
    float ms_h2d = 0.f, ms_pack = 0.f;
    const uint32_t *plane(int i) const { return planes + (uint64_t)i * plane_words; }
};

struct crp_result {
    const crp_genome *g = nullptr;
    uint64_t capacity = 0;
    uint64_t n_plus = 0, n_minus = 0;
    bool scored = false;
    uint32_t *pos[2] = {nullptr, nullptr};
    unsigned long long *packed[2] = {nullptr, nullptr};
    double *x[2] = {nullptr, nullptr};
    unsigned long long *status = nullptr;
    unsigned int *ticket = nullptr;
    unsigned long long *d_counts = nullptr;    // [2*n_seg]
    std::vector<uint64_t> seg_plus, seg_minus;
    float ms_scan = 0.f;
};

static int need_ctx() {
    if (!g_ctx.ready) return fail(CRP_ERR_STATE, "crp_init has not been called");
    return 0;
}

// ------------------------------------------------------------------ C ABI
extern "C" {

int crp_abi_version(void) { return CRP_ABI_VERSION; }

int crp_tile_size(void) { return kTile; }

/* debug only (not part of the public header): device buffer of 8 x u64 per tile, or NULL */
int crp_debug_set_tile_times(void *dev_ptr) {
    unsigned long long *p = (unsigned long long *)dev_ptr;
    CUDA_TRY(cudaMemcpyToSymbol(g_dbg_times, &p, sizeof p));
    return 0;
}

const char *crp_last_error(void) { return g_err; }

int crp_device_count(int *count) {
    if (!count) return fail(CRP_ERR_ARG, "count is NULL");
    CUDA_TRY(cudaGetDeviceCount(count));
    return 0;
}

int crp_init(int device) {
    if (g_ctx.ready) {
        if (g_ctx.device == device) return 0;
        return fail(CRP_ERR_STATE, "already initialised on device %d", g_ctx.device);
    }
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(CRP_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                    prop.major, prop.minor);
    CUDA_TRY(cudaStreamCreateWithFlags(&g_ctx.stream, cudaStreamNonBlocking));
    {
        std::vector<double> tab;
        char msg[128];
        if (build_rs1_tables(tab, msg, sizeof msg)) return fail(CRP_ERR_STATE, "%s", msg);
        CUDA_TRY(cudaMalloc(&g_ctx.d_tables, tab.size() * sizeof(double)));
        CUDA_TRY(cudaMemcpy(g_ctx.d_tables, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    g_ctx.device = device;
    g_ctx.sm_count = prop.multiProcessorCount;
    g_ctx.launches = 0;
    g_ctx.ready = true;
    return 0;
}

int crp_shutdown(void) {
    if (!g_ctx.ready) return 0;
    cudaStreamSynchronize(g_ctx.stream);
    cudaStreamDestroy(g_ctx.stream);
    cudaFree(g_ctx.d_tables);
    g_ctx = Context();
    return 0;
}

int crp_launch_count(uint64_t *n) {
    if (!n) return fail(CRP_ERR_ARG, "n is NULL");
    *n = g_ctx.launches;
    return 0;
}

int crp_host_alloc(void **ptr, uint64_t bytes) {
    if (!ptr) return fail(CRP_ERR_ARG, "ptr is NULL");
    if (int rc = need_ctx()) return rc;
    CUDA_TRY(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return 0;
}

int crp_host_free(void *ptr) {
    if (ptr) CUDA_TRY(cudaFreeHost(ptr));
    return 0;
}

int crp_genome_new(crp_genome **g) {
    if (!g) return fail(CRP_ERR_ARG, "g is NULL");
    if (int rc = need_ctx()) return rc;
    *g = new (std::nothrow) crp_genome();
    if (!*g) return fail(CRP_ERR_NOMEM, "out of host memory");
    return 0;
}

int crp_genome_add_segment(crp_genome *g, uint32_t token_id, const uint8_t *token_ascii,
                           uint64_t token_len, uint64_t seg_begin, uint64_t seg_end) {
    if (!g) return fail(CRP_ERR_ARG, "g is NULL");
    if (g->committed) return fail(CRP_ERR_STATE, "genome already committed");
    if (!token_ascii && token_len) return fail(CRP_ERR_ARG, "token_ascii is NULL");
    if (seg_begin > seg_end || seg_end > token_len)
        return fail(CRP_ERR_ARG, "segment [%llu,%llu) outside token of length %llu",
                    (unsigned long long)seg_begin, (unsigned long long)seg_end, (unsigned long long)token_len);
    if (seg_begin % kAlign) return fail(CRP_ERR_ARG, "seg_begin must be a multiple of %u", kAlign);
    if (token_len >= (1ull << 31))
        return fail(CRP_ERR_RANGE, "token of %llu positions exceeds the 31-bit position range",
                    (unsigned long long)token_len);
    Segment s{};
    s.token_id = token_id;
    s.token = token_ascii;
    s.token_len = token_len;
    s.begin = seg_begin;
    s.end = seg_end;
    g->segs.push_back(s);
    return 0;
}

int crp_genome_num_segments(const crp_genome *g, uint32_t *n) {
    if (!g || !n) return fail(CRP_ERR_ARG, "NULL argument");
    *n = (uint32_t)g->segs.size();
    return 0;
}

int crp_genome_num_positions(const crp_genome *g, uint64_t *n) {
    if (!g || !n) return fail(CRP_ERR_ARG, "NULL argument");
    *n = g->n_positions;
    return 0;
}

int crp_genome_commit(crp_genome *g) {
    if (!g) return fail(CRP_ERR_ARG, "g is NULL");
    if (int rc = need_ctx()) return rc;
    if (g->committed) return fail(CRP_ERR_STATE, "genome already committed");
    cudaStream_t st = g_ctx.stream;

    // ---- layout: every segment gets [128 positions of left context][data][>=32 right context]
    uint64_t gp = 0;
    std::vector<TileDesc> tiles;
    std::vector<uint32_t> seg_first, seg_count;
    g->n_positions = 0;
    for (Segment &s : g->segs) {
        s.stage_begin = s.begin >= kAlign ? s.begin - kAlign : 0;
        s.stage_end = s.end + 64 < s.token_len ? s.end + 64 : s.token_len;
        // plane position of token position s.begin is a multiple of 128, with 128 positions before it
        const uint64_t g_begin = gp + kAlign;
        s.gpos0 = g_begin - (s.begin - s.stage_begin);
        const uint64_t g_end = g_begin + (s.stage_end - s.begin);
        gp = (g_end + 64 + kAlign - 1) / kAlign * kAlign;
        s.first_tile = (uint32_t)tiles.size();
        for (uint64_t t = s.begin; t < s.end; t += kTile) {
            TileDesc td;
            const uint64_t gw = (g_begin + (t - s.begin)) >> 5;
            if (gw >= (1ull << 32)) return fail(CRP_ERR_RANGE, "shard too large for 32-bit plane word index");
            td.gword = (uint32_t)gw;
            td.t_start = (uint32_t)t;
            td.L = (uint32_t)s.token_len;
            td.n = (uint32_t)((s.end - t) < (uint64_t)kTile ? (s.end - t) : (uint64_t)kTile);
            tiles.push_back(td);
        }
        s.n_tiles = (uint32_t)tiles.size() - s.first_tile;
        seg_first.push_back(s.first_tile);
        seg_count.push_back(s.n_tiles);
        g->n_positions += s.end - s.begin;
    }
    g->g_total = gp + kAlign;
    g->plane_words = g->g_total / 32 + kTile / 32 + 128;   // any tile may read its full extent
    g->n_tiles = (uint32_t)tiles.size();

    uint8_t *d_ascii = nullptr;
    CUDA_TRY(cudaMalloc(&d_ascii, g->g_total));
    if (cudaMalloc(&g->planes, 4 * g->plane_words * sizeof(uint32_t)) != cudaSuccess) {
        cudaFree(d_ascii);
        return fail(CRP_ERR_NOMEM, "cudaMalloc of %llu plane bytes failed",
                    (unsigned long long)(4 * g->plane_words * sizeof(uint32_t)));
    }
    cudaEvent_t e0, e1, e2;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    CUDA_TRY(cudaEventCreate(&e2));
    CUDA_TRY(cudaEventRecord(e0, st));
    CUDA_TRY(cudaMemsetAsync(d_ascii, 0, g->g_total, st));
    CUDA_TRY(cudaMemsetAsync(g->planes, 0, 4 * g->plane_words * sizeof(uint32_t), st));
    for (const Segment &s : g->segs) {
        if (s.stage_end > s.stage_begin)
            CUDA_TRY(cudaMemcpyAsync(d_ascii + s.gpos0, s.token + s.stage_begin, s.stage_end - s.stage_begin,
                                     cudaMemcpyHostToDevice, st));
    }
    CUDA_TRY(cudaEventRecord(e1, st));
    const uint64_t n_words = g->g_total / 32;
    if (n_words) {
        int blocks = (int)((n_words + 255) / 256 < (uint64_t)g_ctx.sm_count * 8 ? (n_words + 255) / 256
                                                                                 : (uint64_t)g_ctx.sm_count * 8);
        uint32_t *P = g->planes;
        k_pack<<<blocks, 256, 0, st>>>((const uint4 *)d_ascii, n_words, P, P + g->plane_words,
                                       P + 2 * g->plane_words, P + 3 * g->plane_words);
        g_ctx.launches++;
        CUDA_TRY(cudaGetLastError());
    }
    CUDA_TRY(cudaEventRecord(e2, st));
    if (g->n_tiles) {
        CUDA_TRY(cudaMalloc(&g->d_tiles, g->n_tiles * sizeof(TileDesc)));
        CUDA_TRY(cudaMemcpyAsync(g->d_tiles, tiles.data(), g->n_tiles * sizeof(TileDesc), cudaMemcpyHostToDevice, st));
    }
    if (!g->segs.empty()) {
        const size_t nb = g->segs.size() * sizeof(uint32_t);
        CUDA_TRY(cudaMalloc(&g->d_seg_first, nb));
        CUDA_TRY(cudaMalloc(&g->d_seg_count, nb));
        CUDA_TRY(cudaMemcpyAsync(g->d_seg_first, seg_first.data(), nb, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(g->d_seg_count, seg_count.data(), nb, cudaMemcpyHostToDevice, st));
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaEventElapsedTime(&g->ms_h2d, e0, e1));
    CUDA_TRY(cudaEventElapsedTime(&g->ms_pack, e1, e2));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaEventDestroy(e2);
    CUDA_TRY(cudaFree(d_ascii));
    for (Segment &s : g->segs) s.token = nullptr;   // host tokens may be released now
    g->committed = true;
    return 0;
}

int crp_genome_timing(const crp_genome *g, float *ms_h2d, float *ms_pack) {
    if (!g) return fail(CRP_ERR_ARG, "g is NULL");
    if (ms_h2d) *ms_h2d = g->ms_h2d;
    if (ms_pack) *ms_pack = g->ms_pack;
    return 0;
}

int crp_genome_free(crp_genome *g) {
    if (!g) return 0;
    cudaFree(g->planes);
    cudaFree(g->d_tiles);
    cudaFree(g->d_seg_first);
    cudaFree(g->d_seg_count);
    delete g;
    return 0;
}

static void free_streams(crp_result *r) {
    for (int s = 0; s < 2; ++s) {
        cudaFree(r->pos[s]);
        cudaFree(r->packed[s]);
        cudaFree(r->x[s]);
        r->pos[s] = nullptr;
        r->packed[s] = nullptr;
        r->x[s] = nullptr;
    }
}

static int alloc_streams(crp_result *r, uint64_t cap, bool scored) {
    r->capacity = cap;
    const uint64_t n = cap ? cap : 1;
    for (int s = 0; s < 2; ++s) {
        if (cudaMalloc(&r->pos[s], n * sizeof(uint32_t)) != cudaSuccess) goto oom;
        if (scored) {
            if (cudaMalloc(&r->packed[s], n * sizeof(unsigned long long)) != cudaSuccess) goto oom;
            if (cudaMalloc(&r->x[s], n * sizeof(double)) != cudaSuccess) goto oom;
        }
    }
    return 0;
oom:
    cudaGetLastError();
    free_streams(r);
    return fail(CRP_ERR_NOMEM, "cudaMalloc of candidate streams (%llu entries per strand) failed",
                (unsigned long long)cap);
}

static int launch_scan(const crp_genome *g, crp_result *r, int guide_len, uint32_t flags, cudaEvent_t e0,
                       cudaEvent_t e1) {
    cudaStream_t st = g_ctx.stream;
    ScanArgs a;
    a.p0 = g->plane(0);
    a.p1 = g->plane(1);
    a.lower = g->plane(2);
    a.other = g->plane(3);
    a.tiles = g->d_tiles;
    a.n_tiles = g->n_tiles;
    a.guide_len = guide_len;
    a.flags = flags;
    a.status = r->status;
    a.ticket = r->ticket;
    a.tables = g_ctx.d_tables;
    a.capacity = r->capacity;
    a.pos_plus = r->pos[0];
    a.pos_minus = r->pos[1];
    a.packed_plus = r->packed[0];
    a.packed_minus = r->packed[1];
    a.x_plus = r->x[0];
    a.x_minus = r->x[1];
    CUDA_TRY(cudaEventRecord(e0, st));
    if (g->n_tiles) {
        CUDA_TRY(cudaMemsetAsync(r->status, 0, (size_t)g->n_tiles * sizeof(unsigned long long), st));
        CUDA_TRY(cudaMemsetAsync(r->ticket, 0, sizeof(unsigned int), st));
        const void *fn = r->scored ? (const void *)k_scan_score<true> : (const void *)k_scan_score<false>;
        int per_sm = 0;
        CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kCtaThreads, kSmemBytes));
        if (per_sm < 1) return fail(CRP_ERR_CUDA, "scan kernel does not fit on an SM");
        // persistent grid, every CTA resident (the look-back spins on predecessors)
        uint64_t blocks = (uint64_t)g_ctx.sm_count * per_sm;
        if (blocks > g->n_tiles) blocks = g->n_tiles;
        void *params[] = {(void *)&a};
        CUDA_TRY(cudaLaunchCooperativeKernel(fn, dim3((unsigned)blocks), dim3(kCtaThreads), params, kSmemBytes, st));
        g_ctx.launches++;
        CUDA_TRY(cudaGetLastError());
    }
    const uint32_t n_seg = (uint32_t)g->segs.size();
    if (n_seg) {
        if (g->n_tiles) {
            k_segment_counts<<<(n_seg + 127) / 128, 128, 0, st>>>(r->status, g->d_seg_first, g->d_seg_count, n_seg,
                                                                 r->d_counts);
            g_ctx.launches++;
            CUDA_TRY(cudaGetLastError());
        } else {
            CUDA_TRY(cudaMemsetAsync(r->d_counts, 0, 2 * (size_t)n_seg * sizeof(unsigned long long), st));
        }
    }
    CUDA_TRY(cudaEventRecord(e1, st));
    return 0;
}

int crp_scan_score(crp_genome *g, int guide_len, uint32_t flags, crp_result **res) {
    if (!g || !res) return fail(CRP_ERR_ARG, "NULL argument");
    if (int rc = need_ctx()) return rc;
    if (!g->committed) return fail(CRP_ERR_STATE, "genome not committed");
    if (guide_len < 1 || guide_len > 1000000) return fail(CRP_ERR_ARG, "guide_len %d out of range", guide_len);
    crp_result *r = new (std::nothrow) crp_result();
    if (!r) return fail(CRP_ERR_NOMEM, "out of host memory");
    r->g = g;
    r->scored = guide_len == 20 && !(flags & CRP_SCAN_NO_SCORE);
    const uint32_t n_seg = (uint32_t)g->segs.size();
    cudaStream_t st = g_ctx.stream;
    int rc = 0;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    std::vector<unsigned long long> counts(2 * (size_t)n_seg);
    auto bail = [&](int code) {
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        crp_result_free(r);
        return code;
    };
    if (cudaMalloc(&r->status, ((size_t)g->n_tiles + 1) * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMalloc(&r->ticket, sizeof(unsigned int)) != cudaSuccess ||
        cudaMalloc(&r->d_counts, (2 * (size_t)n_seg + 1) * sizeof(unsigned long long)) != cudaSuccess)
        return bail(fail(CRP_ERR_NOMEM, "cudaMalloc of scan state failed"));
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess)
        return bail(fail(CRP_ERR_CUDA, "cudaEventCreate failed"));
    // First guess of the per-strand capacity: 1/8 candidate per position (GC 70 %
    // upper-case sequence gives 0.1225); a second pass with the exact counts
    // follows if it was too small.
    {
        uint64_t cap = g->n_positions / 8 + 4096;
        if ((rc = alloc_streams(r, cap, r->scored))) return bail(rc);
    }
    for (int attempt = 0; attempt < 2; ++attempt) {
        if ((rc = launch_scan(g, r, guide_len, flags, e0, e1))) return bail(rc);
        if (n_seg)
            if (cudaError_t e = cudaMemcpyAsync(counts.data(), r->d_counts, counts.size() * sizeof(unsigned long long),
                                                cudaMemcpyDeviceToHost, st))
                return bail(fail(CRP_ERR_CUDA, "D2H of segment counts failed: %s", cudaGetErrorString(e)));
        if (cudaError_t e = cudaStreamSynchronize(st))
            return bail(fail(CRP_ERR_CUDA, "scan kernels failed: %s", cudaGetErrorString(e)));
        r->n_plus = r->n_minus = 0;
        r->seg_plus.assign(n_seg, 0);
        r->seg_minus.assign(n_seg, 0);
        for (uint32_t s = 0; s < n_seg; ++s) {
            r->seg_plus[s] = counts[s];
            r->seg_minus[s] = counts[n_seg + s];
            r->n_plus += counts[s];
            r->n_minus += counts[n_seg + s];
        }
        const uint64_t need = r->n_plus > r->n_minus ? r->n_plus : r->n_minus;
        if (need <= r->capacity) break;
        if (attempt == 1) return bail(fail(CRP_ERR_STATE, "candidate streams overflowed twice"));
        free_streams(r);
        if ((rc = alloc_streams(r, need, r->scored))) return bail(rc);
    }
    cudaEventElapsedTime(&r->ms_scan, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *res = r;
    return 0;
}

int crp_result_totals(const crp_result *res, uint64_t *n_plus, uint64_t *n_minus) {
    if (!res) return fail(CRP_ERR_ARG, "res is NULL");
    if (n_plus) *n_plus = res->n_plus;
    if (n_minus) *n_minus = res->n_minus;
    return 0;
}

int crp_result_segment_counts(const crp_result *res, uint64_t *n_plus, uint64_t *n_minus) {
    if (!res) return fail(CRP_ERR_ARG, "res is NULL");
    for (size_t s = 0; s < res->seg_plus.size(); ++s) {
        if (n_plus) n_plus[s] = res->seg_plus[s];
        if (n_minus) n_minus[s] = res->seg_minus[s];
    }
    return 0;
}

int crp_result_device_counts(const crp_result *res, void **dev_ptr) {
    if (!res || !dev_ptr) return fail(CRP_ERR_ARG, "NULL argument");
    *dev_ptr = res->d_counts;
    return 0;
}

int crp_result_fetch(const crp_result *res, char strand, uint64_t first, uint64_t count, uint32_t *pos,
                     uint64_t *packed, double *x) {
    if (!res) return fail(CRP_ERR_ARG, "res is NULL");
    if (int rc = need_ctx()) return rc;
    if (strand != '+' && strand != '-') return fail(CRP_ERR_ARG, "strand must be '+' or '-'");
    const int s = strand == '+' ? 0 : 1;
    const uint64_t total = s == 0 ? res->n_plus : res->n_minus;
    if (first > total || count > total - first)
        return fail(CRP_ERR_ARG, "range [%llu,+%llu) outside stream of %llu candidates", (unsigned long long)first,
                    (unsigned long long)count, (unsigned long long)total);
    if ((packed || x) && !res->scored && count)
        return fail(CRP_ERR_STATE, "this result carries positions only (guide_len != 20 or CRP_SCAN_NO_SCORE)");
    cudaStream_t st = g_ctx.stream;
    if (count) {
        if (pos) CUDA_TRY(cudaMemcpyAsync(pos, res->pos[s] + first, count * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        if (packed)
            CUDA_TRY(cudaMemcpyAsync(packed, res->packed[s] + first, count * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        if (x) CUDA_TRY(cudaMemcpyAsync(x, res->x[s] + first, count * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

int crp_result_timing(const crp_result *res, float *ms_scan) {
    if (!res) return fail(CRP_ERR_ARG, "res is NULL");
    if (ms_scan) *ms_scan = res->ms_scan;
    return 0;
}

int crp_result_free(crp_result *r) {
    if (!r) return 0;
    free_streams(r);
    cudaFree(r->status);
    cudaFree(r->ticket);
    cudaFree(r->d_counts);
    delete r;
    return 0;
}

int crp_rescore(const crp_genome *g, uint64_t n, const uint32_t *segment, const uint32_t *t, const char *strand,
                const uint8_t *cls, double *x_out) {
    if (!g) return fail(CRP_ERR_ARG, "g is NULL");
    if (int rc = need_ctx()) return rc;
    if (!g->committed) return fail(CRP_ERR_STATE, "genome not committed");
    if (n == 0) return 0;
    if (!segment || !t || !strand || !cls || !x_out) return fail(CRP_ERR_ARG, "NULL argument");
    std::vector<RescoreItem> items(n);
    for (uint64_t i = 0; i < n; ++i) {
        if (segment[i] >= g->segs.size()) return fail(CRP_ERR_ARG, "item %llu: bad segment", (unsigned long long)i);
        const Segment &s = g->segs[segment[i]];
        if (t[i] < s.begin || t[i] >= s.end)
            return fail(CRP_ERR_ARG, "item %llu: t=%u outside segment", (unsigned long long)i, t[i]);
        if (strand[i] != '+' && strand[i] != '-') return fail(CRP_ERR_ARG, "item %llu: bad strand", (unsigned long long)i);
        if ((cls[i] & 15u) > CRP_CLASS_SINGLE || (cls[i] >> 4) > CRP_CLASS_SINGLE) return fail(CRP_ERR_ARG, "item %llu: bad class", (unsigned long long)i);
        items[i].gpos = s.gpos0 + (t[i] - s.stage_begin);
        items[i].strand = (uint32_t)strand[i];
        items[i].cls = cls[i];
    }
    cudaStream_t st = g_ctx.stream;
    RescoreItem *d_items = nullptr;
    double *d_x = nullptr;
    CUDA_TRY(cudaMalloc(&d_items, n * sizeof(RescoreItem)));
    if (cudaMalloc(&d_x, n * sizeof(double)) != cudaSuccess) {
        cudaFree(d_items);
        return fail(CRP_ERR_NOMEM, "cudaMalloc failed");
    }
    int rc = 0;
    do {
        if (cudaMemcpyAsync(d_items, items.data(), n * sizeof(RescoreItem), cudaMemcpyHostToDevice, st) != cudaSuccess) {
            rc = fail(CRP_ERR_CUDA, "H2D of rescore items failed");
            break;
        }
        k_rescore<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(g->plane(0), g->plane(1), g->plane(2), g->plane(3),
                                                              d_items, n, d_x);
        g_ctx.launches++;
        if (cudaMemcpyAsync(x_out, d_x, n * sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess) {
            rc = fail(CRP_ERR_CUDA, "rescore failed: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
    } while (0);
    cudaFree(d_items);
    cudaFree(d_x);
    return rc;
}

}  // extern "C"
