// Primer enumeration over windows of a packed genome: the part of the
// reference's primer designer that needs no aligner, /root/reference/prmrdsgn2.py:
//   get_primers      :115-124   every substring [i, i + N) of the first e bases of the fragment,
//                               i = 0 .. e-1, N = s+1 .. l  (i-major), and the same on the
//                               reverse complement (create_reverse_complement :104-112)
//   Primer           :76-95     GC % = 100 * (gc / N),  Tm = 64.9 + 41 * (gc - 16.4) / N   (N >= 13)
//   filter_primers   :127-137   drop GC % < M, > X, Tm < m, > x
//   main             :260-266   pairs forward x reverse with math.isclose(Tm_f, Tm_r, abs_tol=D)
// GC % and Tm are functions of (N, gc) alone, so the host evaluates them ONCE per class with the
// reference's own double arithmetic (build_primer_classes: one IEEE operation per Python
// operator) and the device only counts: per window a warp builds the G/C bit string of both ends
// from the tile records, histograms the passing reverse primers by Tm rank, and sums for every
// passing forward primer the reverse primers whose Tm is close -- a contiguous rank range.
// Opt-in side output; it never touches the parity CSV (CROPSR.py never calls prmrdsgn2).
#pragma once
#include <math.h>

#include <algorithm>
#include <vector>

#include "scan.cuh"

static constexpr int kPrimerMaxLen = 32;          // l <= 32: a primer's G/C bits fit one funnel shift
static constexpr int kPrimerMaxClasses = 640;     // (l - s) * (l + 1) classes (N, gc)
static constexpr int kPrimerMaxRegion = 1024;     // e + l bases per end, one mask word per lane

struct PrimerClasses {                            // class c = (N - s - 1) * (l + 1) + gc
    uint16_t n_classes, n_len, min_len, stride;   // n_len = l - s lengths, min_len = s + 1, stride = l + 1
    uint16_t rank[kPrimerMaxClasses];             // position of the class in ascending-Tm order (0xFFFF: filtered out)
    uint16_t lo[kPrimerMaxClasses], hi[kPrimerMaxClasses];   // ranks of the passing classes whose Tm is close: [lo, hi)
};

struct PrimerArgs {
    const uint4 *records;
    const PrimerClasses *cls;                     // device copy
    const uint32_t *first_tile, *seg_begin;       // per window: first tile record / first owned position of its segment
    const uint32_t *lo, *hi;                      // per window: token positions [lo, hi)
    uint64_t n;
    uint32_t e;
    uint32_t *n_fwd, *n_rev;
    unsigned long long *n_pairs;
    uint16_t *first;                              // 4 per window: forward (i, N), reverse (i, N); 0xFFFF: no pair
    uint8_t *status;                              // 0 ok, 1 window shorter than e + l
};

// host: the class table for one parameter set, with the reference's arithmetic
static int build_primer_classes(int s, int l, double m, double x, double M, double X, double D, PrimerClasses *pc,
                                char *err, size_t errlen) {
    if (s < 12 || l <= s || l > kPrimerMaxLen || (l - s) * (l + 1) > kPrimerMaxClasses) {
        snprintf(err, errlen, "primer lengths s=%d l=%d unsupported: need 12 <= s < l <= %d", s, l, kPrimerMaxLen);
        return -1;
    }
    pc->n_len = (uint16_t)(l - s);
    pc->min_len = (uint16_t)(s + 1);
    pc->stride = (uint16_t)(l + 1);
    pc->n_classes = (uint16_t)((l - s) * (l + 1));
    std::vector<double> tm(pc->n_classes, 0.0);
    std::vector<int> pass;                       // passing classes
    for (int N = s + 1; N <= l; ++N)
        for (int gc = 0; gc <= l; ++gc) {
            const int c = (N - s - 1) * (l + 1) + gc;
            pc->rank[c] = 0xFFFFu;
            pc->lo[c] = pc->hi[c] = 0;
            if (gc > N) continue;
            volatile double frac = (double)gc / (double)N;                 // prmrdsgn2.py:79
            volatile double gcp = 100.0 * frac;                            // :80
            volatile double d0 = (double)gc - 16.4, d1 = 41.0 * d0, d2 = d1 / (double)N;   // :94
            volatile double t = 64.9 + d2;
            tm[c] = t;
            if (!(gcp < M || gcp > X || t < m || t > x)) pass.push_back(c);   // :133
        }
    std::stable_sort(pass.begin(), pass.end(), [&](int a, int b) { return tm[a] < tm[b]; });
    for (size_t r = 0; r < pass.size(); ++r) pc->rank[pass[r]] = (uint16_t)r;
    auto close = [&](double a, double b) {       // math.isclose(a, b, abs_tol=D), rel_tol = 1e-9
        const double diff = fabs(a - b);
        return diff <= fmax(1e-9 * fmax(fabs(a), fabs(b)), D);
    };
    for (int cf : pass) {
        size_t lo = 0, hi = 0;
        bool seen = false, ended = false;
        for (size_t r = 0; r < pass.size(); ++r) {
            const bool ok = close(tm[cf], tm[pass[r]]);
            if (ok && ended) {
                snprintf(err, errlen, "Tm-close classes are not contiguous in Tm order (D too small?)");
                return -1;
            }
            if (ok && !seen) {
                seen = true;
                lo = r;
            }
            if (ok) hi = r + 1;
            if (!ok && seen) ended = true;
        }
        pc->lo[cf] = (uint16_t)lo;
        pc->hi[cf] = (uint16_t)hi;
    }
    return 0;
}

// G/C-ness (case-insensitive G or C; "other" bytes never count) of token position p of a segment
__device__ __forceinline__ uint32_t primer_gc_word(const uint4 *__restrict__ records, uint32_t first_tile,
                                                   uint32_t rel) {        // the 32 positions from rel (32-aligned) on
    const uint4 w = records[(size_t)(first_tile + rel / kTile) * kRecWords + 2 + (rel % kTile) / 32];
    return w.y & ~w.w;
}

// One warp per window.
__global__ void __launch_bounds__(256)
k_primers(const PrimerArgs a) {
    __shared__ PrimerClasses cls;
    __shared__ uint32_t s_fwd[8][kPrimerMaxRegion / 32 + 2], s_rev[8][kPrimerMaxRegion / 32 + 2];
    __shared__ uint32_t s_hist[8][kPrimerMaxClasses + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < sizeof(PrimerClasses) / 4; i += blockDim.x)
        reinterpret_cast<uint32_t *>(&cls)[i] = reinterpret_cast<const uint32_t *>(a.cls)[i];
    __syncthreads();
    const uint32_t e = a.e, n_len = cls.n_len, min_len = cls.min_len, l = min_len + n_len - 1;
    const uint32_t region = e + l;                         // bases of each end a primer can touch
    const uint32_t n_words = (region + 31) / 32;
    const uint32_t items = e * n_len;                      // primers per end, reference order: i-major
    uint32_t *const fw = s_fwd[warp], *const rv = s_rev[warp], *const hist = s_hist[warp];
    for (uint64_t wi = (uint64_t)blockIdx.x * 8 + warp; wi < a.n; wi += (uint64_t)gridDim.x * 8) {
        const uint32_t lo = a.lo[wi], hi = a.hi[wi], n = hi - lo;
        if (hi < lo || n < region) {
            if (lane == 0) {
                a.status[wi] = 1;
                a.n_fwd[wi] = a.n_rev[wi] = 0;
                a.n_pairs[wi] = 0;
                for (int k = 0; k < 4; ++k) a.first[4 * wi + k] = 0xFFFFu;
            }
            continue;
        }
        const uint32_t ft = a.first_tile[wi], sb = a.seg_begin[wi];
        // ---- G/C bit strings: fw bit k = fragment base k; rv bit k = base k of the reverse complement
        //      = fragment base n-1-k (the complement keeps G/C-ness), k < region
        //      Only words that hold window positions are read (the window may end on the last
        //      position of the segment's last tile); word n_words is the zero high half of the last funnel shift.
        const uint32_t end_rel = hi - sb;                                      // one past the last position, segment-relative
        for (uint32_t w = lane; w <= n_words; w += 32) {
            uint32_t f = 0, r = 0;
            if (w < n_words) {
                const uint32_t p = lo - sb + 32 * w, pa = p & ~31u;            // forward: positions p .. p + 31
                const uint32_t a0 = primer_gc_word(a.records, ft, pa);
                const uint32_t a1 = pa + 32 < end_rel ? primer_gc_word(a.records, ft, pa + 32) : 0u;
                f = __funnelshift_r(a0, a1, p & 31u);
                const uint32_t q = end_rel - 1 - 32 * w;                       // reverse: positions q - 31 .. q, read backwards
                if (q >= 31) {
                    const uint32_t first = q - 31, fa = first & ~31u;
                    const uint32_t b0 = primer_gc_word(a.records, ft, fa);
                    const uint32_t b1 = (first & 31u) ? primer_gc_word(a.records, ft, fa + 32) : 0u;
                    r = __brev(__funnelshift_r(b0, b1, first & 31u));
                } else {                                                       // fewer than 32 positions left of q in the segment
                    r = __brev(primer_gc_word(a.records, ft, 0u) << (31 - q));
                }
            }
            fw[w] = f;
            rv[w] = r;
        }
        for (uint32_t c = lane; c <= cls.n_classes; c += 32) hist[c] = 0;
        __syncwarp();
        // ---- passing reverse primers by Tm rank
        for (uint32_t it = lane; it < items; it += 32) {
            const uint32_t i = it / n_len, N = min_len + it % n_len;
            const uint32_t bits = __funnelshift_r(rv[i >> 5], rv[(i >> 5) + 1], i & 31u) & (0xFFFFFFFFu >> (32 - N));
            const uint32_t r = cls.rank[(N - min_len) * cls.stride + __popc(bits)];
            if (r != 0xFFFFu) atomicAdd(&hist[r + 1], 1u);
        }
        __syncwarp();
        // ---- inclusive scan in place: hist[r] = passing reverse primers of rank < r
        uint32_t carry = 0;
        for (uint32_t base = 0; base <= cls.n_classes; base += 32) {
            const uint32_t idx = base + lane;
            uint32_t v = idx <= cls.n_classes ? hist[idx] : 0u;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, v, o);
                if (lane >= o) v += u;
            }
            v += carry;
            if (idx <= cls.n_classes) hist[idx] = v;
            carry = __shfl_sync(0xFFFFFFFFu, v, 31);
        }
        __syncwarp();
        const uint32_t n_rev = carry;
        // ---- forward primers
        uint32_t n_fwd = 0, best = 0xFFFFFFFFu;
        unsigned long long pairs = 0;
        for (uint32_t it = lane; it < items; it += 32) {
            const uint32_t i = it / n_len, N = min_len + it % n_len;
            const uint32_t bits = __funnelshift_r(fw[i >> 5], fw[(i >> 5) + 1], i & 31u) & (0xFFFFFFFFu >> (32 - N));
            const uint32_t c = (N - min_len) * cls.stride + __popc(bits);
            if (cls.rank[c] != 0xFFFFu) {
                ++n_fwd;
                const uint32_t cnt = hist[cls.hi[c]] - hist[cls.lo[c]];
                pairs += cnt;
                if (cnt && it < best) best = it;
            }
        }
        n_fwd = __reduce_add_sync(0xFFFFFFFFu, n_fwd);
        best = __reduce_min_sync(0xFFFFFFFFu, best);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) pairs += __shfl_xor_sync(0xFFFFFFFFu, pairs, o);
        // ---- the first pair of itertools.product(forward, reverse): first forward primer with a partner,
        //      then the first reverse primer (reference order) whose Tm is close to it
        uint32_t best_r = 0xFFFFFFFFu;
        uint32_t fN = 0, fi = 0;
        if (best != 0xFFFFFFFFu) {
            fi = best / n_len;
            fN = min_len + best % n_len;
            const uint32_t bits = __funnelshift_r(fw[fi >> 5], fw[(fi >> 5) + 1], fi & 31u) & (0xFFFFFFFFu >> (32 - fN));
            const uint32_t c = (fN - min_len) * cls.stride + __popc(bits);
            const uint32_t rlo = cls.lo[c], rhi = cls.hi[c];
            for (uint32_t it = lane; it < items && best_r == 0xFFFFFFFFu; it += 32) {
                const uint32_t i = it / n_len, N = min_len + it % n_len;
                const uint32_t b = __funnelshift_r(rv[i >> 5], rv[(i >> 5) + 1], i & 31u) & (0xFFFFFFFFu >> (32 - N));
                const uint32_t r = cls.rank[(N - min_len) * cls.stride + __popc(b)];
                if (r != 0xFFFFu && r >= rlo && r < rhi) best_r = it;
            }
            best_r = __reduce_min_sync(0xFFFFFFFFu, best_r);
        }
        if (lane == 0) {
            a.status[wi] = 0;
            a.n_fwd[wi] = n_fwd;
            a.n_rev[wi] = n_rev;
            a.n_pairs[wi] = pairs;
            const bool has = best != 0xFFFFFFFFu;
            a.first[4 * wi + 0] = has ? (uint16_t)fi : 0xFFFFu;
            a.first[4 * wi + 1] = has ? (uint16_t)fN : 0xFFFFu;
            a.first[4 * wi + 2] = has ? (uint16_t)(best_r / n_len) : 0xFFFFu;
            a.first[4 * wi + 3] = has ? (uint16_t)(min_len + best_r % n_len) : 0xFFFFu;
        }
        __syncwarp();
    }
}
