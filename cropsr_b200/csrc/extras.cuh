// Opt-in per-candidate side outputs (north_star subsystems 3 and 4).  None of this exists in
// the reference at this commit (SURVEY.md "five facts" 2): the reference parses -L / -g and
// never uses them, so these outputs NEVER enter the parity CSV.  Semantics are pinned by
// oracle/extras_oracle.py instead:
//   gc        G+C count of the 20-base protospacer = scored 30-mer[5:25]; the one hint the
//             reference gives is cropsr_functions.py:174-179 (`sequence[5:-5]`, "< 10 -> low")
//   flags     bit0 poly-T (TTTT inside the protospacer: Pol III terminator)
//             bit1 homopolymer run >= 5      bit2 low GC (gc < 10)
//             bit3 the protospacer holds a base that does not score (N, IUPAC, quote, ...)
//   run       longest homopolymer run in the protospacer
//   cut       cut site = end_pos - 3 of the CSV row (CROPSR.py:155-158)
//   flank     [cut - L, cut + L) clipped to the token: the window handed to primer design
//   feature   index of the innermost annotated interval containing the cut site, or -1
#pragma once
#include "scan.cuh"

static constexpr uint32_t kProto = 0x01FFFFE0u;     // bits 5..24 of the 30-mer

struct ExtrasArgs {
    const uint4 *records;
    const uint32_t *pos;          // t of the candidates of one segment, one strand
    uint64_t n;
    uint32_t first_tile, seg_begin, L, flank;
    int minus;
    uint8_t *gc, *flags, *run;
    uint32_t *cut, *flank_lo, *flank_hi;
};

__global__ void k_extras(const ExtrasArgs a) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const uint32_t t = a.pos[i], rel = t - a.seg_begin;
    const uint4 *rec = a.records + (size_t)(a.first_tile + rel / kTile) * kRecWords;
    const Window w = a.minus ? extract_window<true>(rec, rel % kTile + kWinBiasMinus, t, a.L)
                             : extract_window<false>(rec, rel % kTile + kWinBiasPlus, t, a.L);
    const uint32_t v = w.valid & kProto, s0 = w.s0, s1 = w.s1;
    const uint32_t gc = __popc(s1 & v);                                    // C = 2, G = 3: high code bit
    const uint32_t isT = ~s1 & s0 & v;
    const uint32_t t4 = isT & (isT >> 1) & (isT >> 2) & (isT >> 3);        // bit q: T at q..q+3 (all inside the mask)
    // same[q]: bases q and q+1 both score and are equal
    uint32_t same = ~(s0 ^ (s0 >> 1)) & ~(s1 ^ (s1 >> 1)) & v & (v >> 1);
    uint32_t run = v ? 1u : 0u;
    while (same) {
        same &= same >> 1;
        ++run;
    }
    uint32_t f = 0;
    if (t4) f |= 1u;
    if (run >= 5) f |= 2u;
    if (gc < 10) f |= 4u;
    if (v != kProto) f |= 8u;
    const uint32_t cut = a.minus ? t : t - 3u;
    if (a.gc) a.gc[i] = (uint8_t)gc;
    if (a.flags) a.flags[i] = (uint8_t)f;
    if (a.run) a.run[i] = (uint8_t)run;
    if (a.cut) a.cut[i] = cut;
    if (a.flank_lo) a.flank_lo[i] = cut > a.flank ? cut - a.flank : 0u;
    if (a.flank_hi) a.flank_hi[i] = cut + a.flank < a.L ? cut + a.flank : a.L;
}

// Intervals sorted by start, inclusive [start, end]; maxend[j] = max(end[0..j]).  One thread
// per candidate: binary search of the last start <= cut, then walk back while an earlier
// interval can still contain the cut site (nested / overlapping features).
__global__ void k_annotate(const uint32_t *__restrict__ pos, uint64_t n, int minus, const uint32_t *__restrict__ start,
                           const uint32_t *__restrict__ end, const uint32_t *__restrict__ maxend, uint32_t n_iv,
                           int32_t *__restrict__ feature) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t cut = minus ? pos[i] : pos[i] - 3u;
    uint32_t lo = 0, hi = n_iv;                       // first interval with start > cut
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(start + mid) <= cut) lo = mid + 1;
        else hi = mid;
    }
    int32_t found = -1;
    for (int64_t j = (int64_t)lo - 1; j >= 0 && __ldg(maxend + j) >= cut; --j)
        if (__ldg(end + j) >= cut) {
            found = (int32_t)j;
            break;
        }
    feature[i] = found;
}
