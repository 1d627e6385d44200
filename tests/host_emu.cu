// TEST INFRASTRUCTURE: the scoring functions of the scan kernel, run on the CPU from the SAME source.
//
// rs1.cuh / scan.cuh mark the per-candidate functions __host__ __device__ (the intrinsics they use have
// host twins), so this program can check, without a GPU,
//   A  the table-driven lane sums (rs1_canonical: perfect hashes, swizzled slots, exponent-bit tail
//      multiplies, tables built by build_rs1_tables exactly as crp_init builds them) against the dense
//      column-by-column replay of OpenBLAS' canonical lane order (rs1_dense), bit for bit;
//   B  the emit loop's score_hit (class masks straight from the shifted planes, '+' windows shifted by
//      ws - 2, flags through multiplies and bit selects) against the generic extract_window +
//      rs1_canonical pair that the side outputs and the rescore kernel use: packed word and x, bit for bit,
//      on random staged records with every byte class, both strands, truncated windows included;
//   C  the PAM tests: the emit phase's tile_hits (bounds of CROPSR.py:419 / :430 and ownership as range masks)
//      against a byte-by-byte reading of the token, and k_pack's pack_word_hits -- the counts a tile header
//      holds -- against the same reading on every tile that is not at an end of its token for the guide
//      length in question (the tiles the count phase does not count again), for guides of 1 .. 100,000.
// Built and run by tests/test_host_logic.py (nvcc, no GPU needed).  Nothing here is part of the library.
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../cropsr_b200/csrc/scan.cuh"

static unsigned long long rng_state = 0x9E3779B97F4A7C15ull;
static inline unsigned long long rnd() {   // xorshift64*
    rng_state ^= rng_state >> 12;
    rng_state ^= rng_state << 25;
    rng_state ^= rng_state >> 27;
    return rng_state * 0x2545F4914F6CDD1Dull;
}

static unsigned long long bits_of(double d) {
    unsigned long long u;
    memcpy(&u, &d, 8);
    return u;
}

int main(int argc, char **argv) {
    const long n_a = argc > 1 ? atol(argv[1]) : 400000, n_b = argc > 2 ? atol(argv[2]) : 400000;
    std::vector<double> tab;
    char err[256];
    if (build_rs1_tables(tab, err, sizeof err)) {
        printf("FAIL build_rs1_tables: %s\n", err);
        return 1;
    }
    long bad = 0;
    // ---------------------------------------------------------------- A
    for (long i = 0; i < n_a && bad < 10; ++i) {
        const unsigned long long r = rnd(), q = rnd();
        uint32_t s0 = (uint32_t)r & 0x3FFFFFFFu, s1 = (uint32_t)(r >> 32) & 0x3FFFFFFFu;
        uint32_t valid = 0x3FFFFFFFu;
        if (i % 3 == 0) valid &= (uint32_t)q | (uint32_t)(q >> 32) | (uint32_t)rnd();   // a few bases that do not score
        if (i % 7 == 0) s0 &= (uint32_t)rnd(), s1 |= (uint32_t)rnd() & 0x3FFFFFFFu;    // GC-rich rows
        // bases 2 and 3 of every scanned 30-mer are the PAM's upper-case C (code 2), and they score
        s0 &= ~0xCu, s1 |= 0xCu, valid |= 0xCu;
        s0 &= valid, s1 &= valid;                                                        // planes of bases that do not score are zero
        const double a = rs1_canonical(tab.data(), s0, s1, valid);
        const double b = rs1_dense(s0, s1, valid, CRP_CLASS_CANONICAL, CRP_CLASS_CANONICAL);
        if (bits_of(a) != bits_of(b)) {
            printf("FAIL A: s0=%08x s1=%08x valid=%08x table %.17g dense %.17g\n", s0, s1, valid, a, b);
            ++bad;
        }
    }
    // ---------------------------------------------------------------- B
    static const char alphabet[] = "AAAACCCCGGGGTTTTacgtNnUZ*-RYx";
    std::vector<uint4> rec(kRecWords);
    long done = 0, n_trunc = 0, n_irr = 0, n_uns = 0;
    while (done < n_b && bad < 10) {
        // a fresh record: word 0 descriptor (unused here), words 1 .. 514 planes of random bytes
        const int flavour = (int)(rnd() % 4);       // 0: everything, 1: upper-case only, 2: soft-masked mix, 3: mostly odd bytes
        for (int w = 1; w < kRecWords; ++w) {
            uint32_t o0 = 0, o1 = 0, ol = 0, oo = 0;
            for (int b = 0; b < 32; ++b) {
                char c;
                const unsigned long long r = rnd();
                if (flavour == 1) c = "ACGT"[r & 3];
                else if (flavour == 2) c = "ACGTacgt"[r & 7];
                else if (flavour == 3) c = alphabet[16 + r % (sizeof alphabet - 17)];
                else c = alphabet[r % (sizeof alphabet - 1)];
                const uint32_t nib = classify((uint32_t)(unsigned char)c);
                o0 |= (nib & 1u) << b, o1 |= ((nib >> 1) & 1u) << b, ol |= ((nib >> 2) & 1u) << b, oo |= ((nib >> 3) & 1u) << b;
            }
            rec[w] = make_uint4(o0, o1, ol, oo);
        }
        for (int k = 0; k < 4000 && done < n_b; ++k, ++done) {
            const uint32_t p = (uint32_t)(rnd() % kTile);                  // position of the hit inside the tile
            const uint32_t t_start = (uint32_t)(rnd() % 3 == 0 ? rnd() % 100000u : rnd() % 0x7FFF0000u) / kTile * kTile;
            const uint32_t t = t_start + p;
            const bool minus = rnd() & 1;
            const uint32_t need = minus ? 28u : 5u;
            // token length: mostly far away, sometimes so close that the window is cut short
            uint32_t L = (rnd() % 3 == 0) ? t + 1u + (uint32_t)(rnd() % 40u) : t + 1000u + (uint32_t)(rnd() % 100000u);
            if (L >= 0x7FFF8000u) L = 0x7FFF7FFFu;                          // crp_genome_add_segment: L < 2^31 - 2^15
            if (t >= L) continue;
            unsigned long long word = 0;
            double x, xr;
            Window w;
            if (minus) {
                const uint32_t ws = p + kWinBiasMinus, xt = L - need - (t_start - kWinBiasMinus);
                x = score_hit<true>(tab.data(), rec.data(), ws, xt, word);
                w = extract_window<true>(rec.data(), ws, t, L);
            } else {
                const uint32_t ws = p + kWinBiasPlusHot, xt = L - need - (t_start - kWinBiasPlusHot);
                x = score_hit<false>(tab.data(), rec.data(), ws, xt, word);
                w = extract_window<false>(rec.data(), p + kWinBiasPlus, t, L);
            }
            xr = rs1_canonical(tab.data(), w.s0, w.s1, w.valid);
            n_trunc += (w.packed & CRP_PACKED_TRUNCATED) != 0, n_irr += (w.packed & CRP_PACKED_IRREGULAR) != 0;
            n_uns += (w.packed & CRP_PACKED_UNSCORED) != 0;
            if (word != w.packed || bits_of(x) != bits_of(xr)) {
                printf("FAIL B: %c p=%u t=%u L=%u flavour %d: packed %016llx / %016llx, x %.17g / %.17g\n", minus ? '-' : '+', p, t, L,
                       flavour, word, w.packed, x, xr);
                ++bad;
            }
        }
    }
    // ---------------------------------------------------------------- C
    long n_tiles_c = 0, n_hdr_tiles = 0, n_hits_c = 0;
    {
        static const char gc_alphabet[] = "GGGGCCCCAATTgcatN";
        static const int guides[] = {1, 18, 20, 23, 100, 5000, 20000, 100000};
        std::vector<uint4> rec(kRecWords);
        for (int round = 0; round < 60 && bad < 10; ++round) {
            const uint32_t L = 1u + (uint32_t)(rnd() % (round % 4 == 0 ? 300u : 90000u));
            std::vector<unsigned char> tok(L);
            for (auto &ch : tok) ch = (unsigned char)gc_alphabet[rnd() % (sizeof gc_alphabet - 1)];
            if (round % 5 == 0) for (auto &ch : tok) ch = (rnd() & 1) ? 'G' : 'C';            // dense
            // the segment: the whole token, or a 128-aligned part of it (a shard boundary inside the token)
            uint32_t seg_b = 0, seg_e = L;
            if (round % 3 == 1 && L > 512) {
                seg_b = (uint32_t)(rnd() % (L / 2)) / kAlign * kAlign;
                seg_e = seg_b + (uint32_t)(1 + rnd() % (L - seg_b));
                if (seg_e < L) seg_e = (seg_e + kAlign - 1) / kAlign * kAlign;
                if (seg_e > L) seg_e = L;
            }
            const int64_t st_lo = seg_b >= 32 ? seg_b - 32 : 0, st_hi = seg_e + 32 < L ? seg_e + 32 : L;
            const int l = guides[rnd() % (sizeof guides / sizeof guides[0])];
            for (uint32_t t_start = seg_b; t_start < seg_e; t_start += kTile, ++n_tiles_c) {
                const TileDesc td = {t_start, L, seg_e - t_start < (uint32_t)kTile ? seg_e - t_start : (uint32_t)kTile, 0u};
                for (int k = 1; k < kRecWords; ++k) {                  // the planes k_pack writes
                    uint32_t o0 = 0, o1 = 0, ol = 0, oo = 0;
                    for (int b = 0; b < 32; ++b) {
                        const int64_t q = (int64_t)t_start + ((int64_t)k - 2) * 32 + b;
                        const uint32_t nib = q >= st_lo && q < st_hi ? classify(tok[(size_t)q]) : 8u;
                        o0 |= (nib & 1u) << b, o1 |= ((nib >> 1) & 1u) << b, ol |= ((nib >> 2) & 1u) << b, oo |= ((nib >> 3) & 1u) << b;
                    }
                    rec[k] = make_uint4(o0, o1, ol, oo);
                }
                uint32_t naive[8][2] = {}, emit[8][2] = {}, pack[8][2] = {};
                for (uint32_t t = t_start; t < t_start + td.n; ++t) {    // the reference's tests, byte by byte
                    const int c = (int)((t - t_start) / 2048u);
                    const bool third = (int64_t)t + 2 < (int64_t)L;
                    if (third && (int64_t)t >= l + 5 && tok[t + 1] == 'G' && tok[t + 2] == 'G') ++naive[c][0];
                    if (third && t >= 2 && (int64_t)t <= (int64_t)L - l + 7 && tok[t] == 'C' && tok[t + 1] == 'C') ++naive[c][1];
                }
                for (int w = 0; w < kWarps; ++w)
                    for (int lane = 0; lane < 32; ++lane) {
                        const Hits h = tile_hits(rec.data(), td, l, 64 * w + lane);
                        emit[w][0] += crp_popc(h.pA) + crp_popc(h.pB);
                        emit[w][1] += crp_popc(h.mA) + crp_popc(h.mB);
                    }
                for (int i = 0; i < kTileWords; ++i) {
                    const uint4 a = rec[2 + i], an = rec[3 + i];
                    const uint32_t up = ~(a.z | a.w), upn = ~(an.z | an.w);
                    const uint32_t v = pack_word_hits(a.x & a.y & up, ~a.x & a.y & up, an.x & an.y & upn & 3u, ~an.x & an.y & upn & 3u,
                                                      (int32_t)(t_start + 32u * (uint32_t)i), td);
                    pack[i / 64][0] += v & 0xFFFFu, pack[i / 64][1] += v >> 16;
                }
                const bool l_edge = (int32_t)t_start < l + 5 || (int32_t)t_start + kTile - 1 > (int32_t)L - l + 7;   // the count phase's test
                n_hdr_tiles += !l_edge;
                for (int c = 0; c < 8; ++c)
                    for (int sd = 0; sd < 2; ++sd) {
                        n_hits_c += naive[c][sd];
                        if (emit[c][sd] != naive[c][sd] || (!l_edge && pack[c][sd] != naive[c][sd])) {
                            printf("FAIL C: L=%u segment [%u, %u) tile at %u l=%d chunk %d strand %c: bytes %u, tile_hits %u, header %u%s\n", L, seg_b,
                                   seg_e, t_start, l, c, sd ? '-' : '+', naive[c][sd], emit[c][sd], pack[c][sd], l_edge ? " (counted at scan time)" : "");
                            ++bad;
                        }
                    }
            }
        }
        if (!bad && (n_hdr_tiles == 0 || n_hdr_tiles == n_tiles_c || n_hits_c == 0)) {
            printf("FAIL C: the sample does not hold both kinds of tile\n");
            return 1;
        }
    }
    if (bad) return 1;
    if (!n_trunc || !n_irr || !n_uns || n_trunc == done || n_irr == done || n_uns == done) {
        printf("FAIL B: the sample does not exercise every flag both ways\n");
        return 1;
    }
    printf("OK %ld dense rows, %ld windows (%ld truncated, %ld irregular, %ld unscored), %d table doubles, %ld tiles (%ld on their headers, %ld hits)\n",
           n_a, done, n_trunc, n_irr, n_uns, (int)RS1_TABLE_DOUBLES, n_tiles_c, n_hdr_tiles, n_hits_c);
    return 0;
}
