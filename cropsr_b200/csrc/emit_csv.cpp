// Host-side CSV row formatter: byte-identical to what the reference writes with
// csv.writer(...).writerows(rows) in /root/reference/CROPSR.py:463-474 -- excel dialect
// (',' delimiter, QUOTE_MINIMAL with '"' doubled, "\r\n"), ints through str(), the score through
// str(numpy.float64) == repr(float) (shortest digits that round-trip, Python's switch to
// exponent notation below 1e-4 and from 1e16) -- but multi-threaded C++ instead of one Python
// tuple + csv call per row.  Pure host code: no CUDA in this file.
//
// Row (CROPSR.py:463-469), len(long_sequence) == guide_len + 10:
//     id,cas9,short,long,chrom,start,end,end-3,strand,score,,completed
// otherwise the 11-field "error row":
//     id,cas9,short,long,chrom,start,end,strand,-1,,completed
// short / long are Python slices of the token (silently truncated at its end) pushed through the
// reference's replace chains (CROPSR.py:116-129).
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <charconv>
#include <string>
#include <thread>
#include <mutex>
#include <vector>

#include "../../include/cropsr_b200.h"

namespace {

// '+' strand: gRNA(x) = A->U C->G G->C T->A (and Z->G through the chain), then reversed
// '-' strand: gRNA(revcomp(x)) = forward order, T->U, U->A, Z->C
struct Tables {
    unsigned char plus[256], minus[256];
    Tables() {
        for (int i = 0; i < 256; ++i) plus[i] = minus[i] = (unsigned char)i;
        plus['A'] = 'U'; plus['C'] = 'G'; plus['G'] = 'C'; plus['T'] = 'A'; plus['Z'] = 'G';
        minus['T'] = 'U'; minus['U'] = 'A'; minus['Z'] = 'C';
    }
};
const Tables kTab;

// Rows are written straight into a region of a scratch buffer that is big enough by construction
// (max_row_bytes below); the few std::string methods the formatter used keep their names.
struct Buf {
    char *p;
    void push_back(char c) { *p++ = c; }
    void append(const char *s, size_t n) { memcpy(p, s, n); p += n; }
    void append(size_t n, char c) { memset(p, c, n); p += n; }
    Buf &operator+=(const char *s) { const size_t n = strlen(s); memcpy(p, s, n); p += n; return *this; }
};

inline void put_field(Buf &out, const char *s, size_t n) {   // QUOTE_MINIMAL
    bool quote = false;
    for (size_t i = 0; i < n; ++i) {
        const char c = s[i];
        if (c == ',' || c == '"' || c == '\r' || c == '\n') {
            quote = true;
            break;
        }
    }
    if (!quote) {
        out.append(s, n);
        return;
    }
    out.push_back('"');
    for (size_t i = 0; i < n; ++i) {
        if (s[i] == '"') out.push_back('"');
        out.push_back(s[i]);
    }
    out.push_back('"');
}

inline void put_int(Buf &out, long long v) {
    char buf[24];
    char *e = buf + sizeof buf, *p = e;
    unsigned long long u = v < 0 ? 0ull - (unsigned long long)v : (unsigned long long)v;
    do {
        *--p = (char)('0' + u % 10);
        u /= 10;
    } while (u);
    if (v < 0) *--p = '-';
    out.append(p, (size_t)(e - p));
}

// repr(float): shortest decimal string that parses back to the same double (std::to_chars
// yields exactly those digits, the closest to x among the shortest -- what CPython's
// float_repr_style 'short' prints), laid out the way CPython does: fixed notation for
// 1e-4 <= |x| < 1e16, exponent notation otherwise.
void put_repr(Buf &out, double x) {
    if (isnan(x)) { out += "nan"; return; }
    if (isinf(x)) { out += x < 0 ? "-inf" : "inf"; return; }
    if (x == 0.0) { out += signbit(x) ? "-0.0" : "0.0"; return; }
    char buf[48];
    const std::to_chars_result res = std::to_chars(buf, buf + sizeof buf - 1, x, std::chars_format::scientific);
    *res.ptr = 0;
    // buf = [-]d[.ddd]e[+-]XX
    const char *p = buf;
    if (*p == '-') { out.push_back('-'); ++p; }
    char digits[24];
    int nd = 0;
    for (; *p && *p != 'e'; ++p)
        if (*p != '.') digits[nd++] = *p;
    const int exp10 = atoi(p + 1);                 // value = d.ddd * 10^exp10
    while (nd > 1 && digits[nd - 1] == '0') --nd;  // %.{p}e of a shorter-representable value never pads, but be safe
    const int decpt = exp10 + 1;                   // digits before the decimal point
    if (decpt > 16 || decpt < -3) {                // exponent notation
        out.push_back(digits[0]);
        if (nd > 1) {
            out.push_back('.');
            out.append(digits + 1, (size_t)(nd - 1));
        }
        char e[16];
        snprintf(e, sizeof e, "e%c%02d", exp10 < 0 ? '-' : '+', abs(exp10));
        out += e;
    } else if (decpt <= 0) {                       // 0.000ddd
        out += "0.";
        out.append((size_t)(-decpt), '0');
        out.append(digits, (size_t)nd);
    } else if (decpt >= nd) {                      // ddd000.0
        out.append(digits, (size_t)nd);
        out.append((size_t)(decpt - nd), '0');
        out += ".0";
    } else {                                       // dd.ddd
        out.append(digits, (size_t)decpt);
        out.push_back('.');
        out.append(digits + decpt, (size_t)(nd - decpt));
    }
}

struct Job {
    uint64_t n_rows;
    const char *ids;                 // n_ids x 7 bytes
    const uint64_t *id_index;        // per row
    const uint32_t *token_of, *t;
    const uint8_t *minus, *scored;
    const double *score;
    const uint8_t *const *tokens;
    const uint64_t *token_len;
    const char *const *chrom;
    const uint32_t *chrom_len;
    int guide_len;
};

// upper bound of one row: every field that can hold token or header bytes fully quoted
size_t max_row_bytes(const Job &j, uint32_t n_tokens) {
    size_t chrom = 0;
    for (uint32_t k = 0; k < n_tokens; ++k)
        if (j.chrom_len[k] > chrom) chrom = j.chrom_len[k];
    const size_t l = (size_t)j.guide_len;
    return 7 + 6 + (2 * l + 2) + 1 + (2 * (l + 10) + 2) + 1 + (2 * chrom + 2) + 1 + 3 * 21 + 2 + 40 + 16;
}

char *format_range(const Job &j, uint64_t lo, uint64_t hi, char *dst) {
    const int l = j.guide_len;
    std::string seq;
    Buf out{dst};
    for (uint64_t i = lo; i < hi; ++i) {
        const uint32_t k = j.token_of[i];
        const uint8_t *tok = j.tokens[k];
        const int64_t L = (int64_t)j.token_len[k], t = j.t[i];
        const bool minus = j.minus[i] != 0;
        out.append(j.ids + 7 * j.id_index[i], 7);
        out += ",cas9,";
        // Python slices [a, b) clipped to the token
        auto slice = [&](int64_t a, int64_t b) {
            if (a < 0) a = 0;              // cannot happen for rows the scan emits (t >= l+5 resp. t >= 2)
            if (b > L) b = L;
            seq.clear();
            if (minus)
                for (int64_t q = a; q < b; ++q) seq.push_back((char)kTab.minus[tok[q]]);
            else
                for (int64_t q = b - 1; q >= a; --q) seq.push_back((char)kTab.plus[tok[q]]);
            put_field(out, seq.data(), seq.size());
        };
        int64_t first, second;
        if (minus) {
            slice(t + 3, t + 3 + l);
            out.push_back(',');
            slice(t - 2, t + l + 8);
            first = t + 3 + l;
            second = t + 3;
        } else {
            slice(t - l, t);
            out.push_back(',');
            slice(t - l - 5, t + 5);
            first = t - l;
            second = t;
        }
        out.push_back(',');
        put_field(out, j.chrom[k], j.chrom_len[k]);
        out.push_back(',');
        put_int(out, first);
        out.push_back(',');
        put_int(out, second);
        out.push_back(',');
        if (j.scored[i]) {
            put_int(out, second - 3);                    // apply_cutsite, CROPSR.py:155-158
            out.push_back(',');
            out.push_back(minus ? '-' : '+');
            out.push_back(',');
            put_repr(out, j.score[i]);
        } else {
            out.push_back(minus ? '-' : '+');
            out += ",-1";
        }
        out += ",,completed\r\n";
    }
    return out.p;
}

}  // namespace

extern "C" int crp_format_rows(uint64_t n_rows, const char *ids, const uint64_t *id_index, const uint32_t *token_of,
                               const uint32_t *t, const uint8_t *minus, const uint8_t *scored, const double *score,
                               uint32_t n_tokens, const uint8_t *const *tokens, const uint64_t *token_len,
                               const char *const *chrom, const uint32_t *chrom_len, int guide_len, int n_threads,
                               char *out, uint64_t out_capacity, uint64_t *out_bytes) {
    if (!out_bytes) return CRP_ERR_ARG;
    *out_bytes = 0;
    if (n_rows == 0) return 0;
    if (!ids || !id_index || !token_of || !t || !minus || !scored || !score || !tokens || !token_len || !chrom ||
        !chrom_len || guide_len < 1)
        return CRP_ERR_ARG;
    for (uint64_t i = 0; i < n_rows; ++i)
        if (token_of[i] >= n_tokens) return CRP_ERR_ARG;
    Job j = {n_rows, ids, id_index, token_of, t, minus, scored, score, tokens, token_len, chrom, chrom_len, guide_len};
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 1;
    uint64_t nt = n_threads > 0 ? (uint64_t)n_threads : (hw > 16 ? 16 : hw);
    if (nt > (n_rows + 4095) / 4096) nt = (n_rows + 4095) / 4096;     // >= 4096 rows per thread
    // every thread formats its rows into its own region of a scratch buffer that lives as long as
    // the library (fresh 100+ MB buffers per call cost more in page faults than the formatting),
    // then the regions are copied, again in parallel, to their final offsets in out
    static std::mutex scratch_mu;                 // one call at a time per process: concurrent callers queue here
    std::lock_guard<std::mutex> scratch_lock(scratch_mu);
    static std::vector<char> scratch;
    const size_t row_max = max_row_bytes(j, n_tokens);
    const uint64_t per = (n_rows + nt - 1) / nt;
    const size_t stride = (size_t)per * row_max;
    if (scratch.size() < stride * nt) scratch.resize(stride * nt);
    std::vector<size_t> len(nt, 0);
    auto rows_of = [&](uint64_t w, uint64_t *lo, uint64_t *hi) {
        *lo = w * per < n_rows ? w * per : n_rows;
        *hi = (w + 1) * per < n_rows ? (w + 1) * per : n_rows;
    };
    auto run = [&](auto &&fn) {
        std::vector<std::thread> pool;
        for (uint64_t w = 1; w < nt; ++w) pool.emplace_back([&, w] { fn(w); });
        fn(0);
        for (std::thread &th : pool) th.join();
    };
    run([&](uint64_t w) {
        uint64_t lo, hi;
        rows_of(w, &lo, &hi);
        char *base = scratch.data() + w * stride;
        len[w] = (size_t)(format_range(j, lo, hi, base) - base);
    });
    uint64_t total = 0;
    std::vector<uint64_t> off(nt, 0);
    for (uint64_t w = 0; w < nt; ++w) {
        off[w] = total;
        total += len[w];
    }
    *out_bytes = total;
    if (total > out_capacity || !out) return CRP_ERR_RANGE;          // caller retries with *out_bytes
    run([&](uint64_t w) { memcpy(out + off[w], scratch.data() + w * stride, len[w]); });
    return 0;
}

// ---- crispr_id characters: get_id() of the reference (CROPSR.py:316-318) ------------------------
// np.random.choice(alphanum36, [n, 7]) on numpy's legacy global generator is, value for value,
// randint(0, 36, (n, 7)): MT19937 32-bit outputs masked with 63, those above 35 rejected
// (numpy random/src/distributions: buffered_bounded_masked_uint32).  The caller hands over the
// generator's state (np.random.get_state(): key[624], pos) and puts the advanced state back, so a
// seeded run keeps the reference's ids and whatever is drawn afterwards is unchanged too.  numpy's
// own bounded-integer path costs ~27 ns per character; this loop ~3 ns.
static inline void mt19937_refill(uint32_t *key) {
    constexpr int N = 624, M = 397;
    constexpr uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MATRIX = 0x9908b0dfu;
    int kk = 0;
    uint32_t y;
    for (; kk < N - M; ++kk) {
        y = (key[kk] & UPPER) | (key[kk + 1] & LOWER);
        key[kk] = key[kk + M] ^ (y >> 1) ^ (-(y & 1u) & MATRIX);
    }
    for (; kk < N - 1; ++kk) {
        y = (key[kk] & UPPER) | (key[kk + 1] & LOWER);
        key[kk] = key[kk + (M - N)] ^ (y >> 1) ^ (-(y & 1u) & MATRIX);
    }
    y = (key[N - 1] & UPPER) | (key[0] & LOWER);
    key[N - 1] = key[M - 1] ^ (y >> 1) ^ (-(y & 1u) & MATRIX);
}

extern "C" int crp_legacy_ids(uint32_t *mt_key, int32_t *mt_pos, uint64_t n_ids, uint8_t *out) {
    if (!mt_key || !mt_pos || (n_ids && !out)) return CRP_ERR_ARG;
    if (*mt_pos < 0 || *mt_pos > 624) return CRP_ERR_ARG;
    static const char alnum[65] = "ABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789............................";
    int pos = *mt_pos;
    const uint64_t n = n_ids * 7;
    uint64_t i = 0;
    while (i < n) {
        if (pos == 624) {
            mt19937_refill(mt_key);
            pos = 0;
        }
        // the rest of this block of state words, or as many as can still be accepted without
        // running past the end of out: branch-free (44 % of the draws are rejected)
        uint64_t take = (uint64_t)(624 - pos);
        if (take > n - i) take = n - i;
        for (uint64_t k = 0; k < take; ++k) {
            uint32_t y = mt_key[pos + (int)k];
            y ^= y >> 11;
            y ^= (y << 7) & 0x9d2c5680u;
            y ^= (y << 15) & 0xefc60000u;
            y ^= y >> 18;
            const uint32_t v = y & 63u;
            out[i] = (uint8_t)alnum[v];            // overwritten by the next accepted draw if this one is rejected
            i += v <= 35u;
        }
        pos += (int)take;
    }
    *mt_pos = pos;
    return 0;
}
