// Rule-Set-1 scoring on the device (reference: /root/reference/CROPSR.py:285-313).
//
// The reference evaluates  matrix @ weights  with OpenBLAS; which fp64 additions happen in
// which order depends on the row's position inside the np.matmul call:
//   canonical  4 sequential accumulators by column mod 4, combined (p0+p2)+(p1+p3)
//   pair       2 accumulators by column mod 2, q0+q1
//   single     ddot (AVX-512) lane order
// rs1_canonical() is the fused-scan fast path (table driven, exact); rs1_dense() replays any
// of the three classes column by column for the few rows that need it (k_rescore).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

#include "../../include/cropsr_b200.h"
#include "rs1_weights.inc"

static constexpr size_t kRs1TableBytes = (size_t)RS1_TABLE_DOUBLES * sizeof(double);

// Canonical lane order: one sequential accumulator per column-mod-4 lane, i.e. per base
// class for the first-order term and per SECOND base for the dinucleotide term, columns in
// ascending order; lanes combined (p0+p2)+(p1+p3) = (A+C)+(T+G).  A lane's value is a
// function of which of its entries match, so the leading entries of every lane come from a
// table of exact sequential fp64 sums (built on the host at crp_init, staged in shared
// memory; rs1_weights.inc).  s0/s1: planar code bits of the scored 30-mer (bit q = base q),
// valid: bases that score.
__constant__ double c_rs1_k[] = RS1_K_VALUES;
#define RS1_K(i) c_rs1_k[i]

// lane table entry: `off` doubles into the tables, byte index idx8
#define RS1_LD(T, off, idx8) (*reinterpret_cast<const double *>(reinterpret_cast<const char *>(T) + 8u * (off) + (idx8)))
// top `bits` bits of the hash, scaled to bytes, as multiplies (FMA pipe; shifts would go to the ALU pipe)
#define RS1_TOP(h, bits) (__umulhi((h), 1u << (bits)) * 8u)
// the same with the top 4 hash bits added to the slot (spreads the probable states over the banks;
// the add is the accumulate operand of the multiply-high)
#define RS1_SWZ(h, bits) ((__umulhi((h), 1u << (bits)) + __umulhi((h), 16u)) * 8u)
// bit select: a where the mask is set, b elsewhere (one LOP3)
#define RS1_SEL(mask, a, b) (((a) & (mask)) | ((b) & ~(mask)))
#define RS1_SHL1(m) ((m) << 1)
#define RS1_SHL(m, k) ((m) << (k))
// acc + w iff the match bit is set, as ONE DFMA: bit (a single bit p >= 20 of a class mask)
// read as the high word of a double is a power of two, w_scaled = w / that power.
#define RS1_FMA_BIT(bit, w_scaled, acc) __fma_rn(__hiloint2double((int)(bit), 0), (w_scaled), (acc))

__device__ __forceinline__ double rs1_canonical(const double *__restrict__ T, uint32_t s0, uint32_t s1,
                                                uint32_t valid) {
    const uint32_t mA = ~s1 & ~s0 & valid, mT = ~s1 & s0 & valid, mC = s1 & ~s0 & valid, mG = s1 & s0 & valid;
    RS1_LANE_SUMS(T, mA, mT, mC, mG)
    const double first = __dadd_rn(__dadd_rn(fA, fC), __dadd_rn(fT, fG));
    const double second = __dadd_rn(__dadd_rn(dA, dC), __dadd_rn(dT, dG));
    // (score_first + score_second + intersect + low_gc) * -1, CROPSR.py:312
    // -(t + k) == (-t) + (-k) in IEEE arithmetic: the final negation rides on the operands
    return __dadd_rn(-__dadd_rn(__dadd_rn(first, second), RS1_K(RS1_K_INTERCEPT)), -RS1_K(RS1_K_LOW_GC));
}

// Host side: exact sequential sums of every valid subset of each lane's table entries
// (forced entries -- the PAM bases -- are part of every sum).
static int build_rs1_tables(std::vector<double> &tab, char *err, size_t errlen) {
    tab.assign(RS1_TABLE_DOUBLES, 0.0);
    std::vector<char> used(RS1_TABLE_DOUBLES, 0);
    for (const Rs1Lane &ln : kRs1Lanes) {
        if (ln.bits < 0) continue;                // no table: the lane is summed entry by entry
        int free_idx[16], n_free = 0;
        for (int i = 0; i < ln.n_table; ++i)
            if (!ln.entries[i].forced) free_idx[n_free++] = i;
        for (uint32_t sub = 0; sub < (1u << n_free); ++sub) {
            bool ok = true;                       // two entries at one position are mutually exclusive
            for (int i = 0; i < n_free && ok; ++i)
                for (int j = i + 1; j < n_free; ++j)
                    if ((sub >> i & 1) && (sub >> j & 1) && ln.entries[free_idx[i]].pos == ln.entries[free_idx[j]].pos)
                        ok = false;
            if (!ok) continue;
            uint32_t h = 0;
            for (int g = 0; g < ln.n_groups; ++g) {
                uint32_t x = 0;                   // match bits of this first-base group
                for (int i = 0; i < n_free; ++i) {
                    const Rs1Entry &e = ln.entries[free_idx[i]];
                    if ((sub >> i & 1) && ln.groups[g].first_base == e.first_base) x |= 1u << e.bit;
                }
                if (x & ~ln.groups[g].mask) {
                    snprintf(err, errlen, "rs1 lane %s: entry outside its group mask", ln.name);
                    return -1;
                }
                h += x * ln.groups[g].magic;
            }
            volatile double sum = 0.0;            // one IEEE add per entry, in ascending column order
            for (int i = 0, f = 0; i < ln.n_table; ++i) {
                const bool on = ln.entries[i].forced ? true : ((sub >> f++) & 1);
                if (on) sum = sum + ln.entries[i].weight;
            }
            uint32_t slot = ln.bits ? h >> (32 - ln.bits) : 0u;
            if (ln.swizzle) slot += slot >> (ln.bits - 4);
            const uint32_t idx = ln.offset + slot;
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
