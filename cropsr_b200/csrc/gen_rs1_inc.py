#!/usr/bin/env python3
"""Generate rs1_weights.inc from cropsr_b200/rs1.py.

The reference sums  matrix @ weights  with OpenBLAS, whose 4-row dgemv_t kernel
keeps one sequential accumulator per column-mod-4 lane.  Lane =
base code of a first-order term / code of the SECOND base of a dinucleotide
term; inside a lane the non-zero weights are added in ascending column order.
So a lane's value depends only on WHICH of its <= 14 entries match, and the
sequential fp64 sum of every possible subset can be tabulated once, exactly.

This script emits
  * the dense weight vectors (for the dense re-scoring kernel),
  * per lane: the entries, a perfect hash of the match bits (found here by
    random search and re-verified on the host at crp_init), the number of
    leading entries covered by the table and the tail entries that are still
    added one by one,
  * RS1_LANE_SUMS -- the device code that evaluates the 8 lanes.

Hash of a lane (chosen for the instruction mix of the scan kernel, whose
bottleneck is the integer ALU pipe): one IMAD.HI per group of match bits on the
FMA pipe, one AND,
    byte offset = (sum_g umulhi(bits_g, magic_g)) & (((1 << bits) - 1) << 3)
i.e. bits [35, 35 + bits) of the 64-bit products, already scaled to doubles.

Tail entries cost one AND and one DFMA: the match bit p (>= 20) of the class
mask, used as the HIGH word of a double, is the power of two 2^(2^(p-20) - 1023);
multiplied by the weight pre-scaled with the inverse power the product is the
weight exactly (or +-0), so fma(bit, w', acc) rounds exactly like acc + w.

usage: python gen_rs1_inc.py > rs1_weights.inc
"""
import importlib.util
import itertools
import os

import numpy as np

here = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("rs1", os.path.join(here, "..", "rs1.py"))
rs1 = importlib.util.module_from_spec(spec)
spec.loader.exec_module(rs1)

BASES = rs1.BASES
MAX_TABLE_ENTRIES = 10          # leading entries of a lane covered by its table
TABLE_ENTRIES = {"fG": 9}       # per-lane override (keeps every table at <= 1024 doubles)
FIELD_SHIFT = 3                 # index field sits at bits [3, 3 + bits) of the summed high words


def lanes():
    out = []
    for c, name in enumerate(BASES):
        out.append(("f" + name, c, [(p, None, v) for (p, cc), v in sorted(rs1.FIRST.items()) if cc == c]))
    for c, name in enumerate(BASES):
        out.append(("d" + name, c, [(p, c1, v) for (p, c1, c2), v in sorted(rs1.SECOND.items()) if c2 == c]))
    return out


def valid_states(entries):
    by_pos = {}
    for i, (p, c1, _) in enumerate(entries):
        by_pos.setdefault(p, []).append(i)
    opts = [[()] + [(i,) for i in idxs] for idxs in by_pos.values()]
    for combo in itertools.product(*opts):
        yield tuple(sorted(i for t in combo for i in t))


def find_hash(entries, seed):
    """Smallest index width with an injective hash
        h = sum over first-base groups of (umulhi(x_g, hi_g) + x_g * lo_g)   (mod 2^32)
    whose bits [3, 3 + bits) are the table slot.  lo_g is tried as 0 first (one multiply
    per group); match bits below position 4 cannot reach the field through the high word,
    so lanes that have them need the low product as well."""
    groups = {}
    for i, (p, c1, _) in enumerate(entries):
        groups.setdefault(c1, []).append(i)
    glist = list(groups.items())
    states = list(valid_states(entries))
    pat = np.zeros((len(glist), len(states)), dtype=np.uint64)
    for si, s in enumerate(states):
        for gi, (_, idxs) in enumerate(glist):
            pat[gi, si] = sum(1 << entries[i][0] for i in s if i in idxs)
    need = int(np.ceil(np.log2(len(states))))
    rng = np.random.default_rng(seed)
    batch = 1024
    m32 = np.uint64(0xFFFFFFFF)

    def sparse():
        m = rng.integers(0, 2 ** 32, size=batch, dtype=np.uint64)
        for _ in range(int(rng.integers(0, 4))):
            m &= rng.integers(0, 2 ** 32, size=batch, dtype=np.uint64)
        return m

    for bits in (need, need + 1, need + 2):
        field = np.uint64((1 << bits) - 1)
        for with_lo in (False, True):
            for t in range(1500 if bits == need else 400):
                acc = np.zeros((batch, len(states)), dtype=np.uint64)
                his, los = [], []
                for gi in range(len(glist)):
                    hi = sparse()
                    lo = sparse() if with_lo else np.zeros(batch, dtype=np.uint64)
                    his.append(hi)
                    los.append(lo)
                    acc += (pat[gi][None, :] * hi[:, None]) >> np.uint64(32)
                    acc += (pat[gi][None, :] * lo[:, None]) & m32
                h = (acc >> np.uint64(FIELD_SHIFT)) & field
                h.sort(axis=1)
                ok = (np.diff(h.astype(np.int64), axis=1) != 0).all(axis=1)
                if ok.any():
                    k = int(np.argmax(ok))
                    return bits, [(c1, sum(1 << entries[i][0] for i in idxs), int(his[gi][k]), int(los[gi][k]))
                                  for gi, (c1, idxs) in enumerate(glist)]
    raise SystemExit("no perfect hash found")


def hexf(v):
    return float(v).hex()


def main():
    out = []
    emit = out.append
    emit("// GENERATED by gen_rs1_inc.py from cropsr_b200/rs1.py -- do not edit.")
    emit(f"#define RS1_INTERCEPT {hexf(rs1.INTERCEPT)} /* {rs1.INTERCEPT!r} */")
    emit(f"#define RS1_LOW_GC {hexf(rs1.LOW_GC)} /* {rs1.LOW_GC!r} */")
    w1, w2 = rs1.dense_first(), rs1.dense_second()
    emit("#define RS1_DENSE_FIRST { " + ", ".join(hexf(v) if v else "0.0" for v in w1) + " }")
    emit("#define RS1_DENSE_SECOND { " + ", ".join(hexf(v) if v else "0.0" for v in w2) + " }")

    # ---- lane descriptors for the host-side table builder
    emit("struct Rs1Entry { int pos; int first_base; double weight; };     // first_base < 0: first-order term")
    emit("struct Rs1Group { int first_base; unsigned mask; unsigned magic_hi; unsigned magic_lo; };")
    emit("// table slot of a set of matching entries = ((sum over groups of umulhi(x_g, magic_hi) + x_g * magic_lo) >> 3) & (2^bits - 1)")
    emit("struct Rs1Lane { const char *name; int lane_base; int n_entries; int n_table; int bits; int offset; "
         "int n_groups; Rs1Group groups[4]; Rs1Entry entries[16]; };")
    descs, code, offset, consts = [], [], 0, []
    mask_name = {0: "mA", 1: "mT", 2: "mC", 3: "mG"}
    next_name = {0: "nA", 1: "nT", 2: "nC", 3: "nG"}
    for li, (name, base, entries) in enumerate(lanes()):
        n_table = min(len(entries), TABLE_ENTRIES.get(name, MAX_TABLE_ENTRIES))
        # never split a group of mutually exclusive entries (same position) across table / tail
        while n_table < len(entries) and n_table > 0 and entries[n_table][0] == entries[n_table - 1][0]:
            n_table -= 1
        bits, groups = find_hash(entries[:n_table], seed=100 + li)
        ents = ", ".join(f"{{{p}, {-1 if c1 is None else c1}, {hexf(v)}}}" for p, c1, v in entries)
        grps = ", ".join(f"{{{-1 if c1 is None else c1}, 0x{m:x}u, 0x{mh:x}u, 0x{ml:x}u}}" for c1, m, mh, ml in groups)
        descs.append(f'    {{"{name}", {base}, {len(entries)}, {n_table}, {bits}, {offset}, {len(groups)}, {{{grps}}}, {{{ents}}}}}')
        # ---- device code of this lane
        second = name[0] == "d"
        terms = []
        for c1, m, mh, ml in groups:
            src = f"{mask_name[c1]} & {next_name[base]} & 0x{m:x}u" if second else f"{mask_name[base]} & 0x{m:x}u"
            terms.append(f"__umulhi({src}, 0x{mh:x}u)")
            if ml:
                terms.append(f"({src}) * 0x{ml:x}u")
        code.append(f"    double {name} = RS1_LD(T, {8 * offset}u, (" + " + ".join(terms) + f") & 0x{((1 << bits) - 1) << FIELD_SHIFT:x}u);")
        for p, c1, v in entries[n_table:]:
            # entries at one position are mutually exclusive: at most one of the fma's adds a non-zero
            assert p >= 20, "tail entry below bit 20: extend the generator with a shift"
            e = 1 << (p - 20)
            cond = f"{mask_name[c1]} & {next_name[base]} & 0x{1 << p:x}u" if second else f"{mask_name[base]} & 0x{1 << p:x}u"
            code.append(f"    {name} = RS1_FMA_BIT({cond}, RS1_K({len(consts)}), {name});   // {v!r} * 2^{1023 - e}")
            consts.append(hexf(v * 2.0 ** (1023 - e)))
        offset += 1 << bits
    emit(f"#define RS1_TABLE_DOUBLES {offset}")
    emit("// pre-scaled tail weights, then intercept and low_gc: kept in a __constant__ array so that DFMA / DADD")
    emit("// read them as constant-bank operands (64-bit immediates would cost two UMOV each)")
    emit(f"#define RS1_K_INTERCEPT {len(consts)}")
    emit(f"#define RS1_K_LOW_GC {len(consts) + 1}")
    emit("#define RS1_K_VALUES { " + ", ".join(consts + ["RS1_INTERCEPT", "RS1_LOW_GC"]) + " }")
    emit("static const Rs1Lane kRs1Lanes[8] = {")
    emit(",\n".join(descs))
    emit("};")
    emit("// The 8 lane sums of one scored 30-mer: class masks m* (bit q = base q has that code")
    emit("// and scores), T = the lane tables in shared memory.  Lane tables cover the leading")
    emit("// entries of each lane; the few remaining ones are added in order.")
    emit("#define RS1_LANE_SUMS(T, mA, mT, mC, mG) \\")
    body = ["    const uint32_t nA = (mA) >> 1, nT = (mT) >> 1, nC = (mC) >> 1, nG = (mG) >> 1;"] + code
    emit(" \\\n".join(line.split("   // ")[0] for line in body))
    print("\n".join(out))


if __name__ == "__main__":
    main()
