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
bottleneck is the integer ALU pipe): one AND per group of match bits, then only
multiplies, which run on the FMA pipe,
    slot = (sum_g bits_g * magic_g  mod 2^32) >> (32 - bits)
the shift written as a multiply-high by 2^bits.  Dinucleotide match bits are
(first-base mask << 1) & second-base mask, so that the shift is a multiply too.
Entries that can never match a scanned 30-mer (its bases 2 and 3 are the PAM's
upper-case CC) are dropped; the two that always match are folded into the table.

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
A, T, C, G = 0, 1, 2, 3
MAX_TABLE_ENTRIES = 10          # leading free entries of a lane covered by its table
TABLE_ENTRIES = {"fG": 9}       # per-lane override (keeps the tables of one CTA at ~25 KB)
for _kv in filter(None, os.environ.get("RS1_TABLE_ENTRIES", "").split(",")):      # e.g. RS1_TABLE_ENTRIES=fC=9
    TABLE_ENTRIES[_kv.split("=")[0]] = int(_kv.split("=")[1])
# Every 30-mer the scan scores carries the PAM the way the reference orients it: '+' windows
# are reverse-complemented (.GG -> CC.), '-' windows are read forwards (CC.), so bases 2 and 3
# are upper-case C on both strands (CROPSR.py:415-433; tests/test_oracle_golden.py pins it).
KNOWN = {2: C, 3: C}


def lanes():
    """-> [(name, lane base, entries)], entries = (pos, first base or None, weight, forced) in
    ascending column order; entries that cannot match a scanned 30-mer are dropped, entries
    that always match are `forced` (always added, no hash bit)."""
    out = []
    for c, name in enumerate(BASES):
        ents = []
        for (p, cc), v in sorted(rs1.FIRST.items()):
            if cc != c:
                continue
            if p in KNOWN and KNOWN[p] != c:
                continue
            ents.append((p, None, v, p in KNOWN, c))
        out.append(("f" + name, c, ents))
    for c, name in enumerate(BASES):
        ents = []
        for (p, c1, c2), v in sorted(rs1.SECOND.items()):
            if c2 != c:
                continue
            if (p in KNOWN and KNOWN[p] != c1) or (p + 1 in KNOWN and KNOWN[p + 1] != c2):
                continue
            ents.append((p, c1, v, p in KNOWN and p + 1 in KNOWN, c))
        out.append(("d" + name, c, ents))
    return out


def bit_of(entry):
    p, c1 = entry[0], entry[1]
    return p if c1 is None else p + 1       # pair bits: first-base mask shifted left by one


def valid_states(entries):
    by_pos = {}
    for i, e in enumerate(entries):
        by_pos.setdefault(e[0], []).append(i)
    opts = [[()] + [(i,) for i in idxs] for idxs in by_pos.values()]
    for combo in itertools.product(*opts):
        yield tuple(sorted(i for t in combo for i in t))


GC_MODEL = 0.42                 # base composition of the Monte-Carlo that ranks hash candidates
SEQUENTIAL_MAX = 3               # lanes with at most this many entries are summed without a table
SWIZZLE_MIN_BITS = 6            # lanes with at least this many index bits add their top 4 hash bits to the slot              # bit-flip hill climbing on the best candidate
N_CANDIDATES = int(os.environ.get("RS1_NCAND", "80"))               # injective hashes collected per lane before the best one is kept
EFFORT = int(os.environ.get("RS1_EFFORT", "2"))   # multiplies the length of the random search
MERGE_GROUPS = True             # dinucleotide groups at disjoint positions: one bit select + one multiply for the set


def sample_states(entries, rng, n):
    """n random states of the free table entries under i.i.d. bases: -> list of index tuples"""
    f = {A: (1 - GC_MODEL) / 2, T: (1 - GC_MODEL) / 2, C: GC_MODEL / 2, G: GC_MODEL / 2}
    by_pos = {}
    for i, (p, c1, _, _, _) in enumerate(entries):
        by_pos.setdefault(p, []).append(i)
    out = [[] for _ in range(n)]
    for p, idxs in by_pos.items():
        # entries at one position are mutually exclusive: pick at most one
        probs = []
        for i in idxs:
            c1 = entries[i][1]
            lane_c = entries[i][4]
            probs.append(f[lane_c] if c1 is None else f[c1] * f[lane_c])
        u = rng.random(n)
        edges = np.cumsum(probs)
        pick = np.searchsorted(edges, u, side="right")
        for k in np.nonzero(pick < len(idxs))[0]:
            out[k].append(idxs[pick[k]])
    return out


def expected_wavefronts(slots, offset):
    """LDS.64 of a warp = two half-warps of 16 lanes; a half-warp costs as many wavefronts as the
    largest number of DISTINCT slots that fall on one of the 16 bank pairs."""
    n = len(slots) // 16 * 16
    s = (slots[:n] + offset).reshape(-1, 16)
    total = 0
    for row in s:
        u = np.unique(row)
        total += np.bincount(u & 15, minlength=16).max()
    return 2.0 * total / len(s)


SWIZZLE_SPARE = (8, 16, 24, 32, 48, 64)    # spare slots tried for a swizzled table (its size is 2^bits + spare)


def slot_of(h, bits, mul):
    """table slot of the 32-bit hash h (uint64 array holding values < 2^32).
    mul == 0 (narrow tables): the top `bits` bits.  Otherwise the high word of h * mul, mul = 2^bits + spare
    -- ONE multiply-high on the device: the top bits, plus the top few bits again (which spreads the probable
    states over the shared-memory banks), plus a carry out of the low part; find_hash demands that the
    result is injective over the lane's states, and the table has `mul` slots."""
    if mul:
        return (h * np.uint64(mul)) >> np.uint64(32)
    return h >> np.uint64(32 - bits)


def find_hash(entries, seed, offset):
    """Smallest index width with an injective multiplicative hash
        slot = (sum over first-base groups of x_g * magic_g  mod 2^32) >> (32 - bits)
    over every state the free table entries can take; among the injective candidates the one
    with the fewest expected shared-memory bank conflicts (Monte-Carlo over random 30-mers)."""
    groups = {}
    for i, e in enumerate(entries):
        groups.setdefault(e[1], []).append(i)
    glist = list(groups.items())
    # first-base groups whose match bits sit at disjoint positions share ONE magic: the device then
    # merges them with bit selects and multiplies once (MERGE_GROUPS; the host table builder does not care)
    gmask = [sum(1 << bit_of(entries[i]) for i in idxs) for _, idxs in glist]
    sets = []
    for gi in sorted(range(len(glist)), key=lambda g: -bin(gmask[g]).count("1")):
        for st_ in sets:
            if MERGE_GROUPS and glist[gi][0] is not None and all(gmask[gi] & gmask[o] == 0 for o in st_):
                st_.append(gi)
                break
        else:
            sets.append([gi])
    states = list(valid_states(entries))
    pat = np.zeros((len(glist), len(states)), dtype=np.uint64)
    for si, st in enumerate(states):
        for gi, (_, idxs) in enumerate(glist):
            pat[gi, si] = sum(1 << bit_of(entries[i]) for i in st if i in idxs)
    need = max(int(np.ceil(np.log2(len(states)))), 1)
    rng = np.random.default_rng(seed)
    batch = 2048
    m32 = np.uint64(0xFFFFFFFF)
    mc = sample_states(entries, np.random.default_rng(seed + 7), 16 * 1500)
    mc_pat = np.zeros((len(glist), len(mc)), dtype=np.uint64)
    for k, st in enumerate(mc):
        for gi, (_, idxs) in enumerate(glist):
            mc_pat[gi, k] = sum(1 << bit_of(entries[i]) for i in st if i in idxs)

    def sparse():
        m = rng.integers(0, 2 ** 32, size=batch, dtype=np.uint64)
        for _ in range(int(rng.integers(0, 4))):
            m &= rng.integers(0, 2 ** 32, size=batch, dtype=np.uint64)
        return m

    for bits in (need, need + 1, need + 2):
        muls = [(1 << bits) + sp for sp in SWIZZLE_SPARE] if bits >= SWIZZLE_MIN_BITS else [0]
        found = []                                  # (magics, mul)
        for t in range((6000 if bits == need else 800) * EFFORT):
            acc = np.zeros((batch, len(states)), dtype=np.uint64)
            mags = [None] * len(glist)
            for st_ in sets:
                mg = sparse()
                for gi in st_:
                    mags[gi] = mg
                    acc += (pat[gi][None, :] * mg[:, None]) & m32
            acc &= m32
            for mul in muls:
                h = slot_of(acc, bits, mul)
                h.sort(axis=1)
                ok = (np.diff(h.astype(np.int64), axis=1) != 0).all(axis=1) if len(states) > 1 else np.ones(batch, bool)
                for k in np.nonzero(ok)[0]:
                    found.append(([int(mags[gi][k]) for gi in range(len(glist))], mul))
            if len(found) >= N_CANDIDATES or (found and t > 1500 * EFFORT):
                break
        if found:
            def cost_of(mg, mul):
                hh = np.zeros(len(mc), dtype=np.uint64)
                for gi in range(len(glist)):
                    hh += (mc_pat[gi] * np.uint64(mg[gi])) & m32
                return expected_wavefronts(slot_of(hh & m32, bits, mul).astype(np.int64), offset)

            cost, (mg, mul) = min(((cost_of(*c), c) for c in found[:N_CANDIDATES]), key=lambda t: t[0])
            import sys
            print(f"lane hash: {len(states)} states, {bits} bits, {len(found)} candidates, expected wavefronts "
                  f"{cost:.2f}" + (f" (table of {mul} slots)" if mul else ""), file=sys.stderr)
            return bits, mul, [(c1, sum(1 << bit_of(entries[i]) for i in idxs), mg[gi])
                               for gi, (c1, idxs) in enumerate(glist)], sets
    raise SystemExit("no perfect hash found")


def hexf(v):
    return float(v).hex()


def main():
    out = []
    emit = out.append
    emit("// GENERATED by gen_rs1_inc.py from cropsr_b200/rs1.py -- do not edit.  Regenerate with `make regen` (minutes: the")
    emit("// hash search is seeded, so the same rs1.py gives this file bit for bit); tests/test_host_logic.py compares the")
    emit("// digest below with the rs1.py in the tree.")
    import hashlib
    with open(os.path.join(here, "..", "rs1.py"), "rb") as f:
        emit(f"// rs1.py sha256 {hashlib.sha256(f.read()).hexdigest()}")
    emit(f"#define RS1_INTERCEPT {hexf(rs1.INTERCEPT)} /* {rs1.INTERCEPT!r} */")
    emit(f"#define RS1_LOW_GC {hexf(rs1.LOW_GC)} /* {rs1.LOW_GC!r} */")
    w1, w2 = rs1.dense_first(), rs1.dense_second()
    emit("#define RS1_DENSE_FIRST { " + ", ".join(hexf(v) if v else "0.0" for v in w1) + " }")
    emit("#define RS1_DENSE_SECOND { " + ", ".join(hexf(v) if v else "0.0" for v in w2) + " }")

    # ---- lane descriptors for the host-side table builder
    emit("// bit: position of the entry's match bit (first order: p; dinucleotide: p + 1, the first-base mask is")
    emit("// shifted left by one); forced: the entry matches every scanned 30-mer (PAM bases) and has no bit")
    emit("struct Rs1Entry { int pos; int first_base; double weight; int bit; int forced; };     // first_base < 0: first-order term")
    emit("struct Rs1Group { int first_base; unsigned mask; unsigned magic; };")
    emit("// table slot of a set of matching entries: h = sum over groups of x_g * magic  mod 2^32, slot = h >> (32 - bits),")
    emit("// or, if swizzle != 0, slot = (h * swizzle) >> 32 with swizzle = 2^bits + spare: one multiply-high, injective by")
    emit("// search, spreads the probable states over the shared-memory banks; the table then has `swizzle` slots")
    emit("struct Rs1Lane { const char *name; int lane_base; int n_entries; int n_table; int bits; int swizzle; int offset; "
         "int n_groups; Rs1Group groups[4]; Rs1Entry entries[16]; };")
    descs, code, offset, consts = [], [], 0, []
    mask_name = {0: "mA", 1: "mT", 2: "mC", 3: "mG"}
    shift_name = {0: "sA", 1: "sT", 2: "sC", 3: "sG"}
    for li, (name, base, entries) in enumerate(lanes()):
        free = [e for e in entries if not e[3]]
        n_free_table = min(len(free), TABLE_ENTRIES.get(name, MAX_TABLE_ENTRIES))
        if len(free) <= SEQUENTIAL_MAX and not any(e[3] for e in entries):
            n_free_table = 0                        # a short lane is cheaper as fused multiply-adds than as a lookup
        # never split a group of mutually exclusive entries (same position) across table / tail
        while 0 < n_free_table < len(free) and free[n_free_table][0] == free[n_free_table - 1][0]:
            n_free_table -= 1
        table_free = free[:n_free_table]
        tail = free[n_free_table:]
        assert all(e[0] < (tail[0][0] if tail else 99) for e in entries if e[3]), "forced entry after a tail entry"
        second = name[0] == "d"
        if table_free or any(e[3] for e in entries):
            bits, mul, groups, sets = find_hash(table_free, seed=100 + li, offset=offset)
        else:
            bits, mul, groups, sets = -1, 0, [], []
        n_table = len(entries) - len(tail)          # forced + tabulated entries come first
        ents = ", ".join(f"{{{p}, {-1 if c1 is None else c1}, {hexf(v)}, {bit_of((p, c1))}, {int(f)}}}"
                         for p, c1, v, f, _ in entries)
        grps = ", ".join(f"{{{-1 if c1 is None else c1}, 0x{m:x}u, 0x{mg:x}u}}" for c1, m, mg in groups)
        descs.append(f'    {{"{name}", {base}, {len(entries)}, {n_table}, {bits}, {mul}, {offset}, {len(groups)}, {{{grps}}}, {{{ents}}}}}')
        # ---- device code of this lane
        if bits >= 0:
            terms = []
            for st_ in sets:
                c1, m, mg = groups[st_[0]]
                if not second:
                    terms.append(f"({mask_name[base]} & 0x{m:x}u) * 0x{mg:x}u")
                    continue
                sel, union = shift_name[c1], m
                for gi in st_[1:]:                      # (sel & union) | (next & ~union): positions are disjoint
                    c1n, mn, mgn = groups[gi]
                    assert mgn == mg and union & mn == 0
                    sel = f"RS1_SEL(0x{union:x}u, {sel}, {shift_name[c1n]})"
                    union |= mn
                terms.append(f"({sel} & {mask_name[base]} & 0x{union:x}u) * 0x{mg:x}u")
            top, arg = ("RS1_SWZ", f"{mul}u") if mul else ("RS1_TOP", f"{bits}")
            code.append(f"    double {name} = RS1_LD(T, {offset}u, {top}(" + " + ".join(terms) + f", {arg}));")
            offset += mul if mul else 1 << bits
        else:
            code.append(f"    double {name} = 0.0;")
        for p, c1, v, _, _ in tail:
            # entries at one position are mutually exclusive: at most one of the fma's adds a non-zero
            b = bit_of((p, c1))
            cond = f"{shift_name[c1]} & {mask_name[base]} & 0x{1 << b:x}u" if second else f"{mask_name[base]} & 0x{1 << b:x}u"
            if b < 20:                              # move the bit into the exponent field first
                cond = f"RS1_SHL({cond}, {20 - b})"
                b = 20
            e = 1 << (b - 20)
            code.append(f"    {name} = RS1_FMA_BIT({cond}, RS1_K({len(consts)}), {name});   // {v!r} * 2^{1023 - e}")
            consts.append(hexf(v * 2.0 ** (1023 - e)))
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
    body = ["    const uint32_t sA = RS1_SHL1(mA), sT = RS1_SHL1(mT), sC = RS1_SHL1(mC), sG = RS1_SHL1(mG);"] + code
    emit(" \\\n".join(line.split("   // ")[0] for line in body))
    print("\n".join(out))


if __name__ == "__main__":
    main()
