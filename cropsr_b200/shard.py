"""Contig-chunk sharding of a genome across the GPUs of one box.

Every PAM position is independent given 25 bases of left and 27 of right
context, so tokens are cut into segments at multiples of the scan tile and
dealt to ranks as ONE contiguous run of positions each (balanced by length);
the library stages the halo around every segment itself.  The only exchange is
an all-gather of per-segment, per-strand candidate counts, from which every
rank derives the global position of its candidates in reference order
(token order; inside a token all '+' by ascending t, then all '-').
Pure index arithmetic -- unit-tested on the CPU (tests/test_shard.py) and
under gloo with world_size 2.
"""
from .engine import TILE


def plan(token_lengths, world_size, granule=TILE):
    """-> per rank: list of (token_id, begin, end) segments, in genome order.
    Rank r owns the positions [r*share, (r+1)*share) of the concatenated genome,
    with boundaries moved to multiples of `granule` inside a token."""
    total = sum(token_lengths)
    starts = []
    acc = 0
    for n in token_lengths:
        starts.append(acc)
        acc += n
    # cut points in concatenated coordinates, snapped down to a granule inside their token
    cuts = [0]
    for r in range(1, world_size):
        target = total * r // world_size
        k = 0
        while k + 1 < len(token_lengths) and starts[k + 1] <= target:
            k += 1
        if token_lengths:
            inside = target - starts[k]
            inside -= inside % granule
            target = starts[k] + inside
        cuts.append(max(target, cuts[-1]))
    cuts.append(total)
    out = []
    for r in range(world_size):
        lo, hi = cuts[r], cuts[r + 1]
        segs = []
        for k, n in enumerate(token_lengths):
            a, b = max(lo, starts[k]), min(hi, starts[k] + n)
            if a < b or (n == 0 and lo <= starts[k] < hi) :
                segs.append((k, a - starts[k], b - starts[k]))
        out.append(segs)
    return out


def global_offsets(plans, counts):
    """plans[r] = segments of rank r; counts[r] = (plus_counts, minus_counts)
    per segment of rank r (what the all-gather delivers).  Returns
    offsets[r][s] = (first global row of the segment's '+' hits, first global row
    of its '-' hits) in the reference's unique-candidate order, and the total."""
    n_tokens = 1 + max((seg[0] for segs in plans for seg in segs), default=-1)
    plus_tot = [0] * n_tokens
    minus_tot = [0] * n_tokens
    for r, segs in enumerate(plans):
        for s, (k, _, _) in enumerate(segs):
            plus_tot[k] += int(counts[r][0][s])
            minus_tot[k] += int(counts[r][1][s])
    token_base = []
    acc = 0
    for k in range(n_tokens):
        token_base.append(acc)
        acc += plus_tot[k] + minus_tot[k]
    run_plus = [0] * n_tokens
    run_minus = [0] * n_tokens
    offsets = []
    for r, segs in enumerate(plans):          # ranks hold consecutive pieces of a token
        row = []
        for s, (k, _, _) in enumerate(segs):
            row.append((token_base[k] + run_plus[k], token_base[k] + plus_tot[k] + run_minus[k]))
            run_plus[k] += int(counts[r][0][s])
            run_minus[k] += int(counts[r][1][s])
        offsets.append(row)
    return offsets, acc
