"""Contig-chunk sharding of a genome across the GPUs of one box.

Every PAM position is independent given 25 bases of left and 27 of right
context, so tokens are cut into segments at multiples of the scan tile and
dealt to ranks as ONE contiguous run of positions each (balanced by length);
the library stages the halo around every segment itself.  The only exchange is
an all-gather of per-segment, per-strand candidate counts, from which every
rank derives the global position of its candidates in reference order
(token order; inside a token all '+' by ascending t, then all '-').
Pure index arithmetic -- unit-tested on the CPU (tests/test_host_logic.py) and
under gloo with world_size 2 and 3 (tests/test_multi_rank_gloo.py).
"""
from .engine import TILE


def plan(token_lengths, world_size, granule=TILE):
    """-> per rank: list of (token_id, begin, end) segments, in genome order.
    Every rank gets one contiguous run of the genome; the runs are balanced by TILES, not by
    positions (a token's last tile is scanned whole even when it is half empty, so 10,000 short
    scaffolds weigh more than one chromosome of the same length), and boundaries inside a token
    sit on multiples of `granule`."""
    padded = [(n + granule - 1) // granule * granule for n in token_lengths]
    total = sum(token_lengths)
    total_padded = sum(padded)
    starts, pstarts = [], []
    acc = pacc = 0
    for n, pn in zip(token_lengths, padded):
        starts.append(acc)
        pstarts.append(pacc)
        acc += n
        pacc += pn
    # cut points in padded coordinates, snapped down to a granule, mapped back to real positions
    cuts = [0]
    k = 0
    for r in range(1, world_size):
        target = total_padded * r // world_size
        while k + 1 < len(token_lengths) and pstarts[k + 1] <= target:
            k += 1
        cut = 0
        if token_lengths:
            inside = target - pstarts[k]
            inside -= inside % granule
            cut = starts[k] + min(inside, token_lengths[k])
        cuts.append(max(cut, cuts[-1]))
    cuts.append(total)
    out = [[] for _ in range(world_size)]
    r = 0
    for k, n in enumerate(token_lengths):
        s, e = starts[k], starts[k] + n
        if n == 0:                                   # an empty token goes to the rank whose run holds its position
            while r + 1 < world_size and cuts[r + 1] <= s and cuts[r + 1] < total:
                r += 1
            out[r].append((k, 0, 0))
            continue
        while r + 1 < world_size and cuts[r + 1] <= s:
            r += 1
        q = r
        while True:                                  # the token may span several ranks
            a, b = max(cuts[q], s), min(cuts[q + 1], e)
            if a < b:
                out[q].append((k, a - s, b - s))
            if cuts[q + 1] >= e or q + 1 >= world_size:
                break
            q += 1
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
