"""The N > 1 host path on CPU: two ranks over gloo share a genome by contiguous,
tile-aligned shards, exchange per-segment per-strand counts with ONE all-gather
(what bench.py does over NCCL) and derive the global row of every candidate.
The candidates themselves come from the oracle here (no GPU in this test); the
GPU test test_sharded_scan_equals_whole_scan covers the device side of the same
arithmetic."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, lengths_seed, out_dir):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    import cropsr_oracle as oracle
    from helpers import synthetic_fasta
    from cropsr_b200 import ingest, shard

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    text = synthetic_fasta(lengths_seed, [70000, 20000, 50000], gc=0.5)
    tokens = list(ingest.fasta_text_to_tokens(text).values())
    lengths = [len(t) for t in tokens]
    plans = shard.plan(lengths, world)
    mine = plans[rank]
    # this rank's candidates: the oracle's hits that fall inside its segments
    rows = []
    for k, a, b in mine:
        plus, minus = oracle.pam_hits(tokens[k], 20)
        rows.append(([t for t in plus if a <= t < b], [t for t in minus if a <= t < b]))
    n_slots = len(lengths) + 1
    buf = torch.zeros(2 * n_slots, dtype=torch.int64)
    for s, (p, m) in enumerate(rows):
        buf[s] = len(p)
        buf[n_slots + s] = len(m)
    gathered = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf)                      # the one exchange step of the path
    h = torch.stack(gathered).numpy().reshape(world, 2, n_slots)
    counts = [(h[r, 0, :len(plans[r])], h[r, 1, :len(plans[r])]) for r in range(world)]
    offsets, total = shard.global_offsets(plans, counts)
    table = {}
    for s, (k, a, b) in enumerate(mine):
        for si, hits in enumerate(rows[s]):
            for i, t in enumerate(hits):
                table[offsets[rank][s][si] + i] = (k, "+-"[si], t)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), np.array([(g, k, si == "-", t) for g, (k, si, t) in table.items()],
                                                               dtype=np.int64).reshape(-1, 4))
    if rank == 0:
        ref = []
        for k, tok in enumerate(tokens):
            plus, minus = oracle.pam_hits(tok, 20)
            ref += [(k, 0, t) for t in plus] + [(k, 1, t) for t in minus]
        np.save(os.path.join(out_dir, "ref.npy"), np.array(ref, dtype=np.int64).reshape(-1, 3))
        with open(os.path.join(out_dir, "total.txt"), "w") as f:
            f.write(str(total))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_two_ranks_place_candidates_in_reference_order(tmp_path, world):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_rank_main, args=(world, port, 5, str(tmp_path)), nprocs=world, join=True)
    ref = np.load(tmp_path / "ref.npy")
    assert int((tmp_path / "total.txt").read_text()) == len(ref)
    got = np.concatenate([np.load(tmp_path / f"rank{r}.npy") for r in range(world)])
    got = got[np.argsort(got[:, 0])]
    assert np.array_equal(got[:, 0], np.arange(len(ref)))          # every global row exactly once
    assert np.array_equal(got[:, 1:], ref)                          # and in the reference's order


def _rdv_main(rank, world, port, out_dir):
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from cropsr_b200 import launch
    rdv = launch.Rendezvous(rank, world, "127.0.0.1", port)
    got = rdv.all_gather({"rank": rank, "payload": bytes([rank]) * (1 << 18)})
    assert [g["rank"] for g in got] == list(range(world))
    assert all(g["payload"] == bytes([r]) * (1 << 18) for r, g in enumerate(got))
    uid = rdv.broadcast(b"id-from-rank-0" if rank == 0 else None)
    assert uid == b"id-from-rank-0"
    for _ in range(20):
        rdv.barrier()
    rdv.close()
    with open(os.path.join(out_dir, f"ok{rank}"), "w") as f:
        f.write("ok")


@pytest.mark.parametrize("world", [1, 2, 4])
def test_tcp_rendezvous_of_the_launcher(tmp_path, world):
    """cropsr_b200/launch.py: the channel that hands NCCL's unique id round (no PyTorch)."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    port = _free_port()
    procs = [ctx.Process(target=_rdv_main, args=(r, world, port, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(60)
    assert all(p.exitcode == 0 for p in procs)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
