"""One genome on several GPUs of one box: one process per GPU, contiguous tile-aligned shards
(shard.plan), ONE exchange step -- the all-gather of the per-segment candidate counts that
crp_scan_score_sharded performs (fused into the scan kernel over peer memory, or an NCCL
all-gather behind it) -- and rows written in the reference's order.

The reference has no counterpart (CROPSR.py:409 loops over the chromosomes in one process); what
is preserved is its output order: token by token, '+' hits by ascending t, then '-' hits
(CROPSR.py:417-434).  Every rank derives the global row of each of its segments from the gathered
counts (shard.global_offsets) and copies its `pos` / `x` rows there, into shared memory; the
parent -- rank 0, which also drives the first device -- then runs the ordinary emission loop.

Processes: the caller's process is rank 0; ranks 1..N-1 are spawned (multiprocessing "spawn")
and talk to it over the TCP rendezvous of launch.py on a free loopback port.  The tokens are
handed over in one shared-memory block, so a worker copies nothing but its own shard to its GPU.
"""
import multiprocessing as mp
import socket
from multiprocessing import shared_memory

import numpy as np

from . import engine, launch, shard


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _attach(name):
    """SharedMemory that the resource tracker of this process will not unlink behind the owner's back"""
    try:
        return shared_memory.SharedMemory(name=name, track=False)       # Python >= 3.13
    except TypeError:
        shm = shared_memory.SharedMemory(name=name)
        try:
            from multiprocessing import resource_tracker
            resource_tracker.unregister(shm._name, "shared_memory")
        except Exception:
            pass
        return shm


def _rank_body(rank, world, port, device, tok_shm_name, tok_off, tok_len, guide_len, flags):
    """What every rank does; returns (rdv, per-token counts, shared row arrays, timing) on rank 0."""
    engine.init(device)
    rdv = launch.Rendezvous(rank, world, "127.0.0.1", port)
    launch.init_comm(rdv, engine)
    tok_shm = _attach(tok_shm_name) if rank else None
    buf = np.frombuffer((tok_shm.buf if rank else _rank_body.parent_buf), dtype=np.uint8)
    plans = shard.plan([int(n) for n in tok_len], world)
    mine = plans[rank]
    slots = max(1, max(len(p) for p in plans))
    genome = engine.Genome()
    for k, a, b in mine:
        genome.add_segment(k, buf[tok_off[k]:tok_off[k] + tok_len[k]], a, b)
    genome.commit()
    engine.comm_barrier()                                          # shards differ in size: start the collective scan together
    res = genome.scan_sharded(slots, guide_len, flags)
    gathered = res.gathered_counts()                               # [world, 2, slots]: the all-gather's result
    counts = [(gathered[r, 0, :len(plans[r])], gathered[r, 1, :len(plans[r])]) for r in range(world)]
    offsets, total = shard.global_offsets(plans, counts)
    n_tok = len(tok_len)
    plus_tot = np.zeros(n_tok, dtype=np.uint64)
    minus_tot = np.zeros(n_tok, dtype=np.uint64)
    for r, segs in enumerate(plans):
        for s, (k, _, _) in enumerate(segs):
            plus_tot[k] += int(counts[r][0][s])
            minus_tot[k] += int(counts[r][1][s])
    # rank 0 owns the row arrays; their names go round once the total is known
    if rank == 0:
        pos_shm = shared_memory.SharedMemory(create=True, size=max(4 * total, 8))
        x_shm = shared_memory.SharedMemory(create=True, size=max(8 * total, 8))
        names = (pos_shm.name, x_shm.name)
    else:
        names = None
    names = rdv.broadcast(names)
    if rank:
        pos_shm, x_shm = _attach(names[0]), _attach(names[1])
    pos = np.frombuffer(pos_shm.buf, dtype=np.uint32, count=total)
    x = np.frombuffer(x_shm.buf, dtype=np.float64, count=total)
    scored = res.scored
    for s in range(len(mine)):
        for si, strand in enumerate("+-"):
            got = res.fetch_segment(s, strand, want=("pos", "x"))
            row0 = offsets[rank][s][si]
            n = len(got["pos"])
            pos[row0:row0 + n] = got["pos"]
            x[row0:row0 + n] = got["x"] if scored else np.nan
    fused, why = engine.comm_exchange_info()
    timing = rdv.all_gather({"rank": rank, "device": device, "fused_exchange": fused, "exchange_note": why,
                             **res.timing_detail(), **genome.timing(),
                             "positions": int(sum(b - a for _, a, b in mine))})
    res.free()
    genome.free()
    rdv.barrier()                                                  # every rank's rows are in place
    if rank:
        del pos, x, buf, got
        pos_shm.close()
        x_shm.close()
        tok_shm.close()
        engine.comm_shutdown()
        rdv.close()
        return None
    engine.comm_shutdown()
    rdv.close()
    return plus_tot, minus_tot, pos_shm, x_shm, pos, x, timing


def _worker(rank, world, port, device, tok_shm_name, tok_off, tok_len, guide_len, flags):
    _rank_body(rank, world, port, device, tok_shm_name, tok_off, tok_len, guide_len, flags)


class MultiScanOutput:
    """ScanOutput (pipeline.py) of a scan that ran on several GPUs: rows already in reference order."""

    def __init__(self, token_bytes, guide_len, plus_tot, minus_tot, pos_shm, x_shm, pos, x, timing):
        from .pipeline import HostRescorer
        self.token_bytes = token_bytes
        self.guide_len = guide_len
        self.seg_plus, self.seg_minus = plus_tot, minus_tot
        self._shm = (pos_shm, x_shm)
        self._pos, self._x = pos, x
        self.rank_timing = timing
        self._base = np.concatenate(([0], np.cumsum(plus_tot.astype(np.int64) + minus_tot.astype(np.int64))))
        self._host = HostRescorer(token_bytes)
        self._scored = guide_len == 20

    def token_rows(self, k):
        a = int(self._base[k])
        b = a + int(self.seg_plus[k])
        c = int(self._base[k + 1])
        xs = (lambda lo, hi: self._x[lo:hi]) if self._scored else (lambda lo, hi: None)
        return self._pos[a:b], xs(a, b), self._pos[b:c], xs(b, c)

    def rescore(self, tok_index, t, strand, cls):
        return self._host.rescore(tok_index, t, strand, cls)

    def scan_ms(self):
        return max(t["total_ms"] for t in self.rank_timing)         # kernel + all-gather, the slowest rank

    def timing(self):
        return {"h2d_ms": max(t["h2d_ms"] for t in self.rank_timing),
                "pack_ms": max(t["pack_ms"] for t in self.rank_timing), "ranks": self.rank_timing}

    def locate(self, k):
        raise NotImplementedError("the packed genome of a multi-GPU scan is gone once the rows are gathered")

    def free(self):
        self._pos = self._x = None
        for shm in self._shm:
            try:
                shm.unlink()                    # the name goes first: a view someone still holds only delays the unmap
            except Exception:
                pass
            try:
                shm.close()
            except BufferError:
                pass
        self._shm = ()


def scan_on_devices(token_bytes, devices, guide_len=20, flags=0):
    """Scan the tokens on the given CUDA devices (one process each; the calling process drives
    devices[0]).  -> MultiScanOutput."""
    world = len(devices)
    tok_len = np.array([len(b) for b in token_bytes], dtype=np.int64)
    tok_off = np.concatenate(([0], np.cumsum((tok_len + 127) // 128 * 128)))[:-1].astype(np.int64)
    total = int(((tok_len + 127) // 128 * 128).sum())
    tok_shm = shared_memory.SharedMemory(create=True, size=max(total, 8))
    parent_buf = tok_shm.buf
    view = np.frombuffer(parent_buf, dtype=np.uint8)
    for b, off in zip(token_bytes, tok_off.tolist()):
        view[off:off + len(b)] = np.frombuffer(b, dtype=np.uint8)
    port = _free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, world, port, devices[r], tok_shm.name, tok_off, tok_len, guide_len, flags),
                         daemon=True) for r in range(1, world)]
    for p in procs:
        p.start()
    ok = False
    try:
        _rank_body.parent_buf = parent_buf
        out = _rank_body(0, world, port, devices[0], tok_shm.name, tok_off, tok_len, guide_len, flags)
        ok = True
    finally:
        _rank_body.parent_buf = None
        for p in procs:
            p.join(timeout=120 if ok else 2)       # rank 0 failed: the others wait for it in vain
            if p.is_alive():
                p.terminate()
                p.join(timeout=10)
        del view
        tok_shm.close()
        tok_shm.unlink()
    bad = [p.exitcode for p in procs if p.exitcode != 0]
    if bad:
        raise RuntimeError(f"a GPU worker process failed (exit codes {bad})")
    return MultiScanOutput(token_bytes, guide_len, *out)
