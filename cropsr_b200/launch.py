"""One process per GPU: rank discovery and a small TCP rendezvous.

The library's only collective is an NCCL all-gather of per-segment candidate counts
(crp_scan_score_sharded); NCCL needs its 128-byte unique id handed from rank 0 to every other
rank before the communicator exists.  That hand-over -- plus the handful of host-side gathers a
launcher needs (shared-memory names, timings) -- goes over plain TCP on MASTER_ADDR:MASTER_PORT,
the variables torchrun / `python -m torch.distributed.run` export, so the same code runs under
torchrun and under the CLI's own process spawner (cropsr_b200/multi.py).  No PyTorch here.

The reference has no counterpart: CROPSR.py:409 is one serial loop over chromosomes.
"""
import os
import pickle
import socket
import struct
import time


def env_rank():
    """(rank, world, local_rank) from the launcher's environment (1-process defaults)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", os.environ.get("RANK", "0"))))


def _send(sock, obj):
    data = pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)
    sock.sendall(struct.pack("<Q", len(data)) + data)


def _recv(sock):
    hdr = b""
    while len(hdr) < 8:
        part = sock.recv(8 - len(hdr))
        if not part:
            raise ConnectionError("rendezvous peer closed the connection")
        hdr += part
    (n,) = struct.unpack("<Q", hdr)
    buf = bytearray()
    while len(buf) < n:
        part = sock.recv(min(1 << 20, n - len(buf)))
        if not part:
            raise ConnectionError("rendezvous peer closed the connection")
        buf += part
    return pickle.loads(bytes(buf))


class Rendezvous:
    """Star topology: rank 0 listens, every other rank keeps one connection to it.
    all_gather / broadcast / barrier are host-side and latency-bound (~0.1 ms on loopback):
    bootstrap and reporting only, never inside a timed region."""

    def __init__(self, rank, world, addr=None, port=None, timeout=300.0):
        self.rank, self.world = int(rank), int(world)
        self.peers = []
        self.sock = None
        if self.world == 1:
            return
        addr = addr or os.environ.get("MASTER_ADDR", "127.0.0.1")
        if addr == "localhost":
            addr = "127.0.0.1"
        base = int(port if port is not None else os.environ.get("MASTER_PORT", "29500"))
        # MASTER_PORT itself may be taken: torchrun's agent keeps its own store there.  Rank 0 listens on
        # the first free port of a fixed sequence derived from it; the others walk the same sequence and
        # keep the first listener that answers the handshake (so a foreign service on one of the ports,
        # or a rank 0 that is not up yet, only costs a retry).
        ports = [base] if port is not None else [1024 + (base + 997 + 61 * i) % 60000 for i in range(16)]
        hello = ("cropsr-rdv", base, self.world, os.environ.get("TORCHELASTIC_RUN_ID", ""))
        if self.rank == 0:
            srv = None
            for p in ports:
                try:
                    srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
                    srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
                    srv.bind((addr, p))
                    break
                except OSError:
                    srv.close()
                    srv = None
            if srv is None:
                raise OSError(f"rendezvous: none of the ports {ports} is free")
            srv.listen(self.world + 8)
            srv.settimeout(timeout)
            got = {}
            while len(got) < self.world - 1:
                conn, _ = srv.accept()
                conn.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
                conn.settimeout(timeout)
                try:
                    msg = _recv(conn)
                    if not (isinstance(msg, tuple) and msg[:-1] == hello):
                        raise ValueError("not a rank of this job")
                    _send(conn, hello)
                    got[int(msg[-1])] = conn
                except Exception:
                    conn.close()
            srv.close()
            self.peers = [got[r] for r in range(1, self.world)]
        else:
            deadline = time.time() + timeout
            s = None
            while s is None:
                for p in ports:
                    try:
                        c = socket.create_connection((addr, p), timeout=2.0)
                        c.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
                        c.settimeout(10.0)
                        _send(c, hello + (self.rank,))
                        if _recv(c) == hello:
                            s = c
                            break
                        c.close()
                    except Exception:
                        pass
                if s is None:
                    if time.time() > deadline:
                        raise TimeoutError("rendezvous: rank 0 did not answer")
                    time.sleep(0.05)
            s.settimeout(timeout)
            self.sock = s

    def all_gather(self, obj):
        """-> [obj of rank 0, obj of rank 1, ...] on every rank"""
        if self.world == 1:
            return [obj]
        if self.rank == 0:
            out = [obj] + [_recv(p) for p in self.peers]
            for p in self.peers:
                _send(p, out)
            return out
        _send(self.sock, obj)
        return _recv(self.sock)

    def broadcast(self, obj=None):
        """rank 0's obj on every rank"""
        return self.all_gather(obj if self.rank == 0 else None)[0]

    def barrier(self):
        self.all_gather(None)

    def close(self):
        for p in self.peers:
            p.close()
        if self.sock:
            self.sock.close()
        self.peers, self.sock = [], None


def init_comm(rdv, engine):
    """NCCL communicator of this process (engine.init must have bound the GPU): rank 0 draws the
    unique id, the rendezvous hands it round, every rank joins."""
    if rdv.world == 1:
        return
    uid = rdv.broadcast(engine.comm_unique_id() if rdv.rank == 0 else None)
    engine.comm_init(rdv.rank, rdv.world, uid)
