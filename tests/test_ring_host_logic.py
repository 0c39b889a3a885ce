"""world_size-2 gloo test of the host-side ring orchestration (concurrentproject_b200/ring.py): handle exchange,
neighbour wiring, max/OR reduction of the ranks' partial results, the collective retry with re-based lanes when the score leaves the s16 range.
The CUDA ring end points are replaced by stand-ins that behave like the C ABI (no GPU here); the real kernels
are covered by tests/test_gpu_parity.py::test_ring_of_virtual_ranks_on_one_gpu and the multi-GPU bench."""
import os
import sys
from pathlib import Path

import pytest
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


class FakeCtx:
    def __init__(self, device):
        self.device = device

    def last_run(self):
        return {}

    def close(self):
        pass


class FakeRing:
    """Mimics swb200_ring_*: a 64-byte handle naming the rank; partial() returns a rank-dependent share."""
    log = []

    def __init__(self, ctx, rank, world, max_len):
        import ctypes as C
        self.rank, self.world = rank, world
        self.ipc = (C.c_char * 64)()
        self.ipc.raw = (b"rank%02d" % rank).ljust(64, b"\0")
        self.next = None

    def connect_ipc(self, h):
        self.next = int(h[4:6])

    def connect_local(self, other):
        self.next = other.rank

    def connect_root_ipc(self, h):
        self.root = int(h[4:6])

    def combine_pending(self):
        return self.two_sided            # same on every rank, like the real plan

    def combine(self, stream=0):
        self.two_sided = False
        return 7777 if self.rank == 0 else 0        # only the root holds the two middle rows

    two_sided = False

    def partial(self, d1, n, d2, m, params, *, lanes, rebase=0, stream=0, **kw):
        FakeRing.log.append((lanes, rebase))
        self.two_sided = n == 55                    # n == 55: a pair the plan sweeps from both ends; its best alignment crosses the middle
        true_score = 40000 if n == 99 else 1234          # n == 99: a pair whose score leaves the s16 range
        if lanes == 16 and rebase < 0 and true_score > 32000:
            return (32700, 1) if self.rank == 1 else (17, 0)     # only ONE rank notices the overflow
        return (true_score if self.rank == self.world - 1 else 5 * self.rank, 0)

    def close(self):
        pass


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, str(ROOT))
    import torch.distributed as dist
    from concurrentproject_b200.ring import DistributedRingAligner
    dist.init_process_group("gloo", rank=rank, world_size=world)
    al = DistributedRingAligner(0, 1000, _ctx_factory=FakeCtx, _ring_factory=FakeRing)
    out = {"rank": rank, "next": al.ring.next}
    out["plain"] = al.score(0, 10, 0, 10)
    out["root"] = al.ring.root
    out["crossing"] = al.score(0, 55, 0, 55)
    FakeRing.log.clear()
    out["overflow"] = al.score(0, 99, 0, 99)
    out["widths"] = list(FakeRing.log)
    try:
        al.score(0, 99, 0, 99, lanes=16)
        out["forced16"] = "no error"
    except RuntimeError as e:
        out["forced16"] = str(e)
    q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_ring_orchestration_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in procs), key=lambda d: d["rank"])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r["next"] for r in res] == [1, 0]                       # each rank mapped its successor's buffer
    assert [r["plain"] for r in res] == [1234, 1234]                # max over partial scores, same on every rank
    assert [r["root"] for r in res] == [0, 0]                       # every rank mapped rank 0's region (two-sided sweeps)
    assert [r["crossing"] for r in res] == [7777, 7777]             # rank 0's combination result reaches every rank
    assert [r["overflow"] for r in res] == [40000, 40000]           # every rank repeated the call ...
    assert [r["widths"] for r in res] == [[(16, -1), (16, 1)]] * 2       # ... (re-based lanes) although only rank 1 saw the overflow
    assert all("16-bit" in r["forced16"] for r in res)
