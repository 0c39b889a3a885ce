"""One very long pair spread over a ring of GPUs, one process per GPU (BASELINE config 3).

Host-side orchestration only: the boundary hand-off between GPUs happens inside the wavefront kernel
(peer stores into the next GPU's memory, include/swb200.h "ring").  torch.distributed is used for the
plumbing the C ABI leaves to the caller: exchanging the 64-byte CUDA IPC handles once, and the final
max over the ranks' partial scores (one 3-int all-reduce per call; no data-path collective)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

from . import _lib
from .api import DEFAULT_PARAMS, Context, SwbError, _options, _params

STATUS_S16_OVERFLOW, STATUS_TIMEOUT, STATUS_REBASE_RANGE = 1, 2, 8


class Ring:
    """One rank's end of the ring.  `exchange(handle_bytes) -> list[bytes]` is any all-gather of the
    64-byte handles (torch.distributed.all_gather_object by default)."""

    def __init__(self, ctx: Context, rank: int, world: int, max_stream_len: int):
        self.ctx, self.rank, self.world = ctx, rank, world
        self.handle = C.c_void_p()
        self.ipc = (C.c_char * 64)()
        rc = _lib.load().swb200_ring_create(ctx.handle, rank, world, max_stream_len, C.byref(self.handle), self.ipc)
        if rc != 0:
            raise SwbError(rc, "swb200_ring_create")

    def connect_ipc(self, next_handle: bytes):
        buf = (C.c_char * 64).from_buffer_copy(next_handle)
        rc = _lib.load().swb200_ring_connect(self.handle, buf)
        if rc != 0:
            raise SwbError(rc, "swb200_ring_connect")

    def connect_local(self, nxt: "Ring"):
        rc = _lib.load().swb200_ring_connect_local(self.handle, nxt.handle)
        if rc != 0:
            raise SwbError(rc, "swb200_ring_connect_local")

    def connect_root_ipc(self, root_handle: bytes):
        """Maps rank 0's region too: enables the two-sided sweep over the ring (swb200_ring_connect_root)."""
        buf = (C.c_char * 64).from_buffer_copy(root_handle)
        rc = _lib.load().swb200_ring_connect_root(self.handle, buf)
        if rc != 0:
            raise SwbError(rc, "swb200_ring_connect_root")

    def connect_root_local(self, root: "Ring"):
        rc = _lib.load().swb200_ring_connect_root_local(self.handle, root.handle)
        if rc != 0:
            raise SwbError(rc, "swb200_ring_connect_root_local")

    def combine_pending(self) -> bool:
        return bool(_lib.load().swb200_ring_combine_pending(self.handle))

    def combine(self, stream: int = 0) -> int:
        """After EVERY rank's partial() has returned: the best alignment crossing the middle row (rank 0; 0 elsewhere)."""
        out = C.c_int(0)
        rc = _lib.load().swb200_ring_combine(self.handle, C.c_void_p(stream), C.byref(out))
        if rc != 0:
            raise SwbError(rc, "swb200_ring_combine")
        return out.value

    def partial(self, d_seq1: int, n: int, d_seq2: int, m: int, params: Sequence[int] = DEFAULT_PARAMS, *, lanes: int,
                stream: int = 0, rows: int = 0, config: int = 0, ctas: int = 0, no_linear: bool = False, rebase: int = 0,
                two_sided: int = 0):
        """This rank's share of one collective call: (partial best score, status bits)."""
        score, status = C.c_int(0), C.c_int(0)
        p, o = _params(params), _options(lanes, rows, config, ctas, no_linear, 0, rebase, two_sided)
        rc = _lib.load().swb200_ring_score_device(self.handle, C.c_void_p(d_seq1), n, C.c_void_p(d_seq2), m, C.byref(p),
                                                  C.byref(o), C.c_void_p(stream), C.byref(score), C.byref(status))
        if rc != 0:
            raise SwbError(rc, "swb200_ring_score_device")
        return score.value, status.value

    def close(self):
        if self.handle:
            _lib.load().swb200_ring_destroy(self.handle)
            self.handle = C.c_void_p()


class DistributedRingAligner:
    """torch.distributed front end: rank r of the default (or given) process group drives GPU `device`."""

    def __init__(self, device: int, max_stream_len: int, group=None, _ctx_factory=Context, _ring_factory=Ring):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = device
        # the 3-int reduction travels on whatever the process group runs on (NCCL: the GPU; gloo in CPU tests: host)
        self.reduce_device = f"cuda:{device}" if dist.get_backend(group) == "nccl" else "cpu"
        self.ctx = _ctx_factory(device)
        self.ring = _ring_factory(self.ctx, self.rank, self.world, max_stream_len)
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(self.ring.ipc.raw), group=group)
        if self.world > 1:
            self.ring.connect_ipc(handles[(self.rank + 1) % self.world])
            self.ring.connect_root_ipc(handles[0])           # two-sided sweeps: the two middle rows meet on rank 0
        dist.barrier(group=group)

    def score(self, d_seq1: int, n: int, d_seq2: int, m: int, params: Sequence[int] = DEFAULT_PARAMS, *, lanes: int = 0,
              stream: int = 0, **opts) -> int:
        """Collective.  Lane-width policy as in swb200_score: plain 16-bit lanes while match*min(n,m) is within 8x
        the s16 range (the kernel reports leaving it), then re-based 16-bit lanes, 32-bit lanes as the last resort.
        Every rank takes the same decisions because they depend only on the arguments and on all-reduced flags."""
        torch, dist = self.torch, self.dist
        bound = int(params[0]) * min(n, m)
        if lanes == 32:
            attempts = [(32, -1)]
        elif lanes == 16:
            attempts = [(16, -1)]
        else:
            attempts = ([(16, -1)] if bound <= 8 * 32767 else []) + [(16, 1), (32, -1)]
        for width, rebase in attempts:
            try:
                part, status = self.ring.partial(d_seq1, n, d_seq2, m, params, lanes=width, rebase=rebase, stream=stream, **opts)
            except SwbError as e:
                if rebase == 1 and e.code == -2:      # re-based lanes not safe for these parameters: same on every rank
                    continue
                raise
            if self.ring.combine_pending():      # same answer on every rank: the plan depends only on the arguments
                dist.barrier(group=self.group)   # every rank's kernel has finished: the middle rows are complete on rank 0
                part = max(part, self.ring.combine(stream))
            t = torch.tensor([part, status & STATUS_S16_OVERFLOW, status & STATUS_TIMEOUT, status & STATUS_REBASE_RANGE],
                             dtype=torch.int32, device=self.reduce_device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            best, overflow, timeout, rb_range = (int(x) for x in t.tolist())
            if timeout:
                raise RuntimeError("ring hand-off timed out on some rank")
            if not overflow and not rb_range:
                return best
            if lanes == 16:
                raise RuntimeError("score leaves the 16-bit lane range")
        raise RuntimeError("no lane width could score this pair")

    def last_run(self) -> dict:
        return self.ctx.last_run()

    def close(self):
        self.ring.close()
        self.ctx.close()
