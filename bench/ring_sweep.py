#!/usr/bin/env python
"""Kernel-variant sweep of the multi-GPU ring (one long pair over all ranks), launched with torchrun like bench.py:
  python -m torch.distributed.run --nproc-per-node N ... bench/ring_sweep.py [workload] [variants]
variants: comma-separated rows:config:two_sided triples, e.g. 14:3:1,14:2:1,10:2:0 (two_sided: 1 force, -1 never, 0 auto).
Rank 0 appends one JSON line per variant to gpurun_out/ring_sweep.jsonl: kernel ms (max over ranks), wall ms, GCUPS.
Used to fit the planner (swb200.cu: estimate) for the multi-GPU regime; not a benchmark."""
import json
import os
import sys
import time
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from concurrentproject_b200 import rng                                  # noqa: E402
from concurrentproject_b200.ring import DistributedRingAligner          # noqa: E402

WORK = {"cfg3": (4000000, 3, 456968), "ring1m": (1000000, 6, 114366), "ring400k": (400000, 7, 45661)}
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
variants = [tuple(int(x) for x in v.split(":")) for v in (sys.argv[2] if len(sys.argv) > 2 else "0:0:0").split(",")]
n, seed, want = WORK[name]
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
a = torch.from_numpy(rng.random_acgt(seed, 0, n).copy()).cuda()
b = torch.from_numpy(rng.random_acgt(seed, 1, n).copy()).cuda()
al = DistributedRingAligner(local, n)
out = ROOT / "gpurun_out"
out.mkdir(exist_ok=True)
for rows, config, ts in variants:
    best_k, best_w, info, s = None, None, {}, None
    try:
        for rep in range(3):
            dist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            s = al.score(a.data_ptr(), n, b.data_ptr(), n, rows=rows, config=config, two_sided=ts)
            torch.cuda.synchronize()
            w_ms = (time.perf_counter() - t0) * 1e3
            info = al.last_run()
            t = torch.tensor([info["engine_ms"], w_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if rep > 0:
                best_k = float(t[0]) if best_k is None else min(best_k, float(t[0]))
                best_w = float(t[1]) if best_w is None else min(best_w, float(t[1]))
    except Exception as e:
        if rank == 0:
            print(json.dumps({"workload": name, "gpus": world, "asked": [rows, config, ts], "error": str(e)[:200]}), flush=True)
        continue
    if rank == 0:
        rec = {"workload": name, "gpus": world, "asked": [rows, config, ts], "rows": info["rows"], "config": info["config"],
               "two_sided": info["two_sided"], "rebased": info["rebased"], "bands": info["bands"], "warps_per_gpu": info["warps"],
               "kernel_ms": round(best_k, 3), "wall_ms": round(best_w, 3), "gcups": round(n * n / best_w / 1e6, 1), "score_ok": s == want}
        with open(out / "ring_sweep.jsonl", "a") as f:
            f.write(json.dumps(rec) + "\n")
        print(json.dumps(rec), flush=True)
al.close()
dist.barrier()
dist.destroy_process_group()
