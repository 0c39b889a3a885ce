#!/usr/bin/env python
"""Kernel-variant sweep on one GPU: GCUPS of the wavefront engine per (N, lanes, linear, rows, config).
Writes one JSON line per run to gpurun_out/sweep.jsonl.  Used to calibrate the planner in swb200.cu."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from concurrentproject_b200 import api, rng  # noqa: E402


def main():
    out = Path("gpurun_out"); out.mkdir(exist_ok=True)
    sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [3000, 20000, 100000]
    ctx = api.Context(0)
    with open(out / "sweep.jsonl", "a") as f:
        for n in sizes:
            a = torch.from_numpy(rng.random_acgt(2, 0, n).copy()).cuda()
            b = torch.from_numpy(rng.random_acgt(2, 1, n).copy()).cuda()
            for lanes, no_linear in ((16, False), (16, True), (32, True)):
                for config in (1, 2):
                    for rows in (1, 2, 3, 4, 6, 8, 12, 16):
                        if n >= 100000 and lanes == 32 and rows < 2:
                            continue
                        best = None
                        for rep in range(3):
                            s = ctx.score_device(a.data_ptr(), n, b.data_ptr(), n, lanes=lanes, rows=rows, config=config,
                                                 no_linear=no_linear)
                            info = ctx.last_run()
                            ms = info["engine_ms"]
                            best = ms if best is None else min(best, ms)
                        rec = {"n": n, "lanes": lanes, "linear": int(not no_linear), "rows": rows, "config": config,
                               "ctas": info["ctas"], "bands": info["bands"], "ms": round(best, 4),
                               "gcups": round(n * n / best / 1e6, 1), "score": s}
                        f.write(json.dumps(rec) + "\n"); f.flush()
                        print(rec, flush=True)


if __name__ == "__main__":
    main()
