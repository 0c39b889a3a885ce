#!/usr/bin/env python
"""Kernel-variant sweep on one GPU: GCUPS of the wavefront engine per (N, lanes, linear, rows, config).
usage: sweep.py SIZES [ROWS] [CONFIGS] [MODES]   e.g.  sweep.py 100000,1000000 4,8,16 1,2 aff,lin,s32
Appends one JSON line per run to gpurun_out/sweep.jsonl.  Used to calibrate the planner in swb200.cu."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from concurrentproject_b200 import api, rng  # noqa: E402


def main():
    out = Path("gpurun_out"); out.mkdir(exist_ok=True)
    sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [3000, 20000, 100000]
    rows_l = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 3, 4, 6, 8, 12, 16]
    cfgs = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 2]
    modes = sys.argv[4].split(",") if len(sys.argv) > 4 else ["lin", "aff", "s32"]
    reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
    ts = int(sys.argv[6]) if len(sys.argv) > 6 else 0        # two_sided option: 0 auto, 1 force, -1 never
    rb = int(sys.argv[7]) if len(sys.argv) > 7 else 0        # rebase option: 0 auto, 1 force, -1 never
    tag = sys.argv[8] if len(sys.argv) > 8 else ""
    mm = {"lin": (16, False), "aff": (16, True), "s32": (32, True)}
    ctx = api.Context(0)
    with open(out / "sweep.jsonl", "a") as f:
        for n in sizes:
            a = torch.from_numpy(rng.random_acgt(2, 0, n).copy()).cuda()
            b = torch.from_numpy(rng.random_acgt(2, 1, n).copy()).cuda()
            for mode in modes:
                lanes, no_linear = mm[mode]
                for config in cfgs:
                    for rows in rows_l:
                        best = None
                        try:
                            for rep in range(reps):
                                s = ctx.score_device(a.data_ptr(), n, b.data_ptr(), n, lanes=lanes, rows=rows, config=config,
                                                     no_linear=no_linear, two_sided=ts, rebase=rb)
                                info = ctx.last_run()
                                ms = info["engine_ms"]
                                best = ms if best is None else min(best, ms)
                        except Exception as e:
                            print("ERR", n, mode, config, rows, e, flush=True)
                            continue
                        rec = {"tag": tag, "n": n, "mode": mode, "ts": info["two_sided"], "rb": info["rebased"], "rows": rows, "config": config, "ctas": info["ctas"], "bands": info["bands"],
                               "ms": round(best, 4), "gcups": round(n * n / best / 1e6, 1), "score": s}
                        f.write(json.dumps(rec) + "\n"); f.flush()
                        print(rec, flush=True)


if __name__ == "__main__":
    main()
