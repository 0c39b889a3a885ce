#!/usr/bin/env python
"""One-GPU diagnostics of the pair engine (not a benchmark): where a cfg2-sized run spends its time.
  python bench/diag.py [tag]
Writes gpurun_out/diag_<tag>.jsonl:
  * "pure": every band started at once with boundary stores and polls switched off (swb200_configure dbg=3; scores are
            WRONG by construction) -> cycles per step of the bare step loop + chunk prologue, per (mode, R, config)
  * "prof": the real run with per-warp counters and %globaltimer stamps -> per-band start lag, prologue / step cycles
  * "sizes": engine ms and score of cfg2 / 1 M / (optionally) cfg3 pairs on one GPU
"""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from concurrentproject_b200 import api, rng  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "x"
what = sys.argv[2].split(",") if len(sys.argv) > 2 else ["pure", "prof", "sizes"]
out = ROOT / "gpurun_out"
out.mkdir(exist_ok=True)
f = open(out / f"diag_{tag}.jsonl", "a")
ctx = api.Context(0)
MHZ = 1965.0


def emit(rec):
    f.write(json.dumps(rec) + "\n"); f.flush()
    print(rec, flush=True)


def pair(n, seed):
    a = torch.from_numpy(rng.random_acgt(seed, 0, n).copy()).cuda()
    b = torch.from_numpy(rng.random_acgt(seed, 1, n).copy()).cuda()
    return a, b


n = 100000
a, b = pair(n, 2)
if "pure" in what:
    api.configure("dbg", 3)
    for mode, no_lin in (("lin", False), ("aff", True)):
        for config in (1, 3, 2):
            for R in (2, 3, 4, 6, 8):
                best = None
                for _ in range(3):
                    ctx.score_device(a.data_ptr(), n, b.data_ptr(), n, lanes=16, rows=R, config=config, no_linear=no_lin,
                                     two_sided=-1, rebase=-1)
                    info = ctx.last_run()
                    best = info["engine_ms"] if best is None else min(best, info["engine_ms"])
                skew = 31 * (2 + (1 if config == 1 else 0)) + 1
                nsteps = (n + skew + 31) // 32 * 32
                rounds = -(-info["bands"] // info["warps"])
                emit({"kind": "pure", "mode": mode, "R": R, "config": config, "ms": round(best, 4), "bands": info["bands"],
                      "warps": info["warps"], "rounds": rounds,
                      "cyc_per_step": round(best * 1e-3 * MHZ * 1e6 / (nsteps * rounds), 2)})
    api.configure("dbg", 0)

if "cfgsweep" in what:
    configs = [int(x) for x in os.environ.get("DIAG_CONFIGS", "1,4,5,6").split(",")]
    rows = [int(x) for x in os.environ.get("DIAG_ROWS", "2,3,4,6").split(",")]
    for nn, seed, want in ((100000, 2, 11446), (1000000, 6, 114366)):
        if nn > 100000 and "big" not in what:
            continue
        x, y = pair(nn, seed)
        for mode, no_lin in (("lin", False), ("aff", True)):
            for config in configs:
                for R in (rows if nn == 100000 else [10, 12, 14, 16]):
                    for pure in ((0, 3) if nn == 100000 else (0,)):
                        api.configure("dbg", pure)
                        best = None
                        try:
                            for _ in range(3 if nn == 100000 else 2):
                                s = ctx.score_device(x.data_ptr(), nn, y.data_ptr(), nn, rows=R, config=config, no_linear=no_lin,
                                                     two_sided=(-1 if pure else 1), rebase=int(os.environ.get("DIAG_REBASE", "0")))
                                info = ctx.last_run()
                                best = info["engine_ms"] if best is None else min(best, info["engine_ms"])
                        except Exception as e:
                            emit({"kind": "cfgsweep", "n": nn, "mode": mode, "R": R, "config": config, "pure": pure, "error": str(e)[:200]})
                            continue
                        finally:
                            api.configure("dbg", 0)
                        rec = {"kind": "cfgsweep", "n": nn, "mode": mode, "R": R, "config": config, "pure": pure, "ms": round(best, 4),
                               "bands": info["bands"], "warps": info["warps"], "rebased": info["rebased"], "ok": bool(pure or s == want)}
                        if pure:
                            slack, hs = {1: (1, 0), 4: (1, 1), 5: (1, 0)}.get(config, (0, 0))
                            nsteps = (nn + 31 * (2 + slack + hs) + 1 + hs + 255) // 256 * 256
                            rounds = -(-info["bands"] // info["warps"])
                            rec["cyc_per_step"] = round(best * 1e-3 * MHZ * 1e6 / (nsteps * rounds), 2)
                        else:
                            rec["gcups"] = round(nn * nn / best / 1e6, 1)
                        emit(rec)

if "prof" in what:
    pf = out / f"prof_{tag}.jsonl"
    if pf.exists():
        pf.unlink()
    for kw in [json.loads(x) for x in os.environ.get("DIAG_PROF", '{}|{"rows": 3}|{"rows": 4, "config": 3}|{"no_linear": true}').split("|")]:
        ctx.score_device(a.data_ptr(), n, b.data_ptr(), n, **kw)     # warm
        api.configure("prof", str(pf))
        s = ctx.score_device(a.data_ptr(), n, b.data_ptr(), n, **kw)
        api.configure("prof", "")
        info = ctx.last_run()
        line = json.loads(pf.read_text().strip().split("\n")[-1])
        w = np.array(line["warps"], dtype=np.float64)
        split = line["split"] or len(w)
        for half, ws in (("fwd", w[:split]), ("rev", w[split:])):
            ws = ws[ws[:, 3] > 0]
            if len(ws) < 3:
                continue
            t0 = ws[:, 4] - ws[:, 4].min()
            lag_ns = np.diff(ws[:, 4])
            emit({"kind": "prof", "kw": kw, "half": half, "score": s, "ms_prof": line["ms"], "R": info["rows"], "config": info["config"],
                  "bands": int(len(ws)), "prologue_cyc_per_chunk": round(float(np.median(ws[:, 0] / ws[:, 3])), 1),
                  "steps_cyc_per_chunk": round(float(np.median(ws[:, 1] / ws[:, 3])), 1),
                  "failed_polls_per_chunk": round(float(np.mean(ws[:, 2] / ws[:, 3])), 3),
                  "band_start_lag_ns_median": float(np.median(lag_ns)), "band_start_lag_ns_mean": float(np.mean(lag_ns)),
                  "last_band_start_us": float(t0.max() / 1e3),
                  "band_duration_us_median": float(np.median(ws[:, 5] - ws[:, 4]) / 1e3),
                  "first_band_duration_us": float((ws[0, 5] - ws[0, 4]) / 1e3),
                  "last_band_duration_us": float((ws[-1, 5] - ws[-1, 4]) / 1e3)})

if "sizes" in what:
    golden = json.loads((ROOT / "tests" / "golden" / "large_scores.json").read_text()) if (ROOT / "tests" / "golden" / "large_scores.json").exists() else {}
    for name, nn, seed in (("cfg2", 100000, 2), ("ring400k", 400000, 7), ("n1m", 1000000, 6)) + ((("cfg3", 4000000, 3),) if "cfg3" in what else ()):
        x, y = pair(nn, seed)
        for kw in ({}, {"no_linear": True}):
            best = None
            for _ in range(2):
                s = ctx.score_device(x.data_ptr(), nn, y.data_ptr(), nn, **kw)
                info = ctx.last_run()
                best = info["engine_ms"] if best is None else min(best, info["engine_ms"])
            emit({"kind": "size", "name": name, "n": nn, "kw": kw, "score": s, "golden": golden.get(name, {}).get("score"),
                  "ms": round(best, 3), "gcups": round(nn * nn / best / 1e6, 1), "info": info})
