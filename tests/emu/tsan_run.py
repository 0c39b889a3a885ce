#!/usr/bin/env python
"""Race check of the hand-off protocol WITHOUT a GPU tool (compute-sanitizer is closed on the GPU pool, see
profiles/r02_compute_sanitizer_closed.txt): the CPU emulator (one pthread per lane, a barrier per warp-synchronous point,
atomics for the tagged entries -- the very engine source the kernels compile) built with -fsanitize=thread.  Every
shared-memory ring access that is not ordered by a __syncwarp-equivalent barrier, and every plain access to a boundary
entry, would be reported as a data race.   python tests/emu/tsan_run.py  -> profiles-style summary on stdout"""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import oracle_lib as O                            # noqa: E402
from concurrentproject_b200 import rng            # noqa: E402

EMU = ROOT / "tests" / "emu"
exe = EMU / "engine_emu_tsan"
subprocess.run(["nvcc", "-O1", "-g", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fsanitize=thread",
                "-o", str(exe), str(EMU / "engine_emu.cu"), "-lpthread", "-ltsan"], check=True, cwd=EMU)
a = rng.random_acgt(5, 0, 400)
b = rng.mutate(a, 5, 1, 0.06, 0.03)
Path("/tmp/tsan_q.bin").write_bytes(bytes(a)); Path("/tmp/tsan_t.bin").write_bytes(bytes(b))
res = []
for (R, mode, slack, W, G, hs, p) in [(1, 0, 1, 3, 1, 0, (1, -1, 1, 1)), (1, 1, 0, 2, 2, 0, (1, -1, 1, 1)), (1, 3, 1, 2, 1, 1, (10, -8, 10, 5)),
                                       (1, 2, 1, 2, 1, 0, (2, -3, 5, 1)), (2, 4, 1, 2, 2, 0, (10, -8, 7, 7))]:
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0")
    if hs:
        env["EMU_HS"] = "1"
    out = subprocess.run([str(exe), "/tmp/tsan_q.bin", "/tmp/tsan_t.bin", *map(str, [R, mode, slack, W, G, 5, *p, 4096])],
                         capture_output=True, text=True, env=env, timeout=3000)
    d = dict(kv.split("=") for kv in out.stdout.split())
    rec = {"R": R, "mode": mode, "slack": slack, "warps": W, "gpus": G, "hs": hs, "params": p, "score": int(d["score"]),
           "oracle": O.gotoh_rolling(a, b, p), "status": int(d["status"]), "tsan_reports": out.stderr.count("WARNING: ThreadSanitizer")}
    res.append(rec)
    print(json.dumps(rec), flush=True)
    if rec["tsan_reports"]:
        print(out.stderr[:4000])
assert all(r["score"] == r["oracle"] and r["status"] == 0 and r["tsan_reports"] == 0 for r in res)
