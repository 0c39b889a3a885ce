#!/usr/bin/env python
"""End-to-end time of swb200_score_banded_batch_packed (2-bit host words) against the copy/compute chunk size.
usage: python bench/banded_e2e.py [npairs] [len]   -- one JSON line per chunk size (0 = the library's default)"""
import sys, json, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from concurrentproject_b200 import api
npairs = int(sys.argv[1]) if len(sys.argv) > 1 else 250000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
g = torch.Generator(device="cuda"); g.manual_seed(5)
lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device="cuda")
def seqs():
    out = torch.empty((npairs, L), dtype=torch.uint8, device="cuda")
    for k in range(0, npairs, 25000):
        out[k:k + 25000] = lut[torch.randint(0, 4, (min(25000, npairs - k), L), device="cuda", generator=g)]
    return out.cpu().numpy().reshape(-1)
f1, f2 = seqs(), seqs()
off = (np.arange(npairs, dtype=np.int64) * L); ln = np.full(npairs, L, dtype=np.int32)
w1, s1, w2, s2 = api.pack_banded_host(f1, off, ln, f2, off, ln)
pin = lambda x: torch.from_numpy(x).pin_memory().numpy()
w1, w2, ln = pin(w1), pin(w2), pin(ln)
cells = float(npairs) * 64 * L
want = None
for chunk in (0, 16 << 20, 96 << 20, 144 << 20, 288 << 20, 1 << 40):
    api.configure("batch_chunk_bytes", str(chunk))
    api.score_banded_batch_packed(w1, s1, w2, s2, ln, ln)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); out = api.score_banded_batch_packed(w1, s1, w2, s2, ln, ln); ts.append(time.perf_counter() - t0)
    if want is None: want = out
    print(json.dumps({"chunk_bytes": chunk, "ms": round(min(ts) * 1e3, 2), "gcups": round(cells / min(ts) / 1e9, 1), "same": bool(np.array_equal(out, want)),
                      "h2d_gb": round((w1.nbytes + w2.nbytes) / 1e9, 3)}), flush=True)
api.configure("batch_chunk_bytes", "0")
