#!/usr/bin/env python
"""End-to-end time of swb200_score_batch_packed (2-bit host words, cfg4-shaped pairs) against the copy/compute chunk size.
usage: python bench/batch_e2e.py [npairs] [read_len] [window_len]   -- one JSON line per chunk size (0 = the library's default)"""
import sys, json, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from concurrentproject_b200 import api
npairs = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
L1 = int(sys.argv[2]) if len(sys.argv) > 2 else 150
L2 = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
g = torch.Generator(device="cuda"); g.manual_seed(5)
lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device="cuda")
def seqs(L):
    return lut[torch.randint(0, 4, (npairs, L), device="cuda", generator=g)].cpu().numpy().reshape(-1)
f1, f2 = seqs(L1), seqs(L2)
o1 = np.arange(npairs, dtype=np.int64) * L1; o2 = np.arange(npairs, dtype=np.int64) * L2
l1 = np.full(npairs, L1, dtype=np.int32); l2 = np.full(npairs, L2, dtype=np.int32)
qw, qs, tw, ts, ql, tl = api.pack_batch_host(f1, o1, l1, f2, o2, l2)
pin = lambda x: torch.from_numpy(x).pin_memory().numpy()
qw, tw, ql, tl = pin(qw), pin(tw), pin(ql), pin(tl)
cells = float(npairs) * L1 * L2
want = None
for chunk in (0, 4 << 20, 8 << 20, 16 << 20, 24 << 20, 48 << 20, 96 << 20):
    api.configure("batch_chunk_bytes", str(chunk))
    api.score_batch_packed(qw, qs, tw, ts, ql, tl)
    t = []
    for _ in range(3):
        t0 = time.perf_counter(); out = api.score_batch_packed(qw, qs, tw, ts, ql, tl); t.append(time.perf_counter() - t0)
    if want is None: want = out
    print(json.dumps({"chunk_bytes": chunk, "ms": round(min(t) * 1e3, 2), "gcups": round(cells / min(t) / 1e9, 1), "same": bool(np.array_equal(out, want)),
                      "h2d_gb": round((qw.nbytes + tw.nbytes) / 1e9, 3)}), flush=True)
api.configure("batch_chunk_bytes", "0")
