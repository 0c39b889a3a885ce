#!/usr/bin/env python
"""Small inputs through every kernel family, each checked against the oracle: the program compute-sanitizer runs
(memcheck / racecheck / synccheck, one tool per call).  The engine hands boundary rows from warp to
warp through __syncwarp-ordered shared-memory rings and tagged relaxed global entries -- exactly what these tools are
for.  Sizes are tiny because the tools slow kernels down 10-100x.

  compute-sanitizer --tool memcheck python tests/sanitize_small.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import oracle_lib as O                                   # noqa: E402
from concurrentproject_b200 import api, rng              # noqa: E402


def planted(seed, n, sub=0.06, indel=0.03):
    a = rng.random_acgt(seed, 0, n)
    return a, rng.mutate(a, seed, 1, sub, indel)


checks = 0
a, b = planted(1, 1500)
want = O.gotoh_rolling(a, b)
for kw in (dict(), dict(no_linear=True), dict(lanes=32), dict(config=2, rows=2), dict(config=3, rows=2, no_linear=True),
           dict(config=4, rows=2), dict(config=4, rows=3, no_linear=True), dict(config=5, rows=2), dict(config=5, rows=1, no_linear=True, rebase=1), dict(rebase=1, rows=2), dict(rebase=1, no_linear=True, config=2, rows=2),
           dict(two_sided=1, rows=1), dict(two_sided=1, rows=1, no_linear=True, rebase=1), dict(two_sided=-1, rows=1, ctas=2)):
    got = api.score(a, b, **kw)
    assert got == want, (kw, got, want)
    checks += 1
p = (2, -3, 5, 1)
assert api.score(a, b, p) == O.gotoh_rolling(a, b, p); checks += 1
assert api.score(b"GATTACA", b"GCATGCU") == 2; checks += 1                     # five symbols: byte-compare kernel
assert api.score(b"ABDAAADB", b"ADDBAABB") == 2; checks += 1                   # remapped alphabet
assert api.score_end(a, b) == O.gotoh_end(a, b); checks += 1
assert api.score_span(a, b) == O.gotoh_span(a, b); checks += 1
reads, wins = zip(*[rng.read_pair(4, k, 150, 1000) for k in range(64)])
assert api.score_batch(list(reads), list(wins)).tolist() == O.gotoh_batch(list(reads), list(wins)).tolist(); checks += 1
assert api.score_batch(list(reads), list(wins), no_linear=True).tolist() == O.gotoh_batch(list(reads), list(wins)).tolist(); checks += 1
la, lb = zip(*[rng.long_pair(5, k, 700) for k in range(16)])
assert api.score_banded_batch(list(la), list(lb), -32, 31).tolist() == O.gotoh_banded_batch(list(la), list(lb), -32, 31).tolist(); checks += 1
assert api.score_banded_batch(list(la), list(lb), -32, 31, no_linear=True).tolist() == O.gotoh_banded_batch(list(la), list(lb), -32, 31).tolist(); checks += 1
import torch                                             # noqa: E402
out = torch.zeros(1000, dtype=torch.uint8, device="cuda")
api.gen_random_device(0, 2, 0, 1000, out.data_ptr())
torch.cuda.synchronize()
assert np.array_equal(out.cpu().numpy(), rng.random_acgt(2, 0, 1000)); checks += 1
print(f"sanitize_small: {checks} checks ok")
