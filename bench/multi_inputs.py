#!/usr/bin/env python
"""Writes the input files of harness/test_multi_b200 (the C++ multi-GPU caller): python bench/multi_inputs.py DIR [pair_n] [npairs] [len1] [len2]
pair.bin = one seeded random pair (BASELINE config 3 for pair_n = 4000000, seed 3); batch.bin = pairs of the config-4 recipe."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from concurrentproject_b200 import rng  # noqa: E402

d = Path(sys.argv[1]); d.mkdir(parents=True, exist_ok=True)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4000000
npairs = int(sys.argv[3]) if len(sys.argv) > 3 else 200000
l1 = int(sys.argv[4]) if len(sys.argv) > 4 else 150
l2 = int(sys.argv[5]) if len(sys.argv) > 5 else 1000
seed = {4000000: 3, 1000000: 6, 400000: 7, 100000: 2}.get(n, 9)
a, b = rng.random_acgt(seed, 0, n), rng.random_acgt(seed, 1, n)
(d / "pair.bin").write_bytes(np.array([n, n], dtype=np.int64).tobytes() + a.tobytes() + b.tobytes())
# whole windows / reads straight from the generator's streams (vectorised; the per-pair recipe is in rng.read_pair)
s1 = np.stack([rng.random_acgt(4, 2 * k + 1, l1) for k in range(npairs)]) if npairs <= 20000 else \
    rng.random_acgt(4, 1, npairs * l1).reshape(npairs, l1)
s2 = np.stack([rng.random_acgt(4, 2 * k, l2) for k in range(npairs)]) if npairs <= 20000 else \
    rng.random_acgt(4, 0, npairs * l2).reshape(npairs, l2)
if npairs > 20000:          # plant every even read into its window so that scores are spread out
    s1[::2] = s2[::2, 100:100 + l1]
    s1[::2, ::17] = ord("A")
(d / "batch.bin").write_bytes(np.array([npairs, l1, l2], dtype=np.int64).tobytes() + s1.tobytes() + s2.tobytes())
print("wrote", d / "pair.bin", d / "batch.bin")
