"""Throughput of the FASTA / ragged-batch path (concurrentproject_b200/fasta.py) on one GPU: N read/window pairs with
ragged lengths, bucketed by length, host buffers in, scores out.  usage: python bench/fasta_bench.py [npairs]"""
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from concurrentproject_b200 import fasta, rng  # noqa: E402


def main():
    npairs = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    r = np.random.default_rng(1)
    rl = r.integers(50, 301, size=npairs).astype(np.int32)          # reads 50..300 bp
    wl = r.integers(400, 1501, size=npairs).astype(np.int32)        # windows 400..1500 bp
    ro = np.concatenate(([0], np.cumsum(rl[:-1], dtype=np.int64))); wo = np.concatenate(([0], np.cumsum(wl[:-1], dtype=np.int64)))
    reads = fasta.FastaRecords([""] * npairs, rng.random_acgt(31, 0, int(rl.sum())), ro.astype(np.int64), rl)
    wins = fasta.FastaRecords([""] * npairs, rng.random_acgt(31, 1, int(wl.sum())), wo.astype(np.int64), wl)
    cells = float((rl.astype(np.int64) * wl).sum())
    fasta.score_pairs(reads, wins)                                    # warm-up (allocations, first launch)
    t0 = time.perf_counter()
    s = fasta.score_pairs(reads, wins)
    dt = time.perf_counter() - t0
    print(json.dumps({"npairs": npairs, "cells": cells, "seconds": round(dt, 4), "gcups_from_host_records": round(cells / dt / 1e9, 1),
                      "bytes_in": int(rl.sum() + wl.sum()), "checksum": int(s.astype(np.int64).sum())}))


if __name__ == "__main__":
    main()
