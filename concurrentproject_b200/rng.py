"""Portable counter-based sequence generator.

The same function exists in C (oracle/gotoh_oracle.c: oracle_mix64 / oracle_random_acgt) and in
CUDA (csrc/swb_gen.cu: swb200_gen_*_device); all three are bit-identical, so the GPU path, the oracle
and the fixtures see the same bytes (tests/test_gpu_gen.py compares them byte for byte).  This replaces the reference harness's unseeded
``rand() % 4`` (TestFileWithGPU.cpp:25-36) and its libstdc++-specific
``mt19937_64 + uniform_int_distribution`` (cudaSmithM.cu:200-213).
"""
from __future__ import annotations

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
_NT = np.frombuffer(b"ACGT", dtype=np.uint8)


def mix64(seed: int, stream: int, index) -> np.ndarray:
    """splitmix64-style finaliser of (seed, stream, index); index may be an array."""
    idx = np.asarray(index, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = (np.uint64(seed & 0xFFFFFFFFFFFFFFFF)
             + np.uint64(0x9E3779B97F4A7C15) * (idx + np.uint64(1))
             + np.uint64((0xD1B54A32D192ED03 * (stream & 0xFFFFFFFFFFFFFFFF)) & 0xFFFFFFFFFFFFFFFF))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def random_codes(seed: int, stream: int, length: int) -> np.ndarray:
    """length symbols in {0,1,2,3}: word k//32 of the stream, bits 2*(k%32)."""
    if length <= 0:
        return np.zeros(0, dtype=np.uint8)
    nwords = (length + 31) // 32
    words = mix64(seed, stream, np.arange(nwords, dtype=np.uint64))
    shifts = (np.arange(32, dtype=np.uint64) * np.uint64(2))[None, :]
    codes = ((words[:, None] >> shifts) & np.uint64(3)).astype(np.uint8).reshape(-1)
    return codes[:length]


def random_acgt(seed: int, stream: int, length: int) -> np.ndarray:
    """ASCII bytes over {A,C,G,T}, identical to oracle_random_acgt(seed, stream, length)."""
    return _NT[random_codes(seed, stream, length)]


def mutate(seq: np.ndarray, seed: int, stream: int, sub_rate: float, indel_rate: float) -> np.ndarray:
    """A copy of ``seq`` with i.i.d. substitutions and single-base insertions/deletions.

    Used for planted-similarity pairs (cfg4 reads, cfg5 long reads).  Decisions come from
    mix64(seed, stream, position) so the result is reproducible from (seed, stream) alone.
    """
    if isinstance(seq, (bytes, bytearray)):
        seq = np.frombuffer(bytes(seq), dtype=np.uint8)
    n = len(seq)
    if n == 0:
        return seq.copy()
    r = mix64(seed, stream, np.arange(n, dtype=np.uint64))
    u = (r >> np.uint64(11)).astype(np.float64) / float(1 << 53)
    newc = ((r >> np.uint64(3)) & np.uint64(3)).astype(np.uint8)
    out = []
    sub_hi = sub_rate
    del_hi = sub_hi + indel_rate / 2.0
    ins_hi = del_hi + indel_rate / 2.0
    kind = np.zeros(n, dtype=np.uint8)          # 0 keep, 1 substitute, 2 delete, 3 insert-after
    kind[u < ins_hi] = 3
    kind[u < del_hi] = 2
    kind[u < sub_hi] = 1
    sub = _NT[newc]
    # make substitutions real changes
    same = (kind == 1) & (sub == seq)
    sub = np.where(same, _NT[(newc + 1) & 3], sub)
    keep = kind != 2
    base = np.where(kind == 1, sub, seq)
    # build with insertions
    counts = np.where(keep, 1, 0) + np.where(kind == 3, 1, 0)
    total = int(counts.sum())
    res = np.empty(total, dtype=np.uint8)
    pos = np.cumsum(counts) - counts
    res[pos[keep]] = base[keep]
    ins = kind == 3
    res[pos[ins] + 1] = _NT[(newc[ins] + 2) & 3]
    return res


# ---- the synthetic inputs of BASELINE configs 4 and 5, one pair at a time --------------------
# csrc/swb_gen.cu generates whole batches of these in HBM; the functions below regenerate any single pair (by its
# global pair id) on the host, for the oracle sample of bench.py and for tests/test_gpu_gen.py.

def read_pair(seed: int, pair_id: int, read_len: int = 150, window_len: int = 1000):
    """(read, window) of BASELINE config 4: the window is window_len random bases (stream 2*pair_id); even pairs carry
    a read cut from the window (offset mix64(seed, 2*pair_id+1, 2**40) % (window_len - read_len - 15)) with 5 %
    substitutions and 1 % indels, odd pairs a read of read_len unrelated random bases (stream 2*pair_id+1)."""
    window = random_acgt(seed, 2 * pair_id, window_len)
    if pair_id % 2 == 0:
        off = int(mix64(seed, 2 * pair_id + 1, np.uint64(1 << 40))) % (window_len - read_len - 15)
        read = _first(mutate(window[off:off + read_len + 16], seed, 2 * pair_id + 1, 0.05, 0.01), read_len)
    else:
        read = random_acgt(seed, 2 * pair_id + 1, read_len)
    return read, window


def long_pair(seed: int, pair_id: int, length: int = 10000):
    """(seq1, seq2) of BASELINE config 5: seq1 = length random bases (stream 2*pair_id); seq2 = the same stretch (plus
    320 spare bases) with 10 % substitutions and 2 % single-base indels, cut to length."""
    src = random_acgt(seed, 2 * pair_id, length + 320)
    return src[:length].copy(), _first(mutate(src, seed, 2 * pair_id + 1, 0.10, 0.02), length)


def _first(a: np.ndarray, n: int) -> np.ndarray:
    if len(a) >= n:
        return a[:n].copy()
    return np.concatenate([a, np.full(n - len(a), ord("A"), dtype=np.uint8)])
