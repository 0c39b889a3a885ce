"""Several GPUs from ONE host process through the C ABI alone (no Python on the data path, no torch, no NCCL):
harness/test_multi_b200.cpp makes the reference's kind of call -- host buffers in, ints out (TestFileWithGPU.cpp:57-94)
-- after swb200_set_devices(G) for G = 1, 2, 4, 8 (as many as the box has) and demands identical results; this test
checks the scores it prints against the oracle.  With one GPU only G = 1 runs (the pool path itself is then covered by
test_set_devices_from_python)."""
import json
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oracle_lib as O
from concurrentproject_b200 import rng

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


def _write_inputs(tmp, n, m, npairs, l1, l2, seed):
    a = rng.random_acgt(seed, 0, n)
    b = rng.mutate(a, seed, 1, 0.10, 0.04)[:m]
    (tmp / "pair.bin").write_bytes(np.array([len(a), len(b)], dtype=np.int64).tobytes() + a.tobytes() + b.tobytes())
    if l1 == l2:
        pairs = [rng.long_pair(seed, k, l1) for k in range(npairs)]
    else:
        pairs = [rng.read_pair(seed, k, l1, l2) for k in range(npairs)]
    s1 = np.stack([p[0] for p in pairs]); s2 = np.stack([p[1] for p in pairs])
    (tmp / "batch.bin").write_bytes(np.array([npairs, l1, l2], dtype=np.int64).tobytes() + s1.tobytes() + s2.tobytes())
    return a, b, [p[0] for p in pairs], [p[1] for p in pairs]


def _run(tmp, *args):
    exe = ROOT / "harness" / "test_multi_b200"
    if not exe.exists():
        subprocess.run(["make", "-s", "-C", str(ROOT / "harness")], check=True)
    out = subprocess.run([str(exe), str(tmp / "pair.bin"), str(tmp / "batch.bin"), str(tmp / "out"), *map(str, args)],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    return [json.loads(line) for line in out.stdout.splitlines() if line.startswith("{")]


def test_cpp_caller_reaches_all_gpus_with_identical_results(tmp_path):
    a, b, s1, s2 = _write_inputs(tmp_path, 60000, 50000, 3001, 150, 1000, 31)
    lines = _run(tmp_path, 8, 1)
    want_pair = O.gotoh_fast(a, b)
    want_batch = O.gotoh_batch(s1, s2)
    assert lines and lines[0]["gpus"] == 1
    for ln in lines:
        assert ln["same_as_one_gpu"] is True and ln["devices_in_use"] == ln["gpus"]
        assert ln["pair_score"] == want_pair, ln
        got = np.fromfile(tmp_path / f"out_g{ln['gpus']}.bin", dtype=np.int32)
        assert np.array_equal(got, want_batch), ln["gpus"]


def test_cpp_caller_banded_batches_are_sharded_too(tmp_path):
    a, b, s1, s2 = _write_inputs(tmp_path, 3000, 3000, 257, 1000, 1000, 32)
    lines = _run(tmp_path, 8, 0)
    want_full = O.gotoh_batch(s1, s2)
    want_banded = O.gotoh_banded_batch(s1, s2, -32, 31)
    for ln in lines:
        got = np.fromfile(tmp_path / f"out_g{ln['gpus']}.bin", dtype=np.int32)
        assert np.array_equal(got[:257], want_full) and np.array_equal(got[257:], want_banded), ln["gpus"]


def test_set_devices_from_python():
    """swb200_set_devices from Python: more devices than present is refused, one device keeps answering; with two or
    more GPUs a pair as small as 5000 x 5000 is pushed over the in-process ring (ring_min_cells = 1) and must score the
    same, default and other parameters."""
    from concurrentproject_b200 import _lib, api
    have = _lib.load().swb200_device_count()
    with pytest.raises(api.SwbError):
        api.set_devices(have + 1)
    api.set_devices(1)
    a = rng.random_acgt(33, 0, 5000)
    b = rng.mutate(a, 33, 1, 0.1, 0.05)
    assert api.score(a, b) == O.gotoh_rolling(a, b)
    if have >= 2:
        api.configure("ring_min_cells", "1")
        try:
            api.set_devices(2)
            assert api.get_devices() == 2
            assert api.score(a, b) == O.gotoh_rolling(a, b)
            assert api.score(a, b, (2, -3, 5, 1)) == O.gotoh_rolling(a, b, (2, -3, 5, 1))
            assert api.last_run()["warps"] >= 8
            # host batches in the 2-bit format, sharded over the two devices (each shard's lengths scanned by its own thread)
            r = np.random.default_rng(5)
            s1 = [bytes(rng.random_acgt(34, k, int(r.integers(0, 150)))) for k in range(301)]
            s2 = [bytes(rng.mutate(np.frombuffer(x, np.uint8), 34, 400 + k, 0.06, 0.02)) if k % 2 and len(x) else
                  bytes(rng.random_acgt(34, 800 + k, int(r.integers(0, 700)))) for k, x in enumerate(s1)]
            f1, o1, l1 = api._flatten(s1); f2, o2, l2 = api._flatten(s2)
            qw, qs, tw, ts, ql, tl = api.pack_batch_host(f1, o1, l1, f2, o2, l2)
            assert np.array_equal(api.score_batch_packed(qw, qs, tw, ts, ql, tl), O.gotoh_batch(s1, s2))
            w1, st1, w2, st2 = api.pack_banded_host(f1, o1, l1, f2, o2, l2)
            assert np.array_equal(api.score_banded_batch_packed(w1, st1, w2, st2, l1, l2, -20, 43), O.gotoh_banded_batch(s1, s2, -20, 43))
        finally:
            api.set_devices(1)
            api.configure("ring_min_cells", str(200 * 10**9))
