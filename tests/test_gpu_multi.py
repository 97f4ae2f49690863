"""Receiver sharding: several ranks - here several contexts, each with its own
stream, connected in-process - must shoot the same patches in the same order as one GPU and end with
bit-identical radiosity on their shards.  The exchange runs inside the kernels through the peer
mailboxes, exactly as between processes; tests/multi_gpu_check.py repeats the check across real GPUs
with CUDA IPC under torchrun."""
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def run_ranks(hr, m, world, steps, devices=None):
    ctxs = [hr.Context(m, opts=hr.default_opts(rank=r, world=world, device=(devices[r] if devices else 0))) for r in range(world)]
    hr.Context.connect_local(ctxs)
    results = [[] for _ in range(world)]

    def work(r):
        for k in steps:
            results[r].append(ctxs[r].step(k))
    th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    [t.start() for t in th]
    [t.join(timeout=300) for t in th]
    assert not any(t.is_alive() for t in th), "a rank hung"
    return ctxs, results


@pytest.mark.parametrize("name,world", [("sbox", 2), ("model3s", 3), ("model1s", 4)])
def test_sharded_ranks_equal_single_gpu(hr, models_dir, name, world):
    m = hr.Model(str(models_dir / f"{name}.tri"))
    one = hr.Context(m)
    done1, conv1 = one.step(-1)
    acc1, dB1, par1 = one.download()
    ctxs, results = run_ranks(hr, m, world, [-1])
    owned_total = 0
    for r, c in enumerate(ctxs):
        assert results[r] == [(done1, conv1)]
        assert c.shooter_log() == one.shooter_log()
        acc, dB, par = c.download()
        own = c.owned_slots()
        owned_total += len(own)
        assert np.array_equal(acc[own].view(np.uint32), acc1[own].view(np.uint32))
        assert np.array_equal(dB[own].view(np.uint32), dB1[own].view(np.uint32))
    assert owned_total == int((par1 >= 0).sum())


def test_sharded_ranks_with_wide_candidate_sets(hr, models_dir):
    """Three in-process ranks, 128 candidates per round: merge of 3 x 128 mailed candidates, records of 128 rows."""
    m = hr.Model(str(models_dir / "model3s.tri"))
    one = hr.Context(m, per_shot_kernels=1)
    done1, conv1 = one.step(-1)
    acc1, dB1, _ = one.download()
    ctxs = [hr.Context(m, opts=hr.default_opts(rank=r, world=3, ff_rows=128)) for r in range(3)]
    hr.Context.connect_local(ctxs)
    res = [None] * 3

    def work(r):
        res[r] = ctxs[r].step(-1)
    th = [threading.Thread(target=work, args=(r,)) for r in range(3)]
    [t.start() for t in th]
    [t.join(timeout=300) for t in th]
    assert not any(t.is_alive() for t in th), "a rank hung"
    for r, c in enumerate(ctxs):
        assert res[r] == (done1, conv1) and c.shooter_log() == one.shooter_log()
        own = c.owned_slots()
        acc, dB, _ = c.download()
        assert np.array_equal(acc[own].view(np.uint32), acc1[own].view(np.uint32))
        assert np.array_equal(dB[own].view(np.uint32), dB1[own].view(np.uint32))


@pytest.mark.parametrize("name,P,world", [("sbox", 4, 2), ("model3s", 10, 3)])
def test_sharded_ranks_with_parallel_sources(hr, models_dir, name, P, world):
    """-parallel_src across ranks: every rank mails its exact local top P and its part of |C|; the merged schedule
    and the results equal one GPU's generic multi-source kernel (itself pinned to the reference goldens)."""
    m = hr.Model(str(models_dir / f"{name}.tri"))
    one = hr.Context(m, parallel_src=P, per_shot_kernels=1)
    done1, conv1 = one.step(-1)
    acc1, dB1, _ = one.download()
    ctxs = [hr.Context(m, opts=hr.default_opts(rank=r, world=world, parallel_src=P)) for r in range(world)]
    hr.Context.connect_local(ctxs)
    res = [None] * world

    def work(r):
        res[r] = ctxs[r].step(-1)
    th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    [t.start() for t in th]
    [t.join(timeout=300) for t in th]
    assert not any(t.is_alive() for t in th), "a rank hung"
    for r, c in enumerate(ctxs):
        assert res[r] == (done1, conv1) and c.shooter_log() == one.shooter_log()
        own = c.owned_slots()
        acc, dB, _ = c.download()
        assert np.array_equal(acc[own].view(np.uint32), acc1[own].view(np.uint32))
        assert np.array_equal(dB[own].view(np.uint32), dB1[own].view(np.uint32))


def test_sharded_steps_in_pieces_and_rewind(hr, models_dir):
    m = hr.Model(str(models_dir / "model3s.tri"))
    one = hr.Context(m, per_shot_kernels=1)
    one.step(90)
    ctxs, results = run_ranks(hr, m, 2, [1, 40, 49])
    assert results[0] == results[1] == [(1, False), (40, False), (49, False)]
    acc1, dB1, _ = one.download()
    for c in ctxs:
        own = c.owned_slots()
        acc, dB, _ = c.download()
        assert np.array_equal(acc[own].view(np.uint32), acc1[own].view(np.uint32))
        assert c.shooter_log() == one.shooter_log()
    # rewind on all ranks alike, run again: same answer (mailbox epochs do not collide)
    for c in ctxs:
        c.rewind()
    th = [threading.Thread(target=c.step, args=(90,)) for c in ctxs]
    [t.start() for t in th]
    [t.join(timeout=300) for t in th]
    for c in ctxs:
        own = c.owned_slots()
        acc, dB, _ = c.download()
        assert np.array_equal(acc[own].view(np.uint32), acc1[own].view(np.uint32))


def test_unconnected_multi_rank_context_refuses_to_step(hr, models_dir):
    m = hr.Model(str(models_dir / "model2s.tri"))
    c = hr.Context(m, opts=hr.default_opts(rank=0, world=2))
    with pytest.raises(hr.RadkError) as e:
        c.step(1)
    assert "peer_connect" in str(e.value)


def test_real_multi_gpu_under_torchrun():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = min(n, 4)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29631", os.path.join(ROOT, "tests", "multi_gpu_check.py")],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=900)
    assert r.returncode == 0 and b"MULTI_GPU_CHECK OK" in r.stdout, r.stdout.decode(errors="replace")[-3000:]
