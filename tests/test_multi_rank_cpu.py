"""CPU tier of the N > 1 path: several processes (gloo) each hold one receiver shard and run the chained shot loop
of fast_kernels.cuh - per round three exchanges, exactly the mail the GPU ranks send each other: local top
candidates, the candidates' records from their owners, the first violation of the schedule - and must reproduce
the single-rank oracle bit for bit.  Runs the device-core code (rt_core.cuh, shot_core.cuh) via tests/hostemu."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import GOLDEN_DIR, ROOT

ROWS = 8      # cache rows = candidates per round (small, so that rounds end on violations too)


def _worker(rank, world, port, model_gz, shots, out_dir, psrc=1):
    import gzip
    import tempfile

    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import hw_radiosity_b200 as hr
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    emu = C.CDLL(os.path.join(ROOT, "tests", "hostemu", "libhostemu.so"))
    emu.emu_shard_open.restype = C.c_void_p
    emu.emu_shard_open.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float]
    for f in ("emu_shard_top", "emu_shard_begin", "emu_shard_cands"):
        getattr(emu, f).argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    emu.emu_shard_chain2.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong]
    emu.emu_shard_set_parallel_src.argtypes = [C.c_void_p, C.c_int]
    emu.emu_shard_count.restype = C.c_longlong
    emu.emu_shard_count.argtypes = [C.c_void_p]
    emu.emu_shard_apply.argtypes = [C.c_void_p]
    emu.emu_shard_commit.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    emu.emu_shard_state.argtypes = [C.c_void_p] * 4
    emu.emu_shard_sizes.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    nrec, ncand, rows = emu.emu_shooter_rec_bytes(), emu.emu_cand_rec_bytes(), ROWS
    with tempfile.TemporaryDirectory() as tmp:
        tri = os.path.join(tmp, "m.tri")
        with open(tri, "wb") as f:
            f.write(gzip.open(model_gz).read())
        m = hr.Model(tri)
    h = emu.emu_shard_open(C.addressof(m.scene), rank, world, rows, 1200.0)
    emu.emu_shard_set_parallel_src(h, psrc)          # -parallel_src: co-shooters per iteration

    def all_gather(arr):
        mine = torch.from_numpy(arr.view(np.uint8).reshape(-1).copy())
        parts = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        return [p.numpy() for p in parts]

    log, rounds = [], 0
    while len(log) < shots:
        # exchange 1: every rank's local top candidates, key + complete shooter record (predict mail)
        keys, recs = np.zeros(rows, np.uint64), np.zeros(rows * nrec, np.uint8)
        emu.emu_shard_top(h, keys.ctypes.data, recs.ctypes.data)
        gk = np.concatenate([p.view(np.uint64) for p in all_gather(keys)])
        # ... and with -parallel_src every rank's part of |C| (the one extra all-reduced count of SURVEY.md 8e)
        count = sum(int(p.view(np.int64)[0]) for p in all_gather(np.array([emu.emu_shard_count(h)], np.int64)))
        gr = np.concatenate(all_gather(recs)).reshape(world * rows, nrec)
        order = np.argsort(gk, kind="stable")[::-1][:rows]
        top, toprec = np.ascontiguousarray(gk[order]), np.ascontiguousarray(gr[order])
        if top[0] == 0:
            break
        emu.emu_shard_begin(h, top.ctypes.data, toprec.ctypes.data)
        # exchange 2: the candidates' records from their owners
        cands, owned = np.zeros(rows * ncand, np.uint8), np.zeros(rows, np.int32)
        emu.emu_shard_cands(h, cands.ctypes.data, owned.ctypes.data)
        gc, go = all_gather(cands), all_gather(owned)
        merged = cands.reshape(rows, ncand).copy()
        for r in range(world):
            o = go[r].view(np.int32).astype(bool)
            merged[o] = gc[r].reshape(rows, ncand)[o]
        merged = np.ascontiguousarray(merged)
        emu.emu_shard_chain2(h, merged.ctypes.data, shots - len(log), count)
        # exchange 3: the first violation of the schedule
        viol = np.array([emu.emu_shard_apply(h)], np.int32)
        v = min(int(p.view(np.int32)[0]) for p in all_gather(viol))
        slots = np.zeros(64, np.int32)
        n = emu.emu_shard_commit(h, v, slots.ctypes.data)
        assert n > 0
        log += list(slots[:n])
        rounds += 1
    R, U = C.c_int(), C.c_int()
    emu.emu_shard_sizes(h, C.byref(R), C.byref(U))
    slots = np.zeros(R.value, np.int32)
    acc = np.zeros((R.value, 9), np.float32)
    dB = np.zeros((R.value, 3), np.float32)
    emu.emu_shard_state(h, slots.ctypes.data, acc.ctypes.data, dB.ctypes.data)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), slots=slots, acc=acc, dB=dB, log=np.array(log, np.int64), rounds=rounds)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name,world,shots,psrc", [("sbox", 2, 45, 1), ("model3s", 3, 60, 1), ("sbox", 2, 30, 3)])
def test_two_rank_gloo_exchange_matches_oracle(hr, po, models_dir, tmp_path, name, world, shots, psrc):
    from conftest import oracle_scene
    test_hostemu = __import__("test_hostemu")      # builds libhostemu.so if stale
    so = os.path.join(ROOT, "tests", "hostemu", "libhostemu.so")
    if not os.path.exists(so):
        pytest.skip("hostemu not built")
    port = 29500 + (os.getpid() % 400)
    mp.spawn(_worker, args=(world, port, os.path.join(GOLDEN_DIR, "models", f"{name}.tri.gz"), shots, str(tmp_path), psrc),
             nprocs=world, join=True)
    m = hr.Model(str(models_dir / f"{name}.tri"))
    sc = oracle_scene(po, m, parallel_src=psrc)
    seq = sc.run(shots)
    seen = 0
    for r in range(world):
        d = np.load(tmp_path / f"rank{r}.npz")
        assert list(d["log"]) == [s[0] for s in seq]
        assert int(d["rounds"]) < shots or psrc > 1  # shots really were batched (co-shooters: |C| is small here, one iteration per round)
        slots = d["slots"]
        seen += len(slots)
        assert np.array_equal(d["acc"].reshape(-1, 3, 3).view(np.uint32), sc.slots["accB"][slots].view(np.uint32))
        assert np.array_equal(d["dB"].view(np.uint32), sc.slots["deltaB"][slots].view(np.uint32))
    assert seen == int((sc.slots["parent"] >= 0).sum())          # shards are disjoint and cover every receiver
