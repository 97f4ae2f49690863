"""CPU tier of the N > 1 path: two processes (gloo) each hold one receiver shard, exchange their best
candidate (key + shooter record) every shot like the GPU ranks do through their mailboxes, and must
reproduce the single-rank oracle bit for bit.  Runs the same device-core code via tests/hostemu."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import GOLDEN_DIR, ROOT


def _worker(rank, world, port, model_gz, shots, out_dir):
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
    emu.emu_shard_open.argtypes = [C.c_void_p, C.c_int, C.c_int]
    emu.emu_shard_best.argtypes = [C.c_void_p, C.c_float, C.POINTER(C.c_ulonglong), C.c_void_p]
    emu.emu_shard_apply.argtypes = [C.c_void_p, C.c_void_p]
    emu.emu_shard_state.argtypes = [C.c_void_p] * 4
    emu.emu_shard_sizes.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    nrec = emu.emu_shooter_rec_bytes()
    with tempfile.TemporaryDirectory() as tmp:
        tri = os.path.join(tmp, "m.tri")
        with open(tri, "wb") as f:
            f.write(gzip.open(model_gz).read())
        m = hr.Model(tri)
    h = emu.emu_shard_open(C.addressof(m.scene), rank, world)
    log = []
    for _ in range(shots):
        key = C.c_ulonglong()
        rec = C.create_string_buffer(nrec)
        emu.emu_shard_best(h, 1200.0, C.byref(key), rec)
        mine = torch.zeros(8 + nrec, dtype=torch.uint8)
        mine[:8] = torch.frombuffer(bytearray(int(key.value).to_bytes(8, "little")), dtype=torch.uint8)
        mine[8:] = torch.frombuffer(bytearray(rec.raw), dtype=torch.uint8)
        gathered = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)                      # the per-shot exchange
        keys = [int.from_bytes(bytes(g[:8].tolist()), "little") for g in gathered]
        best = max(range(world), key=lambda r: keys[r])
        if keys[best] == 0:
            break
        winner = C.create_string_buffer(bytes(gathered[best][8:].tolist()), nrec)
        emu.emu_shard_apply(h, winner)
        log.append(keys[best] & 0xffffffff)
    R, U = C.c_int(), C.c_int()
    emu.emu_shard_sizes(h, C.byref(R), C.byref(U))
    slots = np.zeros(R.value, np.int32)
    acc = np.zeros((R.value, 9), np.float32)
    dB = np.zeros((R.value, 3), np.float32)
    emu.emu_shard_state(h, slots.ctypes.data, acc.ctypes.data, dB.ctypes.data)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), slots=slots, acc=acc, dB=dB, log=np.array(log, np.int64))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name,world,shots", [("sbox", 2, 25), ("model3s", 3, 15)])
def test_two_rank_gloo_exchange_matches_oracle(hr, po, models_dir, tmp_path, name, world, shots):
    from conftest import oracle_scene
    test_hostemu = __import__("test_hostemu")      # builds libhostemu.so if stale
    so = os.path.join(ROOT, "tests", "hostemu", "libhostemu.so")
    if not os.path.exists(so):
        pytest.skip("hostemu not built")
    port = 29500 + (os.getpid() % 400)
    mp.spawn(_worker, args=(world, port, os.path.join(GOLDEN_DIR, "models", f"{name}.tri.gz"), shots, str(tmp_path)),
             nprocs=world, join=True)
    m = hr.Model(str(models_dir / f"{name}.tri"))
    sc = oracle_scene(po, m)
    seq = sc.run(shots)
    seen = 0
    for r in range(world):
        d = np.load(tmp_path / f"rank{r}.npz")
        assert list(d["log"]) == [s[0] for s in seq]
        slots = d["slots"]
        seen += len(slots)
        assert np.array_equal(d["acc"].reshape(-1, 3, 3).view(np.uint32), sc.slots["accB"][slots].view(np.uint32))
        assert np.array_equal(d["dB"].view(np.uint32), sc.slots["deltaB"][slots].view(np.uint32))
    assert seen == int((sc.slots["parent"] >= 0).sum())          # shards are disjoint and cover every receiver
