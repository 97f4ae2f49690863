"""torchrun entry: N processes, one GPU each, CUDA-IPC mailboxes; compares with a single-GPU run.
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/multi_gpu_check.py [model]"""
import gzip
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hw_radiosity_b200 as hr  # noqa: E402
from hw_radiosity_b200 import dist as rdist  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "church"
    rank, world, local_rank = rdist.init("nccl")
    import torch.distributed as dist
    with tempfile.TemporaryDirectory() as tmp:
        tri = os.path.join(tmp, f"{name}.tri")
        with open(tri, "wb") as f:
            f.write(gzip.open(os.path.join(ROOT, "tests", "golden", "models", f"{name}.tri.gz")).read())
        model = hr.Model(tri)
    ctx = rdist.sharded_context(model, rank, world, local_rank)
    done, conv = ctx.step(-1)
    full = rdist.gather_state(ctx, rank, world)
    log = ctx.shooter_log()
    ok = True
    if rank == 0:
        one = hr.Context(model, device=local_rank)
        d1, c1 = one.step(-1)
        acc1, dB1, par1 = one.download()
        live = par1 >= 0
        ok = (done, conv) == (d1, c1) and log == one.shooter_log() and \
            np.array_equal(full[0][live].view(np.uint32), acc1[live].view(np.uint32)) and \
            np.array_equal(full[1][live].view(np.uint32), dB1[live].view(np.uint32))
        print(f"world={world} model={name} shots={done} converged={conv} device_ms={ctx.device_ms:.3f} "
              f"single_gpu_ms={one.device_ms:.3f} identical={ok}", flush=True)
    flag = [ok]
    dist.broadcast_object_list(flag, src=0)
    dist.barrier()
    if rank == 0 and flag[0]:
        print("MULTI_GPU_CHECK OK", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag[0] else 1)


if __name__ == "__main__":
    main()
