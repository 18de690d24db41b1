"""world_size-2 (and 3) gloo test of the one collective on the path: every rank renders its tile shard, the
shards are all-gathered and de-tiled into the row-major frame.  On CPU the shard content comes
from the oracle (rendering the full frame and cutting the rank's tiles out of it), so this covers the host-side
partitioning / gather / assembly logic that bench.py uses with NCCL."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, w, h, ret):
    sys.path.insert(0, str(REPO / "rendering-fw_b200" / "python"))
    import torch
    import torch.distributed as dist

    sys.path.insert(0, str(REPO))
    import rfwb200 as R
    from oracle.oracle_lib import load_oracle
    import scenes as S

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = S.cornell_box(unit_scale=True)
    ctx = R.RenderContext(load_oracle())
    S.upload(ctx, sc, w, h)
    ctx.render_frame(sc.camera(w, h), R.RESET)
    full = ctx.read_framebuffer().reshape(-1, 4)
    # this rank's tile-major shard, padded to the common stride
    m = R.shard_pixel_map(w, h, rank, world)
    stride = R.shard_stride(w, h, world)
    shard = np.zeros((stride, 4), np.float32)
    shard[: len(m)][m >= 0] = full[m[m >= 0]]
    local = torch.from_numpy(shard.reshape(-1))
    gathered = torch.empty(world * stride * 4, dtype=torch.float32)
    dist.all_gather_into_tensor(gathered, local)
    img = R.assemble_shards_host(list(gathered.numpy().reshape(world, stride, 4)), w, h)
    ok = np.array_equal(img.reshape(-1, 4), full)
    t = torch.tensor([1 if ok else 0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        ret.put(int(t.item()))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,w,h", [(2, 96, 64), (3, 100, 50)])
def test_gather_and_assemble_over_gloo(built, world, w, h):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, w, h, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=5) == 1
