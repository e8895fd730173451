"""Data-parallel plumbing on CPU: the gradient bucketer (flat arena in backward order, bucket-complete -> async
all-reduce, average) over a world_size-2 gloo group.  The CUDA model is replaced by a stub that produces gradients
in the same order the hand-scheduled backward does; the arithmetic checked is the collective, not the kernels."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import b200unet
from b200unet.ddp import GradBucketer, backward_order


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    model = b200unet.UNet(1, 2, 3, 2, False, True, "upsample")  # parameter containers only; nothing runs on CPU
    names = backward_order(model)
    shapes = {n: tuple(p.shape) for n, p in model.named_parameters()}
    gb = GradBucketer(names, shapes, bucket_bytes=2048)
    assert len(gb.buckets) > 3 and set(names) == set(shapes)
    for step in range(2):  # the arena is re-created every backward
        gb.begin("cpu")
        grads = {}
        for i, n in enumerate(names):  # production order of the backward
            g = gb.alloc(n, shapes[n], "cpu")
            g.fill_(float(rank + 1) * (i + 1) + step)
            grads[n] = g
            gb.ready(n)
        gb.finish()
        for i, n in enumerate(names):
            want = sum(float(r + 1) * (i + 1) + step for r in range(world)) / world
            assert torch.allclose(grads[n], torch.full(shapes[n], want)), (n, grads[n].flatten()[0].item(), want)
    # deferred mode (GraphedTrainStep over DataParallel): backward only fills the arena, ONE collective reduces it afterwards
    gb.defer = True
    gb.begin("cpu")
    grads = {}
    for i, n in enumerate(names):
        g = gb.alloc(n, shapes[n], "cpu")
        g.fill_(float(rank + 1) * (i + 1))
        grads[n] = g
        gb.ready(n)                      # must NOT start a reduction
    gb.finish()
    arena = gb.deferred_arena
    assert arena is not None and gb.arena is None
    i0 = names.index(names[0])
    assert torch.allclose(grads[names[0]], torch.full(shapes[names[0]], float(rank + 1) * (i0 + 1)))   # still local
    gb.reduce_all(arena)
    for i, n in enumerate(names):
        want = sum(float(r + 1) * (i + 1) for r in range(world)) / world
        assert torch.allclose(grads[n], torch.full(shapes[n], want)), n
    out.put((rank, "ok"))
    dist.destroy_process_group()


def test_bucketed_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(2))
    assert got == [(0, "ok"), (1, "ok")]


def test_backward_order_is_reverse_forward():
    m = b200unet.UNet(1, 2, 3, 2)
    order = backward_order(m)
    assert order[:2] == ["last.weight", "last.bias"]
    assert order[-2:] == ["down_path.0.block.0.weight", "down_path.0.block.0.bias"]
    assert sorted(order) == sorted(n for n, _ in m.named_parameters())
    # buckets are contiguous, 16-byte aligned slices of one arena
    gb = GradBucketer(order, {n: tuple(p.shape) for n, p in m.named_parameters()}, bucket_bytes=4096)
    prev_end = 0
    for s, e, c in gb.buckets:
        assert s == prev_end and e > s and c > 0 and s % 4 == 0
        prev_end = e
    assert prev_end == gb.total
