"""Data-parallel equivalence on the GPU, driver-run: two ranks (two processes sharing cuda:0, gloo transport so that no
NCCL kernel has to co-run with the other rank's kernels on one device) each run the real CUDA training step on their
own B images through b200unet.ddp.DataParallel; the averaged, bucket-all-reduced gradients must equal the gradients of
ONE process on the concatenated 2B images (the identity batch-sharded data parallelism relies on).
With >= 2 GPUs visible the same test also runs over NCCL, one rank per GPU."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu
ARGS = (1, 2, 3, 6, False, False, "upconv")   # paper graph at 64..256 channels: the tcgen05 kernels, no BatchNorm
B, H = 2, 92


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _data():
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2 * B, 1, H, H, generator=g)
    y = torch.randint(0, 2, (2 * B, 52, 52), generator=g)
    return x, y


def _worker(rank, world, port, backend, out_dir):
    import torch.distributed as dist
    import b200unet
    from b200unet.ddp import DataParallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", rank if backend == "nccl" else 0)
    torch.cuda.set_device(dev)
    dist.init_process_group(backend, rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                    # different init per rank: the wrapper must broadcast rank 0's weights
    model = b200unet.UNet(*ARGS).to(dev).train()
    net = DataParallel(model, bucket_bytes=256 << 10)  # small buckets: several all-reduces are in flight during backward
    assert len(net.bucketer.buckets) >= 3
    x, y = _data()
    xs, ys = x[rank * B:(rank + 1) * B].to(dev), y[rank * B:(rank + 1) * B].to(dev)
    losses = []
    for _ in range(2):                               # the arena is re-created every backward
        model.zero_grad(set_to_none=True)
        loss = net.loss(xs, ys)
        loss.backward()
        losses.append(float(loss))
    torch.cuda.synchronize()
    torch.save({"grads": {k: p.grad.detach().cpu() for k, p in model.named_parameters()},
                "weights": {k: p.detach().cpu() for k, p in model.named_parameters()}, "loss": losses[-1]},
               os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def _run(backend, tmp_path):
    import torch.multiprocessing as mp
    import b200unet
    from gpu_util import rel_l2
    port = _free_port()
    mp.spawn(_worker, args=(2, port, backend, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(os.path.join(tmp_path, "rank0.pt"))
    r1 = torch.load(os.path.join(tmp_path, "rank1.pt"))
    for k in r0["weights"]:
        assert torch.equal(r0["weights"][k], r1["weights"][k]), f"replicas differ in {k}"
    for k in r0["grads"]:                            # the all-reduced gradient is the same tensor on both ranks
        assert torch.equal(r0["grads"][k], r1["grads"][k]), k
    # single process, the same weights, all 2B images at once
    model = b200unet.UNet(*ARGS).cuda().train()
    model.load_state_dict(r0["weights"])
    x, y = _data()
    loss = model.loss(x.cuda(), y.cuda())
    loss.backward()
    keys = list(r0["grads"])
    got = torch.cat([r0["grads"][k].flatten() for k in keys])
    want = torch.cat([dict(model.named_parameters())[k].grad.cpu().flatten() for k in keys])
    e = rel_l2(got, want)
    print(f"[ddp {backend}] 2-rank averaged gradients vs one process on 2B images: rel-L2 {e:.3e}; "
          f"loss {0.5 * (r0['loss'] + r1['loss']):.6f} vs {float(loss):.6f}")
    # different tile plans (n_img differs) and a different summation order: fp32 accumulation noise only
    assert e < 2e-3
    assert abs(0.5 * (r0["loss"] + r1["loss"]) - float(loss)) < 1e-5


def test_two_ranks_one_gpu_gloo_equal_single_process(tmp_path):
    _run("gloo", tmp_path)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_ranks_nccl_equal_single_process(tmp_path):
    _run("nccl", tmp_path)


def _graph_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    import b200unet
    from b200unet.ddp import DataParallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x, y = _data()
    xs, ys = x[rank * B:(rank + 1) * B].to(dev), y[rank * B:(rank + 1) * B].to(dev)
    out = {}
    for mode in ("eager", "graph"):
        torch.manual_seed(7)
        model = b200unet.UNet(*ARGS).to(dev).train()
        net = DataParallel(model, bucket_bytes=256 << 10)
        opt = b200unet.FusedAdam(model.parameters(), lr=1e-3)
        if mode == "graph":
            step = b200unet.GraphedTrainStep(net, opt, xs, ys, warmup=1)   # 1 warm-up step + 1 capture-time reduce
            losses = [float(step(xs, ys)) for _ in range(3)]
        else:
            losses = []
            for _ in range(4):                                             # the same number of optimizer updates
                loss = net.loss(xs, ys)
                opt.zero_grad(set_to_none=True)
                loss.backward()
                opt.step()
                losses.append(float(loss))
        torch.cuda.synchronize()
        out[mode] = {"w": {k: p.detach().cpu() for k, p in model.named_parameters()}, "loss": losses[-1]}
    torch.save(out, os.path.join(out_dir, f"graph_rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_graphed_step_with_data_parallel_two_ranks(tmp_path):
    """GraphedTrainStep over DataParallel = graph(fwd + bwd) -> one all-reduce of the gradient arena -> graph(optimizer):
    after the same number of updates the weights equal the eager bucketed data-parallel run, on both ranks."""
    import torch.multiprocessing as mp
    from gpu_util import rel_l2
    mp.spawn(_graph_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(os.path.join(tmp_path, "graph_rank0.pt"))
    r1 = torch.load(os.path.join(tmp_path, "graph_rank1.pt"))
    for k in r0["graph"]["w"]:
        assert torch.equal(r0["graph"]["w"][k], r1["graph"]["w"][k]), f"replicas differ in {k}"
    keys = list(r0["graph"]["w"])
    got = torch.cat([r0["graph"]["w"][k].flatten() for k in keys])
    want = torch.cat([r0["eager"]["w"][k].flatten() for k in keys])
    assert rel_l2(got, want) < 1e-5
    assert abs(r0["graph"]["loss"] - r0["eager"]["loss"]) < 1e-4
