"""Row N3 measurement: the CVO kernel-Gramian loss at the reference's full size (N = 96 x 128 = 12 288 points per frame,
options.py:109-110; domains xyz 3 ch / hsv_graduv 5 ch / feature 6 ch; pairs (0,0), (1,1), (0,1) as min_dist_mode).

    python experiments/bench_cvo.py [--n 12288] [--steps 20]

Prints one JSON line: ms per loss evaluation (forward + backward to every input) of
  fused        b200unet.cvo.cvo_losses   (one kernel per pair and direction, nothing of size N x N stored)
  matrices     b200unet.cvo.kern_mat per domain and pair + torch product / sum (the reference's structure, our kernels)
  torch_chain  the oracle's pure-PyTorch chain on the same GPU (what the reference would run without its extensions)
plus the pair rate (pairs x domains evaluated per second) and the HBM bytes the materialising forms move."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pytorch-unet_b200"))
from b200unet import cvo  # noqa: E402
from oracle import cvo_oracle as CO  # noqa: E402


def frames(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    h = 96
    w = n // h
    grid = torch.stack(torch.meshgrid(torch.linspace(-1, 1, h), torch.linspace(-1.3, 1.3, w), indexing="ij")).reshape(1, 2, h * w)
    out = []
    for i in range(2):
        depth = 1.5 + 0.3 * torch.rand(1, 1, h * w, generator=g)
        xyz = torch.cat([grid * depth, depth], 1) + 0.01 * i
        out.append({"xyz": xyz.cuda().requires_grad_(True),
                    "img": torch.rand(1, 5, h * w, generator=g).cuda().requires_grad_(True),
                    "feature": (torch.randn(1, 6, h * w, generator=g) * 0.1).cuda().requires_grad_(True)})
    return out


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=96 * 128)
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    items = ["xyz", "img", "feature"]
    coef = {"xyz": 0.2, "img": 0.5, "feature": 0.1}
    f = frames(args.n)
    pairs = [(0, 0), (1, 1), (0, 1)]

    def zero():
        for d in f:
            for t in d.values():
                t.grad = None

    def fused():
        zero()
        L = cvo.cvo_losses(f, items, coef)
        L["func_dist"].backward()
        return L

    def chain(kern):
        zero()
        ip = {}
        for (i, j) in pairs:
            p = None
            for it in items:
                k = kern(f[i][it], f[j][it], coef[it])
                p = k if p is None else p * k
            ip[(i, j)] = p.sum()
        L = ip[(0, 0)] + ip[(1, 1)] - 2 * ip[(0, 1)]
        L.backward()
        return L

    lf = float(fused()["func_dist"])
    lm = float(chain(cvo.kern_mat))
    lt = float(chain(CO.kern_mat))
    t_fused = timed(fused, args.steps)
    t_mat = timed(lambda: chain(cvo.kern_mat), max(3, args.steps // 4))
    t_torch = timed(lambda: chain(CO.kern_mat), max(2, args.steps // 10), warmup=1)
    n = args.n
    pair_domains = 3 * n * n * 3 * 3  # pairs x domains x (forward + two backward directions)
    print(json.dumps({
        "metric": "CVO loss evaluations/s (fwd+bwd, 3 frame pairs x 3 domains, N points per frame)", "n_points": n,
        "fused_ms": t_fused, "matrices_ms": t_mat, "torch_chain_ms": t_torch,
        "fused_over_torch_chain": t_torch / t_fused, "fused_over_matrices": t_mat / t_fused,
        "fused_pair_domain_evals_per_s": pair_domains / (t_fused * 1e-3),
        "matrix_bytes_materialised_by_the_reference_form": 9 * n * n * 4,
        "func_dist": {"fused": lf, "matrices": lm, "torch_chain": lt},
        "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}))


if __name__ == "__main__":
    main()
