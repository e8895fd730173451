import copy, sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-unet_b200")); sys.path.insert(0, ROOT)
from b200unet import FusedAdam
torch.manual_seed(0)
ps = [torch.randn(s, device="cuda").requires_grad_(True) for s in [(64,), (33, 17, 3, 3)]]
a = FusedAdam(ps, lr=1e-3)
for it in range(2):
    for p in ps: p.grad = torch.randn_like(p)
    a.step()
qs = [p.detach().clone().requires_grad_(True) for p in ps]
b = torch.optim.Adam(qs, lr=1e-3)
sd = copy.deepcopy(a.state_dict())
print("saved groups", {k: v for k, v in sd["param_groups"][0].items() if k != "params"})
print("saved steps", [float(s["step"]) for s in sd["state"].values()])
b.load_state_dict(sd)
print("torch groups after load", {k: v for k, v in b.param_groups[0].items() if k != "params"})
for p, q in zip(ps, qs):
    print("exp_avg equal", torch.equal(a.state[p]["exp_avg"], b.state[q]["exp_avg"]), "step", b.state[q]["step"])
for it in range(2):
    for p, q in zip(ps, qs):
        g = torch.randn_like(p); p.grad = g.clone(); q.grad = g.clone()
    a.step(); b.step()
    for p, q in zip(ps, qs):
        print(it, "max |p-q|", float((p - q).abs().max()), "steps", float(a.state[p]["step"]), float(b.state[q]["step"]))
