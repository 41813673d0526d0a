"""Runs each hot kernel once or twice at the bench size (m = 32^4, r = 432, fp32) — the target of the ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from online_gp_b200 import ops

dev = "cuda:0"
m, r = 32 ** 4, int(os.environ.get("PROBE_R", "432"))
sizes = [32, 32, 32, 32]
torch.manual_seed(0)
L = torch.randn(m, r, device=dev) / m ** 0.5
cols = (torch.rand(4, 32, device=dev) * 0.1 + torch.exp(-0.1 * torch.arange(32, device=dev) ** 2)).requires_grad_(True)
for it in range(int(os.environ.get("PROBE_ITERS", "2"))):
    KL = ops.kron_toeplitz_matmul(cols, sizes, L)
    G = ops.gram(L, KL)
    (G.sum()).backward()
    with torch.no_grad():
        p = torch.randn(r, 1, device=dev) / r ** 0.5
        ops.panel_lowrank_update_(L, p, 0.1 * p.t())
        v = torch.randn(r, 1, device=dev)
        w = ops.q_matvec(L, KL.detach(), v)
        KLn = ops.kron_toeplitz_matmul(cols.detach(), sizes, L)
    cols.grad = None
torch.cuda.synchronize()
print("probe done", float(w.sum()))
