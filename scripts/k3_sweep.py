"""K3+K4 device time vs factor width d at a fixed N (tuning aid): lds_k3k4_theta_update_tc includes the pack kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lds_gnn_b200 import kernels as K
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
ds = [int(a) for a in sys.argv[2:]] or [12, 23, 40, 60, 71, 100]
dev = torch.device("cuda")
ld = K.padded_ld(n)
theta = torch.rand((n, ld), device=dev)
for d in ds:
    fa = torch.randn((n, d), device=dev) * 0.01
    fb = torch.randn((n, d), device=dev) * 0.01
    cv = torch.randn((n,), device=dev) * 0.001
    for _ in range(2):
        K.k3k4_theta_update_tc_(theta, n, fa, fb, cv, 0.1)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    ev[0].record()
    for i in range(5):
        K.k3k4_theta_update_tc_(theta, n, fa, fb, cv, 0.1)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(5)]
    best = min(ts)
    print(f"n={n} d={d:4d}  best {best:8.1f} us  median {sorted(ts)[2]:8.1f} us   {8.0 * n * n / best / 1e3:7.1f} GB/s algorithmic")
