"""Times the packed kernels against the bf16 ones (CUDA events, L2 flushed between runs). Usage: python scripts/bench_packed.py [N [ROWS]]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lds_gnn_b200 import kernels as K

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
rows = int(sys.argv[2]) if len(sys.argv) > 2 else n
row0 = 0 if rows == n else (n // 2 // 128) * 128
dev = torch.device("cuda")
ld = K.padded_ld(n)
torch.manual_seed(0)
theta = torch.rand((rows, ld), device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return min(ts), sorted(ts)[len(ts) // 2]


bits = torch.zeros(K.packed_adj_bytes(n, rows), dtype=torch.uint8, device=dev)
res = {}
res["k1_packed"] = timeit(lambda: K.k1_sample_packed(theta, n, 1, 2, row0=row0, rows=rows, bits=bits))
if rows * ld * 2 < 20e9:
    res["k1_bf16"] = timeit(lambda: K.k1_sample_normalize(theta, n, 1, 2, row0=row0, rows=rows))
    adj = K.k1_sample_normalize(theta, n, 1, 2, row0=row0, rows=rows)[0]
else:
    adj = None
for w in (64, 7):
    p = torch.randn(n, w, device=dev)
    z = torch.empty(rows, w, device=dev)
    res[f"k2_packed_w{w}"] = timeit(lambda: K.k2_propagate_packed(bits, n, rows, p, out=z))
    if adj is not None:
        res[f"k2_bf16_w{w}"] = timeit(lambda: K.k2_propagate(adj, n, p, out=z))
nn = float(rows) * n
for k, (mn, med) in res.items():
    extra = ""
    if k.startswith("k2"):
        w = int(k.split("w")[-1])
        extra = f"  {2 * nn * w / mn / 1e6:.1f} TFLOP/s useful"
    print(f"{k:16s} min {mn:9.1f} us  median {med:9.1f} us{extra}")
