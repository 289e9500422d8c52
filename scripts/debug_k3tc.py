import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lds_gnn_b200 import kernels as K

def run(n, d, lr=0.7, seed=0):
    rng = np.random.default_rng(seed)
    t = n*(n+1)//2
    th = rng.random(t).astype(np.float32)
    fa = (rng.standard_normal((n, d))*0.1).astype(np.float32)
    fb = (rng.standard_normal((n, d))*0.1).astype(np.float32)
    cv = (rng.standard_normal(n)*0.01).astype(np.float32)
    dev = lambda a: torch.as_tensor(a).cuda()
    full = K.theta_triu_to_full(dev(th))
    before = full.clone()
    K.k3k4_theta_update_tc_(full, n, dev(fa), dev(fb), dev(cv), lr)
    torch.cuda.synchronize()
    out = full[:, :n].double().cpu().numpy()
    fa64, fb64, c64 = fa.astype(np.float64), fb.astype(np.float64), cv.astype(np.float64)
    g = fa64@fb64.T + fb64@fa64.T + c64[:,None] + c64[None,:]
    np.fill_diagonal(g, 0)
    thf = before[:, :n].double().cpu().numpy()
    ref = np.clip(thf - lr*g, 0, 1)
    err = np.abs(out-ref)
    # implied gradient
    gi = (thf - out)/lr
    inside = (ref > 0) & (ref < 1) & (out > 0) & (out < 1)
    print(f"n={n} d={d}: max err {err.max():.3e}  max|lr*g| {np.abs(lr*g).max():.3e} sym={np.array_equal(out, out.T)}")
    if err.max() > 1e-5:
        bad = err > 1e-5
        print("  bad fraction", bad.mean(), "rows with bad:", np.unique(np.nonzero(bad)[0])[:20], "cols:", np.unique(np.nonzero(bad)[1])[:40])
        # per 32-col slab / 128 tile stats
        for bj in range((n+127)//128):
            for s in range(4):
                c0 = bj*128+s*32
                if c0 >= n: continue
                print(f"   tile col {bj} slab {s}: bad frac {bad[:, c0:c0+32].mean():.3f}", end=";")
            print()
        i, j = np.argwhere(bad)[0]
        print("  first bad", i, j, "out", out[i,j], "ref", ref[i,j], "theta", thf[i,j], "g_ref", g[i,j], "g_implied", gi[i,j])
        # compare implied g with partial terms
        t1 = fa64@fb64.T; t2 = fb64@fa64.T
        print("  t1", t1[i,j], "t2", t2[i,j], "ci+cj", c64[i]+c64[j])
        m = inside
        for name, cand in (("t1+c", t1 + c64[:,None]+c64[None,:]), ("t2+c", t2 + c64[:,None]+c64[None,:]), ("c only", 0*t1 + c64[:,None]+c64[None,:]), ("2*t1+c", 2*t1+ c64[:,None]+c64[None,:]), ("t1+t2", t1+t2)):
            cc = cand.copy(); np.fill_diagonal(cc, 0)
            print(f"   implied-g vs {name}: {np.abs(gi-cc)[m].max():.3e}")

for n, d in [(64, 4), (64, 22), (128, 22), (96, 22), (200, 22), (257, 71), (1000, 22)]:
    run(n, d)
