import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lds_gnn_b200 import kernels as K, _lib
from oracle.make_golden import make_inputs

n, f, h, c = 64, 40, 16, 7
inp = make_inputs(seed=n, n=n, f=f, h=h, c=c, theta_kind="uniform", p=0.0)
dev = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
eng = K.OuterStep(n, dev(inp["x"]), dev(inp["y"]), dev(inp["mask"]), hidden=h, classes=c)
eng.set_weights(dev(inp["w0"]), dev(inp["b0"]), dev(inp["w1"]), dev(inp["b1"]))
full0 = K.theta_triu_to_full(dev(inp["theta_triu"]))
keep = full0.clone()
print("ptrs full0 %x keep %x ws %x..%x" % (full0.data_ptr(), keep.data_ptr(), eng.ws.data_ptr(), eng.ws.data_ptr()+eng.ws_bytes))
def chk(tag):
    torch.cuda.synchronize()
    print(f"  [{tag}] full0 intact: {torch.equal(full0, keep)}")
a = full0.clone(); print("a %x" % a.data_ptr())
eng.run(a, lr=0.5, seed=1, step=0, dropout_p=0.0, update=False); chk("no update")
fa, fb, cv = eng.buffer("fa").clone(), eng.buffer("fb").clone(), eng.buffer("cvec").clone()
b = full0.clone(); print("b %x" % b.data_ptr())
K.k3k4_theta_update_(b, n, fa, fb, cv, 0.5, d=h+c); chk("simt standalone")
c2 = full0.clone(); print("c2 %x" % c2.data_ptr())
K.k3k4_theta_update_tc_(c2, n, fa, fb, cv, 0.5, d=h+c); chk("tc standalone")
print("standalone tc vs simt", (b-c2).abs().max().item())
d2 = full0.clone(); print("d2 %x" % d2.data_ptr())
eng.run(d2, lr=0.5, seed=1, step=0, dropout_p=0.0, update=True, k3_flags=_lib.K3_SIMT); chk("fused simt")
print("fused simt vs standalone simt", (b-d2).abs().max().item())
e2 = full0.clone(); print("e2 %x" % e2.data_ptr())
eng.run(e2, lr=0.5, seed=1, step=0, dropout_p=0.0, update=True); chk("fused tc")
print("fused tc vs standalone simt", (b-e2).abs().max().item(), " e2 unchanged?", torch.equal(e2, keep))
# compare packed operands: fused pack (in ws) vs standalone pack
lib = _lib.load()
import ctypes
print("pm/qm region after fused: nonzero count", int((eng.ws.view(torch.uint8) != 0).sum().item()))
