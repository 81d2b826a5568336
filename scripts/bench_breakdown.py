"""Where does a bench.py step spend its time?  (diagnostic: synchronises between phases)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn as nn
import bench
from repsurf_b200.models import RepSurfSeg
from repsurf_b200.seg import pointops as PS
dev = torch.device("cuda")
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
W = bench.WORKLOADS["seg"]
host = bench.make_inputs("seg", W["clouds"], W["n"], 100, True)
inp = [t.to(dev) for t in host]
PS.register_offsets(inp[2], host[2].tolist())
model = RepSurfSeg().to(dev).train()
crit = nn.CrossEntropyLoss()
params = list(model.parameters())
mode = sys.argv[1] if len(sys.argv) > 1 else "flat"
if mode == "flat":
    flat = torch.zeros(sum(p.numel() for p in params), device=dev)
    o = 0
    for p in params:
        p.grad = flat[o:o + p.numel()].view_as(p); o += p.numel()
opt = torch.optim.SGD(params, lr=1e-3, momentum=0.9, weight_decay=1e-4)
def sync(): torch.cuda.synchronize(); return time.perf_counter()
for it in range(6):
    t0 = sync()
    if mode == "flat": flat.zero_()
    else: opt.zero_grad(set_to_none=True)
    out = model([inp[0], inp[1], inp[2]]); loss = crit(out, inp[3])
    t1 = sync()
    loss.backward()
    t2 = sync()
    opt.step()
    t3 = sync()
    print(f"{mode} it{it}: fwd {1e3*(t1-t0):.1f}  bwd {1e3*(t2-t1):.1f}  opt {1e3*(t3-t2):.1f} ms", flush=True)
