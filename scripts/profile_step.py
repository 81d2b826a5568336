"""torch.profiler kernel table for one fwd+bwd step (diagnostic; not a bench)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn as nn
from torch.profiler import profile, ProfilerActivity
import bench
wl = sys.argv[1] if len(sys.argv) > 1 else "seg"
from repsurf_b200.models import RepSurfCls, RepSurfSeg, SmoothClsLoss
from repsurf_b200.seg import pointops as PS
dev = torch.device("cuda")
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
W = bench.WORKLOADS[wl]
host = bench.make_inputs(wl, W["clouds"], W["n"], 100, True)
inp = [t.to(dev) for t in host]
model = (RepSurfSeg() if wl == "seg" else RepSurfCls()).to(dev).train()
from repsurf_b200.seg.loss import CrossEntropyLoss
crit = CrossEntropyLoss() if wl == "seg" else SmoothClsLoss()
def step():
    model.zero_grad(set_to_none=True)
    if wl == "seg":
        PS.register_offsets(inp[2], host[2].tolist())
        loss = crit(model([inp[0], inp[1], inp[2]]), inp[3])
    else:
        loss = crit(model(inp[0]), inp[1])
    loss.backward()
for _ in range(2): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
print("=== by self CPU time ===")
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=40, max_name_column_width=60))
import time
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); step(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"host launch time {1e3*(t1-t0):.1f} ms, step wall {1e3*(t2-t0):.1f} ms")
