"""Per-call CUDA-event timings of the tensor-core GEMM entries over one training step (diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn as nn
import bench
from repsurf_b200 import _native
from repsurf_b200.models import RepSurfSeg
from repsurf_b200.seg import pointops as PS
dev = torch.device("cuda")
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
W = bench.WORKLOADS["seg"]
host = bench.make_inputs("seg", W["clouds"], W["n"], 100, True)
inp = [t.to(dev) for t in host]
PS.register_offsets(inp[2], host[2].tolist())
model = RepSurfSeg().to(dev).train(); crit = nn.CrossEntropyLoss()
def step():
    model.zero_grad(set_to_none=True)
    crit(model([inp[0], inp[1], inp[2]]), inp[3]).backward()
for _ in range(3): step()
with bench.EntryTimer(_native, {"rsb_gemm_rows", "rsb_gemm_wgrad", "rsb_bn_relu_backward", "rsb_bn_apply", "rsb_pool_forward", "rsb_pool_backward_stats", "rsb_linear_tc_prep_weight", "rsb_grouping_packed_forward", "rsb_grouping_packed_backward", "rsb_interpolation_packed_forward", "rsb_interpolation_packed_backward"}) as et:
    step()
torch.cuda.synchronize()
tot = {}
for name, a, e0, e1 in et.ev:
    t = e0.elapsed_time(e1)
    tot[name] = tot.get(name, 0) + t
    if name == "rsb_gemm_rows":
        print(f"rows  R={a[0]:8d} N={a[1]:4d} K={a[2].K:4d} kind={a[2].kind}  {t:7.3f} ms  {(a[0]*(a[2].K+a[1])*4)/t/1e6:8.1f} GB/s")
    elif name == "rsb_gemm_wgrad":
        print(f"wgrad R={a[0]:8d} M={a[1].K:4d} N={a[2].K:4d} kinds={a[1].kind}/{a[2].kind}  {t:7.3f} ms  {(a[0]*(a[1].K+a[2].K)*4)/t/1e6:8.1f} GB/s")
print({k: round(v, 3) for k, v in tot.items()})
