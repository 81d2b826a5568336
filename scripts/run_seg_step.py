"""Two fwd+bwd steps of the S3DIS workload (for targeted ncu captures; not a bench)."""
import os, sys
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
model = RepSurfSeg().to(dev).train(); crit = nn.CrossEntropyLoss()
for _ in range(int(os.environ.get("ITERS", 2))):
    model.zero_grad(set_to_none=True)
    crit(model([inp[0], inp[1], inp[2]]), inp[3]).backward()
torch.cuda.synchronize()
print("ok")
