"""One SurfaceAbstraction shared-MLP block (fused tcgen05 path) at S3DIS sa1 size, fwd+bwd, for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from repsurf_b200 import tc
from tests.test_mlp_gpu import _Block
dev = torch.device("cuda")
G, ns, pos_c, feat_c, mlp = (int(os.environ.get("G", 81920)), 32, 3, 16, [32, 32, 64])
torch.manual_seed(0)
blk = _Block(pos_c, feat_c, mlp, 1).to(dev).train()
X = torch.randn(G * ns, pos_c + feat_c, device=dev, requires_grad=True)
for it in range(int(os.environ.get("ITERS", 2))):
    out = tc.sa_mlp_fused(X, pos_c, blk, ns)
    out.backward(torch.ones_like(out))
torch.cuda.synchronize()
print("ok", out.shape)
