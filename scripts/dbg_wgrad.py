import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from repsurf_b200 import tc
dev = torch.device("cuda")
def run(R, M, Nn):
    g = torch.Generator().manual_seed(R + M)
    G = torch.randn(R, M, generator=g).to(dev); X = torch.randn(R, Nn, generator=g).to(dev)
    dW = torch.zeros(M, Nn, device=dev)
    tc.gemm_wgrad(R, tc.opnd(tc.OPND_RAW, G, M), tc.opnd(tc.OPND_RAW, X, Nn), dW)
    ref = G.double().t() @ X.double()
    torch.cuda.synchronize()
    err = ((dW.double() - ref).abs().max() / ref.abs().max()).item()
    print(f"R={R} M={M} N={Nn} sbo={os.environ.get('RSB_WG_SBO')} swap={os.environ.get('RSB_WG_SWAP')} rel_err={err:.3e} |dW|max={dW.abs().max().item():.3e} ref max={ref.abs().max().item():.3e}", flush=True)
if len(sys.argv) > 1:
    for cfg in [(32, 128, 32), (64, 16, 16), (1000, 64, 19), (5000, 300, 140), (100000, 64, 32)]:
        try: run(*cfg)
        except Exception as e: print("ERR", cfg, e, flush=True)
else:
    for sbo in ("528", "512"):
        for swap in ("0", "1"):
            env = dict(os.environ, RSB_WG_SBO=sbo, RSB_WG_SWAP=swap)
            subprocess.run([sys.executable, __file__, "child"], env=env, timeout=120)
