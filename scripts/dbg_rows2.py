"""Sweep of gemm_rows generation 2 against fp64 over ragged N tilings / operand kinds (debugging aid)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from repsurf_b200 import _native as N  # noqa: E402
from repsurf_b200 import tc  # noqa: E402
from scripts.ab_gemm import make_opnd, rel  # noqa: E402

dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
bad = 0
for R in (300, 2000):
    for K in (32, 64, 256):
        for Nn in (160, 272, 288, 320, 512):
            for kind in ("raw", "bn", "aff", "dual"):
                A, Ad = make_opnd(kind, R, K, g)
                W = torch.randn(Nn, K, generator=g, device=dev) / K ** 0.5
                Wp, _, _ = tc.prep_weight(W)
                ref = Ad @ W.double().t()
                errs = []
                for gen in (1, 0):
                    N.lib().rsb_tc_set_generation(gen)
                    Y = torch.full((R, Nn), float("nan"), device=dev)
                    tc.gemm_rows(R, Nn, A, Wp, Y=Y)
                    errs.append(rel(Y, ref))
                N.lib().rsb_tc_set_generation(0)
                flag = "" if errs[1] < 1e-5 else "   <-- BAD"
                bad += errs[1] >= 1e-5
                if flag or (K == 256 and Nn == 272):
                    # which columns / rows are wrong
                    d = (Y.double() - ref).abs() / ref.abs().max()
                    cols = (d.max(0).values > 1e-5).nonzero().flatten().tolist()
                    rows = (d.max(1).values > 1e-5).nonzero().flatten().tolist()
                    extra = f" bad cols {cols[:4]}..{cols[-2:]} ({len(cols)}) rows {rows[:3]}..{rows[-2:]} ({len(rows)})" if cols else ""
                    print(f"R={R} K={K} N={Nn} {kind}: v1 {errs[0]:.1e} v2 {errs[1]:.1e}{flag}{extra}", flush=True)
print("bad cases:", bad)
