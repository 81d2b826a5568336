"""A/B of the two GEMM kernel generations (csrc/mlp_tc.cu vs csrc/mlp_tc2.cu) on the S3DIS / ScanObjectNN launch shapes:
result difference against fp64 and CUDA-event time per launch.  Usage: python scripts/ab_gemm.py [quick]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from repsurf_b200 import _native as N  # noqa: E402
from repsurf_b200 import tc  # noqa: E402

dev = torch.device("cuda")
PEAK = 6571.2


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    flush = torch.empty(64 * 1024 * 1024, device=dev)
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


def make_opnd(kind, R, K, g):
    """returns (opnd, fp64 matrix)"""
    if kind == "raw":
        U = torch.randn(R, K, generator=g, device=dev)
        return tc.opnd(tc.OPND_RAW, U, K), U.double()
    if kind == "bn":
        U = torch.randn(R, K, generator=g, device=dev)
        a = torch.rand(K, generator=g, device=dev) + 0.5
        d = torch.randn(K, generator=g, device=dev) * 0.3
        return tc.opnd(tc.OPND_BN_RELU, U, K, a=a, d=d), torch.relu(U.double() * a.double() + d.double())
    if kind == "dual":
        U = torch.randn(R, 2 * K, generator=g, device=dev)
        a = torch.rand(2 * K, generator=g, device=dev) + 0.5
        d = torch.randn(2 * K, generator=g, device=dev) * 0.3
        Ud, ad, dd = U.double(), a.double(), d.double()
        return tc.opnd(tc.OPND_DUAL, U, K, a=a, d=d, ku=K), torch.relu(Ud[:, :K] * ad[:K] + dd[:K] + Ud[:, K:] * ad[K:] + dd[K:])
    if kind in ("aff", "affwrap"):
        ku = K // 2 if kind == "affwrap" else K
        U = torch.randn(R, ku, generator=g, device=dev)
        V = torch.randn(R, K, generator=g, device=dev)
        a = torch.randn(K, generator=g, device=dev)
        b = torch.randn(K, generator=g, device=dev) * 0.1
        d = torch.randn(K, generator=g, device=dev) * 0.1
        Uw = U.double() if ku == K else torch.cat([U.double(), U.double()], 1)
        return tc.opnd(tc.OPND_AFFINE2, U, K, a=a, b=b, d=d, V=V, ku=ku), a.double() * Uw + b.double() * V.double() + d.double()
    raise ValueError(kind)


def run_rows(R, K, Nn, kind, epi, g):
    A, Ad = make_opnd(kind, R, K, g)
    W = torch.randn(Nn, K, generator=g, device=dev) / K ** 0.5
    Wp, _, _ = tc.prep_weight(W)
    bias = torch.randn(Nn, generator=g, device=dev)
    ref = Ad @ W.double().t()
    res = {}
    for gen in (1, 0):
        N.lib().rsb_tc_set_generation(gen)
        Y = torch.empty(R, Nn, device=dev)
        if epi == "stats":
            st = torch.zeros(2 * Nn, dtype=torch.float64, device=dev)
            fn = lambda: tc.gemm_rows(R, Nn, A, Wp, Y=Y, bias=bias, stats=st)  # noqa: E731
            t = timeit(fn)
            st.zero_()
            fn()
            want = ref + bias.double()
            e = rel(Y, want)
            es = max(rel(st[:Nn], want.sum(0)), rel(st[Nn:], (want * want).sum(0)))
        elif epi in ("mask", "maskdual"):
            dual = epi == "maskdual"
            Yl = torch.randn(R, Nn * (2 if dual else 1), generator=g, device=dev)
            w = Nn * (2 if dual else 1)
            sc = torch.rand(w, generator=g, device=dev) + 0.5
            sh = torch.randn(w, generator=g, device=dev) * 0.3
            mu = torch.randn(w, generator=g, device=dev) * 0.1
            inv = torch.rand(w, generator=g, device=dev) + 0.5
            st = torch.zeros((3 if dual else 2) * Nn, dtype=torch.float64, device=dev)
            fn = lambda: tc.gemm_rows(R, Nn, A, Wp, Y=Y, stats=st, mask=(Yl, sc, sh, mu, inv, dual))  # noqa: E731
            t = timeit(fn)
            st.zero_()
            fn()
            z = Yl.double()[:, :Nn] * sc.double()[:Nn] + sh.double()[:Nn]
            if dual:
                z = z + Yl.double()[:, Nn:] * sc.double()[Nn:] + sh.double()[Nn:]
            # exclude elements whose mask is numerically undecidable
            zf = Yl[:, :Nn] * sc[:Nn] + sh[:Nn]
            if dual:
                zf = zf + (Yl[:, Nn:] * sc[Nn:] + sh[Nn:])
            want = torch.where(zf > 0, ref, torch.zeros_like(ref))
            dec = (z.abs() > 1e-4).double()          # elements whose mask fp32 and fp64 agree on
            e = rel(Y.double() * dec, want * dec)
            xh = (Yl.double()[:, :Nn] - mu.double()[:Nn]) * inv.double()[:Nn]
            es = max(rel(st[:Nn], want.sum(0)), rel(st[Nn:2 * Nn], (want * xh).sum(0)))
            if dual:
                xh2 = (Yl.double()[:, Nn:] - mu.double()[Nn:]) * inv.double()[Nn:]
                es = max(es, rel(st[2 * Nn:], (want * xh2).sum(0)))
        else:
            fn = lambda: tc.gemm_rows(R, Nn, A, Wp, Y=Y)  # noqa: E731
            t = timeit(fn)
            e, es = rel(Y, ref), 0.0
        res[gen] = (t, e, es)
    N.lib().rsb_tc_set_generation(0)
    npieces = 2 if kind in ("dual", "aff") else 1
    byts = R * 4 * (K * npieces + Nn + (Nn * (2 if epi == "maskdual" else 1) if epi.startswith("mask") else 0))
    print(f"rows  R={R:8d} K={K:4d} N={Nn:4d} {kind:5s} {epi:8s} | v1 {res[1][0]:7.3f} ms err {res[1][1]:.1e}/{res[1][2]:.1e} | "
          f"v2 {res[0][0]:7.3f} ms err {res[0][1]:.1e}/{res[0][2]:.1e} | x{res[1][0] / res[0][0]:.2f}  v2 {byts / res[0][0] / 1e6:7.1f} GB/s "
          f"({byts / res[0][0] / 1e6 / PEAK:.2f})", flush=True)


def run_wgrad(R, M, Nn, gk, xk, g):
    G, Gd = make_opnd(gk, R, M, g)
    X, Xd = make_opnd(xk, R, Nn, g)
    ref = Gd.t() @ Xd
    res = {}
    for gen in (1, 0):
        N.lib().rsb_tc_set_generation(gen)
        dW = torch.zeros(M, Nn, device=dev)
        fn = lambda: tc.gemm_wgrad(R, G, X, dW)  # noqa: E731
        t = timeit(fn)
        dW.zero_()
        fn()
        res[gen] = (t, rel(dW, ref))
    N.lib().rsb_tc_set_generation(0)
    pieces = lambda k: 2 if k in ("dual", "aff", "affwrap") else 1  # noqa: E731
    byts = R * 4 * (M * pieces(gk) * (0.5 if gk == "affwrap" else 1) + (M if gk == "affwrap" else 0) * 0 + Nn * pieces(xk))
    print(f"wgrad R={R:8d} M={M:4d} N={Nn:4d} {gk:7s} {xk:5s}    | v1 {res[1][0]:7.3f} ms err {res[1][1]:.1e} | v2 {res[0][0]:7.3f} ms err {res[0][1]:.1e} | "
          f"x{res[1][0] / res[0][0]:.2f}  v2 {byts / res[0][0] / 1e6:7.1f} GB/s ({byts / res[0][0] / 1e6 / PEAK:.2f})", flush=True)


def main():
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    g = torch.Generator(device=dev).manual_seed(0)
    small = [(1000, 32, 32, "raw", "stats"), (1000, 20, 64, "raw", "stats"), (777, 64, 32, "dual", "stats"), (5000, 96, 128, "bn", "stats"),
             (3000, 64, 64, "aff", "mask"), (3000, 32, 64, "aff", "maskdual"), (2000, 272, 512, "bn", "stats"), (129, 40, 48, "bn", "none"),
             (300, 256, 272, "aff", "none"), (3000, 512, 256, "raw", "mask"), (3000, 256, 256, "aff", "mask"), (1000, 1024, 512, "raw", "mask")]
    for c in small:
        run_rows(*c, g)
    smallw = [(9000, 32, 32, "aff", "dual"), (10001, 64, 20, "affwrap", "raw"), (20000, 64, 64, "raw", "bn"), (8200, 32, 64, "bn", "aff"),
              (1000, 64, 20, "aff", "raw"), (1000, 32, 64, "aff", "dual"), (2077, 64, 32, "raw", "bn"), (4000, 128, 128, "affwrap", "raw"),
              (3000, 512, 272, "aff", "bn"), (999, 256, 144, "aff", "dual")]
    for c in smallw:
        run_wgrad(*c, g)
    if quick:
        return
    R1, R2, R3, R4 = 2621440, 655360, 163840, 40960
    big = [(R1, 20, 64, "raw", "stats"), (R1, 32, 32, "dual", "stats"), (R1, 32, 64, "bn", "stats"),
           (R1, 64, 32, "raw", "mask"), (R1, 32, 32, "aff", "maskdual"), (R1, 32, 16, "aff", "none"),
           (R2, 80, 128, "raw", "stats"), (R2, 64, 64, "dual", "stats"), (R2, 64, 128, "bn", "stats"), (R2, 128, 64, "raw", "mask"),
           (R3, 144, 256, "raw", "stats"), (R3, 128, 256, "bn", "stats"), (R4, 272, 512, "raw", "stats"), (R4, 256, 512, "bn", "stats"),
           (327680, 128, 128, "raw", "stats"), (327680, 128, 128, "aff", "none"), (R4, 512, 256, "raw", "mask"), (R4, 256, 256, "aff", "mask")]
    for c in big:
        run_rows(*c, g)
    bigw = [(R1, 64, 20, "affwrap", "raw"), (R1, 32, 32, "aff", "dual"), (R1, 64, 32, "raw", "bn"),
            (R2, 128, 80, "affwrap", "raw"), (R2, 64, 64, "aff", "dual"), (R2, 128, 64, "raw", "bn"),
            (R3, 256, 144, "affwrap", "raw"), (R3, 256, 128, "raw", "bn"), (R4, 512, 272, "affwrap", "raw"), (R4, 512, 256, "raw", "bn"),
            (327680, 128, 128, "aff", "raw")]
    for c in bigw:
        run_wgrad(*c, g)


if __name__ == "__main__":
    main()
