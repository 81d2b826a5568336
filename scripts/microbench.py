"""BASELINE.json config 5: FPS / ball query / kNN / grouping sweep, N = 1k..128k points per cloud, nsample 8..64,
m = N/4 centres, B clouds with B*N ~ 327 680 - this package's sm_100a kernels against the REFERENCE's own CUDA kernels
(oracle/_ref/libref_pointops_cls.so, compiled unmodified from /root/reference by oracle/build_ref.sh) on the same GPU and
the same inputs.  CUDA-event time, best of 3 after one warm-up.  `python bench.py --workload micro` -> one JSON line +
gpurun_out/microbench.{json,md}."""
import ctypes
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _time(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    from repsurf_b200 import _native as N
    from repsurf_b200.cls import pointops as P
    from tests import refcuda as R
    dev = torch.device("cuda")
    have_ref = R.available("cls")
    _i, _f = ctypes.c_int, ctypes.c_float
    p = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    rows = []
    total = 327680
    for n in (1024, 2048, 4096, 8192, 16384, 32768, 65536, 131072):
        b = max(1, total // n)
        m = n // 4
        g = torch.Generator().manual_seed(n)
        xyz = torch.rand(b, n, 3, generator=g).to(dev)
        # ---------------- FPS
        idx = torch.empty(b, m, dtype=torch.int32, device=dev)
        t_us = _time(lambda: N.call("rsb_furthestsampling_dense", b, n, m, xyz, None if n + 1024 <= 16 * 512 * 16 else torch.empty(b, n, device=dev), idx, None))
        t_ref = None
        if have_ref:
            ridx = torch.zeros(b, m, dtype=torch.int32, device=dev)
            tmp = torch.empty(b, n, device=dev)

            def ref_fps():      # the reference's wrapper refills the scratch on every call (cls/po/functions/pointops.py:45)
                tmp.fill_(1e10)
                R.lib("cls").furthestsampling_cuda_launcher(_i(b), _i(n), _i(m), p(xyz), p(tmp), p(ridx))
            t_ref = _time(ref_fps, reps=1 if n >= 32768 else 2)
            same = bool(torch.equal(idx, ridx))
        alg = b * ((m - 1) * n * 20 + 4 * m)
        rows.append(dict(op="fps", B=b, N=n, m=m, ms=t_us, ref_ms=t_ref, identical=same if have_ref else None,
                         us_per_sample=1e3 * t_us / m, alg_gbs=alg / t_us / 1e6))
        new_xyz = torch.gather(xyz, 1, idx.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
        for ns in (8, 16, 32, 64):
            # ---------------- ball query: radius for ~ns expected neighbours in the unit cube
            r = (3.0 * ns / (4.0 * math.pi * n)) ** (1.0 / 3.0)
            bi = torch.empty(b, m, ns, dtype=torch.int32, device=dev)
            t_us = _time(lambda: N.call("rsb_ballquery", b, n, m, float(r), ns, new_xyz, xyz, bi))
            t_ref, same = None, None
            if have_ref:
                rbi = torch.zeros(b, m, ns, dtype=torch.int32, device=dev)
                t_ref = _time(lambda: R.lib("cls").ballquery_cuda_launcher_fast(_i(b), _i(n), _i(m), _f(r), _i(ns), p(new_xyz), p(xyz), p(rbi), ctypes.c_void_p(0)), reps=2)
                same = bool(torch.equal(bi, rbi))
            rows.append(dict(op="ballquery", B=b, N=n, m=m, nsample=ns, ms=t_us, ref_ms=t_ref, identical=same,
                             gpairs_per_s=b * m * n / t_us / 1e6))
            # ---------------- kNN
            ki = P.knnquery(ns, xyz, new_xyz)
            t_us = _time(lambda: P.knnquery(ns, xyz, new_xyz))
            t_ref, same = None, None
            if have_ref:
                rki = torch.zeros(b, m, ns, dtype=torch.int32, device=dev)
                rd2 = torch.zeros(b, m, ns, dtype=torch.float32, device=dev)
                t_ref = _time(lambda: R.lib("cls").knnquery_cuda_launcher(_i(b), _i(n), _i(m), _i(ns), p(xyz), p(new_xyz), p(rki), p(rd2), ctypes.c_void_p(0)), reps=1)
                same = bool(torch.equal(ki, rki))
            rows.append(dict(op="knn", B=b, N=n, m=m, nsample=ns, ms=t_us, ref_ms=t_ref, identical=same,
                             gpairs_per_s=b * m * n / t_us / 1e6))
            # ---------------- grouping (channel-first dense API), C = 64; the C sweep at nsample = 32
            for c in ((3, 16, 64, 128) if ns == 32 else (64,)):
                feat = torch.randn(b, c, n, generator=g).to(dev)
                out = torch.empty(b, c, m, ns, device=dev)
                t_us = _time(lambda: N.call("rsb_grouping_forward", b, c, n, m, ns, feat, ki, out))
                t_ref, same = None, None
                if have_ref:
                    rout = torch.empty(b, c, m, ns, device=dev)
                    t_ref = _time(lambda: R.lib("cls").grouping_forward_cuda_launcher_fast(_i(b), _i(c), _i(n), _i(m), _i(ns), p(feat), p(ki), p(rout)), reps=2)
                    same = bool(torch.equal(out, rout))
                byts = b * m * ns * (4 + 8 * c)
                rows.append(dict(op="grouping", B=b, N=n, m=m, nsample=ns, C=c, ms=t_us, ref_ms=t_ref, identical=same,
                                 alg_gbs=byts / t_us / 1e6))
        print(f"N={n} done", file=sys.stderr, flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "microbench.json"), "w"), indent=0)
    with open(os.path.join(ROOT, "gpurun_out", "microbench.md"), "w") as f:
        f.write("| op | B | N | m | nsample | C | ours ms | reference CUDA ms | speed-up | identical | note |\n|---|---|---|---|---|---|---|---|---|---|---|\n")
        for r in rows:
            note = (f"{r['us_per_sample']:.2f} us/sample, {r['alg_gbs']:.0f} GB/s (streaming model)" if r["op"] == "fps" else
                    f"{r['gpairs_per_s']:.1f} G pairs/s (all-pairs equivalent)" if "gpairs_per_s" in r else f"{r['alg_gbs']:.0f} GB/s")
            sp = f"{r['ref_ms'] / r['ms']:.1f}x" if r.get("ref_ms") else "-"
            f.write(f"| {r['op']} | {r['B']} | {r['N']} | {r['m']} | {r.get('nsample', '-')} | {r.get('C', '-')} | {r['ms']:.3f} | "
                    f"{(r['ref_ms'] if r.get('ref_ms') else float('nan')):.3f} | {sp} | {r.get('identical')} | {note} |\n")
    geo = {}
    for op in ("fps", "ballquery", "knn", "grouping"):
        sp = [r["ref_ms"] / r["ms"] for r in rows if r["op"] == op and r.get("ref_ms")]
        if sp:
            geo[op] = math.exp(sum(math.log(x) for x in sp) / len(sp))
    print(json.dumps({"metric": "pointops microbench vs reference CUDA kernels (same GPU)", "n_cases": len(rows),
                      "all_identical": all(r.get("identical") in (True, None) for r in rows), "geomean_speedup": geo,
                      "table": "gpurun_out/microbench.md"}))


if __name__ == "__main__":
    main()
