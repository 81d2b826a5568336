"""A/B of the two FPS kernel generations (barrier.cluster vs st.async + mbarrier exchange): identical picks, us/sample."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from repsurf_b200 import _native as N  # noqa: E402
from repsurf_b200.seg import pointops as PS  # noqa: E402

dev = torch.device("cuda")


def run(nseg, n, m, seed=0):
    g = torch.Generator().manual_seed(seed)
    xyz = (torch.rand(nseg * n, 3, generator=g) * torch.tensor([8.0, 8.0, 3.0])).to(dev)
    off = PS.make_offsets([n * (i + 1) for i in range(nseg)], dev)
    noff = PS.make_offsets([m * (i + 1) for i in range(nseg)], dev)
    res = {}
    for gen in (1, 0):
        N.lib().rsb_fps_set_generation(gen)
        idx = PS.furthestsampling(xyz, off, noff)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            a, b = torch.cuda.Event(True), torch.cuda.Event(True)
            a.record()
            idx = PS.furthestsampling(xyz, off, noff)
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        res[gen] = (best, idx.clone())
    N.lib().rsb_fps_set_generation(0)
    same = torch.equal(res[0][1], res[1][1])
    print(f"FPS {nseg:3d} x {n:6d} -> {m:5d} | v1 {res[1][0]:7.3f} ms ({1e3 * res[1][0] / m:.3f} us/sample) | v2 {res[0][0]:7.3f} ms "
          f"({1e3 * res[0][0] / m:.3f} us/sample) | x{res[1][0] / res[0][0]:.2f} | identical {same}", flush=True)


if __name__ == "__main__":
    for c in [(32, 10240, 2560), (8, 10240, 2560), (8, 40960, 10240), (8, 2560, 640), (32, 12000, 3000), (4, 100000, 5000), (3, 9000, 1000)]:
        run(*c)
