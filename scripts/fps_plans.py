"""FPS per-iteration latency under different (cluster size, threads) plans (RSB_FPS_PLAN), S3DIS level shapes."""
import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import torch
    from repsurf_b200.seg import pointops as P
    dev = torch.device("cuda")
    B, N, stride = [int(v) for v in sys.argv[1:4]]
    g = torch.Generator().manual_seed(0)
    xyz = (torch.rand(B * N, 3, generator=g) * torch.tensor([8.0, 8.0, 3.0])).to(dev)
    off = P.make_offsets([N * (i + 1) for i in range(B)], dev)
    noff = P.make_offsets([N // stride * (i + 1) for i in range(B)], dev)
    for _ in range(3): P.furthestsampling(xyz, off, noff)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): P.furthestsampling(xyz, off, noff)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"B={B} N={N} m={N//stride} plan={os.environ.get('RSB_FPS_PLAN','default'):8s} {ms:7.3f} ms  {1e3*ms/(N//stride):6.3f} us/iter")
else:
    for shape in [(8, 10240, 4), (32, 10240, 4), (8, 40960, 4), (8, 2560, 4), (8, 640, 4)]:
        for plan in [None, "1,512", "2,512", "4,512", "4,256", "8,256", "8,512", "8,128", "16,128", "16,256", "16,512"]:
            env = dict(os.environ)
            if plan: env["RSB_FPS_PLAN"] = plan
            subprocess.run([sys.executable, __file__] + [str(v) for v in shape], env=env, timeout=120)
