"""The training steps captured in a CUDA graph (dense classification; packed segmentation with fixed offsets): same weights
after the same steps as the eager loop."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _inputs(seed, B=8, N=1024, dev="cuda"):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, 3, N, generator=g).to(dev), torch.randint(0, 15, (B,), generator=g).to(dev)


def _make(dev):
    from repsurf_b200.models import RepSurfCls
    torch.manual_seed(3)
    model = RepSurfCls().to(dev).train()
    model.surface_constructor.random_inv = False          # the two loops must see the same draws: none
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    return model


def test_graphed_train_step_matches_eager_loop():
    from repsurf_b200.graph import GraphedTrainStep
    from repsurf_b200.models import SmoothClsLoss
    dev = torch.device("cuda")
    crit = SmoothClsLoss()
    eager = _make(dev)
    graphed = copy.deepcopy(eager)
    batches = [_inputs(10 + i) for i in range(4)]
    init = {n: v.clone() for n, v in eager.state_dict().items()}
    opt_e = torch.optim.SGD(eager.parameters(), lr=1e-4, momentum=0.9, weight_decay=1e-4)
    opt_g = torch.optim.SGD(graphed.parameters(), lr=1e-4, momentum=0.9, weight_decay=1e-4)

    # the graphed loop warms up on batch 0 three times before capturing, and the capture itself does not execute; the eager
    # loop does the same three steps so that both start the compared steps from the same weights / momentum / running stats
    step = GraphedTrainStep(graphed, crit, opt_g, [batches[0][0]], batches[0][1], warmup=3)
    assert step.launches_per_step > 50
    for _ in range(3):
        opt_e.zero_grad(set_to_none=True)
        crit(eager(batches[0][0]), batches[0][1]).backward()
        opt_e.step()

    def same_update(a, b, ref, n):
        # two runs of the SAME eager step already differ (order of the fp32 atomics in the scatter epilogues, amplified through
        # ReLU / max-pool routing and the 8-sample BatchNorm of the head: a few 1e-3 of the first layers' gradients, seen once
        # above rtol 1e-4 on a weight after three steps), so updates are compared, with room for that noise; a step that was
        # dropped, doubled or replayed with stale inputs moves them by O(1)
        da, db = (a - ref).double(), (b - ref).double()
        if a.dim() > 1:
            assert float(da.norm()) > 0, n                  # every weight moved (a bias in front of a BatchNorm has no gradient)
        floor = 2e-7 * (1.0 + float(ref.double().norm()))   # fp32 resolution of the parameter itself: updates below it are rounding
        assert float((da - db).norm()) <= 0.15 * float(da.norm()) + floor, (n, float((da - db).norm()), float(da.norm()))

    start = {n: v.clone() for n, v in eager.state_dict().items()}
    for (n, a), b in zip(start.items(), graphed.state_dict().values()):
        if a.dtype.is_floating_point:
            same_update(a, b, init[n], n)                   # same state after the three warm-up steps
        else:
            assert torch.equal(a, b), n
    losses_e, losses_g = [], []
    for x, y in batches:
        opt_e.zero_grad(set_to_none=True)
        le = crit(eager(x), y)
        le.backward()
        opt_e.step()
        losses_e.append(float(le.detach()))
        losses_g.append(float(step([x], y).detach()))
    # same kernels with the same launch plans; run to run only the order of the fp32 atomics in the scatter / statistics
    # epilogues differs.  The learning rate is small so that this noise is not amplified through the 8-sample BatchNorm of the
    # head from step to step: what is compared is four forward passes and the four accumulated SGD-momentum updates.
    assert losses_g == pytest.approx(losses_e, rel=1e-3)
    for (n, a), b in zip(eager.state_dict().items(), graphed.state_dict().values()):
        if not a.dtype.is_floating_point:
            assert torch.equal(a, b), n                     # num_batches_tracked advances on replay too
            continue
        same_update(a, b, start[n], n)


def test_graphed_step_draws_fresh_randomness_each_replay():
    """Dropout and the umbrella's random flip must not be frozen into the graph."""
    from repsurf_b200.graph import GraphedTrainStep
    from repsurf_b200.models import RepSurfCls, SmoothClsLoss
    dev = torch.device("cuda")
    torch.manual_seed(5)
    model = RepSurfCls().to(dev).train()
    opt = torch.optim.SGD(model.parameters(), lr=0.0)       # weights fixed: only the draws differ between replays
    for m in model.modules():
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            m.momentum = 0.0                                 # and running statistics play no part in train mode anyway
    x, y = _inputs(77)
    step = GraphedTrainStep(model, SmoothClsLoss(), opt, [x], y)
    vals = {round(float(step([x], y)), 7) for _ in range(6)}
    assert len(vals) > 1


# ------------------------------------------------------------------------------------------------ packed segmentation
SEG_SIZES = (12000, 5000)          # cloud 0 takes the sectorized FPS (>= 10000 points), both take the grid kNN


def _seg_inputs(seed, dev="cuda"):
    from repsurf_b200.seg import pointops as PS
    g = torch.Generator().manual_seed(seed)
    n = sum(SEG_SIZES)
    coord = torch.rand(n, 3, generator=g) * torch.tensor([6.0, 6.0, 3.0])
    feat = torch.randn(n, 3, generator=g)
    target = torch.randint(0, 13, (n,), generator=g)
    ends, run = [], 0
    for s_ in SEG_SIZES:
        run += s_
        ends.append(run)
    offset = PS.make_offsets(ends, torch.device(dev))
    return coord.to(dev), feat.to(dev), offset, target.to(dev)


def _make_seg(dev):
    from repsurf_b200.models import RepSurfSeg
    torch.manual_seed(4)
    model = RepSurfSeg().to(dev).train()
    model.surface_constructor.random_inv = False          # the two loops must see the same draws: none
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    return model


@pytest.mark.parametrize("opt_in_graph", [True, False])
def test_graphed_seg_step_matches_eager_loop(opt_in_graph):
    """forward + backward (+ SGD) of the packed model as one graph, geometry-plan side streams included; with the optimizer
    outside the graph (the N > 1 arrangement) the eager tail runs on the gradients the replay wrote, also after an eager
    backward in between has re-pointed p.grad."""
    from repsurf_b200.graph import graphed_seg_step
    from repsurf_b200.seg.loss import CrossEntropyLoss
    dev = torch.device("cuda")
    crit = CrossEntropyLoss()
    eager = _make_seg(dev)
    graphed = copy.deepcopy(eager)
    batches = [_seg_inputs(20 + i) for i in range(4)]
    init = {n: v.clone() for n, v in eager.state_dict().items()}
    opt_e = torch.optim.SGD(eager.parameters(), lr=1e-4, momentum=0.9, weight_decay=1e-4)
    opt_g = torch.optim.SGD(graphed.parameters(), lr=1e-4, momentum=0.9, weight_decay=1e-4)
    calls = []
    c0, f0, o0, t0 = batches[0]
    step = graphed_seg_step(graphed, crit, opt_g, c0, f0, o0, t0, optimizer_in_graph=opt_in_graph,
                            after_backward=(lambda: calls.append(1)) if not opt_in_graph else None, warmup=3)
    assert step.launches_per_step > 100
    for _ in range(3):
        opt_e.zero_grad(set_to_none=True)
        crit(eager([c0, f0, o0]), t0).backward()
        opt_e.step()

    def same_update(a, b, ref, n):
        da, db = (a - ref).double(), (b - ref).double()
        if a.dim() > 1:
            assert float(da.norm()) > 0, n                  # every weight moved (a bias in front of a BatchNorm has no gradient)
        floor = 2e-7 * (1.0 + float(ref.double().norm()))   # fp32 resolution of the parameter itself: updates below it are rounding
        assert float((da - db).norm()) <= 0.15 * float(da.norm()) + floor, (n, float((da - db).norm()), float(da.norm()))

    start = {n: v.clone() for n, v in eager.state_dict().items()}
    for (n, a), b in zip(start.items(), graphed.state_dict().values()):
        if a.dtype.is_floating_point:
            same_update(a, b, init[n], n)
        else:
            assert torch.equal(a, b), n
    losses_e, losses_g = [], []
    for i, (c, f, o, t) in enumerate(batches):
        opt_e.zero_grad(set_to_none=True)
        le = crit(eager([c, f, o]), t)
        le.backward()
        opt_e.step()
        losses_e.append(float(le.detach()))
        if i == 2:                                          # an eager backward through the graphed model re-points p.grad ...
            opt_g.zero_grad(set_to_none=True)
            crit(graphed([c, f, o]), t).backward()
            opt_g.zero_grad(set_to_none=True)               # ... and leaves without stepping
        losses_g.append(float(step([c, f, o], t).detach()))
    assert losses_g == pytest.approx(losses_e, rel=1e-3)
    if not opt_in_graph:
        assert len(calls) == 3 + len(batches)               # after_backward: once per warm-up step and per call
    for (n, a), b in zip(eager.state_dict().items(), graphed.state_dict().values()):
        if not a.dtype.is_floating_point:
            continue                                        # num_batches_tracked: the graphed model ran one extra forward
        if "running_" in n:
            continue                                        # ... which also moved its running statistics once more
        same_update(a, b, start[n], n)
    # other per-cloud sizes than the captured ones are refused, not silently replayed with stale launch plans
    from repsurf_b200.seg import pointops as PS
    other = PS.make_offsets([11000, 17000], dev)
    with pytest.raises(RuntimeError):
        step([c0, f0, other], t0)
