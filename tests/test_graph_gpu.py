"""The dense classification step captured in a CUDA graph: same weights after the same steps as the eager loop."""
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
        assert float(da.norm()) > 0, n                      # every parameter and running statistic moved
        assert float((da - db).norm()) <= 0.15 * float(da.norm()), (n, float((da - db).norm()), float(da.norm()))

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
