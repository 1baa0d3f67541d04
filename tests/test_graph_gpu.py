"""train.GraphedStep on a GPU: the CUDA-graph replay of the WideResNet-40-2 step without CrossNorm must be the same
training step as the eager one -- same loss, same parameters and buffers after several optimizer steps in which graph
replays and eager CrossNorm steps alternate -- and capturing must not disturb the training state."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


def _run(capture, steps=8, cn_prob=0.4, channels_last=False):
    from cnsn_b200.train import GraphedStep, make_optimizer, wrn40_2
    torch.manual_seed(0)
    np.random.seed(0)
    net = wrn40_2(fuse_post=True).to(DEV).train()
    opt, sched = make_optimizer(net, total_steps=steps)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(64, 3, 32, 32, generator=g).to(DEV)
    y = torch.randint(0, 10, (64,), generator=g).to(DEV)
    if channels_last:                                     # the layout train.bench_wrn runs in
        net = net.to(memory_format=torch.channels_last)
        x = x.contiguous(memory_format=torch.channels_last)
    before = {k: v.clone() for k, v in net.state_dict().items()}
    gs = GraphedStep(net, x, y, 1, capture=capture)
    assert (gs.graph is not None) == capture
    for k, v in net.state_dict().items():                 # capture (and its warm-up passes) left the state alone
        assert torch.equal(v, before[k]), k
    torch.manual_seed(2)
    np.random.seed(3)                                     # the coins: a mix of CrossNorm (eager) and plain (graph) steps
    losses = [gs.step(gs.x, gs.y, opt, sched, cn_prob) for _ in range(steps)]
    return losses, {k: v.clone() for k, v in net.state_dict().items()}


@pytest.mark.parametrize("channels_last", [False, True])
def test_graph_replay_is_the_eager_step(channels_last):
    saved = (torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = True, False
    try:
        le, se = _run(False, channels_last=channels_last)
        lg, sg = _run(True, channels_last=channels_last)
    finally:
        torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = saved
    np.random.seed(3)
    coins = [bool(np.random.rand(1) < 0.4) for _ in range(8)]
    assert any(coins) and not all(coins)                  # both kinds of step occurred (the draws above consume more
    assert le[0] == lg[0]                                 # numbers in CrossNorm steps; the first coin is the same)
    for a, b in zip(le, lg):
        assert abs(a - b) <= 1e-4 * max(1.0, abs(a)), (le, lg)
    worst = 0.0
    for k in se:
        if se[k].dtype.is_floating_point:
            d = float((se[k] - sg[k]).abs().max() / se[k].abs().max().clamp_min(1e-12))
            worst = max(worst, d)
        else:
            assert torch.equal(se[k], sg[k]), k
    assert worst <= 2e-3, worst


def _run_r50(capture, steps=3):
    """A small ResNet ([1,1,1,1] bottlenecks) through GraphedStep.step_image_cn: image-space CrossNorm eager in front,
    the network replayed."""
    import cnsn_b200.cnsn as ops
    from cnsn_b200.hosts import ResNet
    from cnsn_b200.train import GraphedStep
    torch.manual_seed(0)
    np.random.seed(0)
    net = ResNet([1, 1, 1, 1], num_classes=10, active_num=1, pos="post", beta=1, crop="neither", cnsn_type="sn", fuse_post=True).to(DEV).train()
    opt = torch.optim.SGD(net.parameters(), 0.05, momentum=0.9, weight_decay=1e-4)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(16, 3, 64, 64, generator=g).to(DEV)
    y = torch.randint(0, 10, (16,), generator=g).to(DEV)
    gs = GraphedStep(net, x, y, 1, capture=capture, loss_fn=lambda n, xx, yy, aug: torch.nn.functional.cross_entropy(n(xx, aug=False), yy))
    torch.manual_seed(2)
    np.random.seed(3)
    losses = [gs.step_image_cn(x, y, opt, None, 0.5, ops, crop="both") for _ in range(steps)]
    return losses, {k: v.clone() for k, v in net.state_dict().items()}


def test_graphed_image_crossnorm_step_is_the_eager_step():
    saved = (torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = True, False
    try:
        le, se = _run_r50(False)
        lg, sg = _run_r50(True)
    finally:
        torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = saved
    for a, b in zip(le, lg):
        assert abs(a - b) <= 1e-4 * max(1.0, abs(a)), (le, lg)
    for k in se:
        if se[k].dtype.is_floating_point:
            assert float((se[k] - sg[k]).abs().max() / se[k].abs().max().clamp_min(1e-12)) <= 2e-3, k
        else:
            assert torch.equal(se[k], sg[k]), k
