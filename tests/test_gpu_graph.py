"""GPU: the training step as a CUDA graph (graph.GraphedStep; SURVEY.md 8d "CUDA-graph-replayed step").  The reference's
loop is eager (Our_UNet/src/train.py:630-670); a replay must compute exactly what the eager step computes."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(dropout):
    from unet_implementations_b200.models.unet import UNet
    torch.manual_seed(1234)
    rates = dict(encoder_dropout_rates=[0, 0, 0.1, 0.2], decoder_dropout_rates=[0.2, 0.1, 0]) if dropout else \
        dict(encoder_dropout_rates=[0, 0, 0, 0], decoder_dropout_rates=[0, 0, 0])
    model = UNet(n_stages=4, features_per_stage=[32, 64, 128, 128], **rates).cuda().train()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 3, 64, 64, generator=g).cuda()
    t = torch.randint(0, 3, (2, 64, 64), generator=g).cuda()
    return model, x, t


def test_replay_of_forward_loss_backward_is_bit_identical_to_the_eager_step():
    from unet_implementations_b200.graph import GraphedStep
    from unet_implementations_b200.models.losses import SimpleLoss
    model, x, t = _setup(dropout=False)
    loss_fn = SimpleLoss()
    xs = x.clone()

    def step():
        model.zero_grad(set_to_none=True)
        loss = loss_fn(model(xs), t)
        loss.backward()
        return loss

    ref_loss = step().detach().clone()  # (keeping the loss itself would keep the eager step's autograd graph -- and its
    ref_grads = [p.grad.clone() for p in model.parameters()]  # default-stream AccumulateGrad nodes -- alive into the capture)
    gs = GraphedStep(step)
    out = gs.replay()
    torch.cuda.synchronize()
    assert torch.equal(out.detach(), ref_loss)
    for p, g in zip(model.parameters(), ref_grads):
        assert torch.equal(p.grad, g)
    # new data through the static input buffer
    xs.copy_(x * 0.5)
    out2 = gs.replay().detach().clone()
    e = step().detach()
    assert torch.equal(out2, e) and not torch.equal(out2, ref_loss)


def test_replays_draw_fresh_dropout_masks():
    from unet_implementations_b200.graph import GraphedStep
    from unet_implementations_b200.models.losses import SimpleLoss
    model, x, t = _setup(dropout=True)
    loss_fn = SimpleLoss()

    def step():
        model.zero_grad(set_to_none=True)
        loss = loss_fn(model(x), t)
        loss.backward()
        return loss

    gs = GraphedStep(step)
    losses = {float(gs.replay()) for _ in range(4)}
    assert len(losses) > 1 and all(l == l for l in losses)


def test_whole_training_step_with_the_fused_optimizer_in_one_graph_follows_the_lr_schedule():
    """forward + loss + backward + FusedSGD(model=..., capturable=True) as ONE graph; LambdaLR (train.py:454-477) keeps
    working through the device-side learning rate.  Against the same loop run eagerly: bit-identical parameters."""
    from unet_implementations_b200.graph import GraphedStep
    from unet_implementations_b200.models.losses import SimpleLoss
    from unet_implementations_b200.optim import FusedSGD
    ma, x, t = _setup(dropout=False)
    mb = copy.deepcopy(ma)
    loss_fn = SimpleLoss()
    kw = dict(lr=0.005, momentum=0.99, nesterov=True, weight_decay=1e-4)
    oa = torch.optim.SGD(ma.parameters(), **kw)
    ob = FusedSGD(mb.parameters(), model=mb, capturable=True, **kw)
    lam = lambda e: (1 - e / 10) ** 0.9  # noqa: E731
    sa, sb = torch.optim.lr_scheduler.LambdaLR(oa, lam), torch.optim.lr_scheduler.LambdaLR(ob, lam)

    def step(model, opt):
        opt.zero_grad(set_to_none=True)
        loss = loss_fn(model(x), t)
        loss.backward()
        opt.step()
        return loss

    for _ in range(2):  # eager: momentum buffers initialised, steady-state table uploaded
        step(ma, oa), sa.step()
        step(mb, ob), sb.step()
    gs = GraphedStep(lambda: step(mb, ob), warmup=1, optimizer=ob)   # the warm-up step is a real training step
    step(ma, oa)
    for _ in range(4):
        sa.step(), sb.step()
        la = step(ma, oa)
        lb = gs.replay()
        assert torch.equal(la.detach(), lb.detach())
    for pa, pb in zip(ma.parameters(), mb.parameters()):
        assert torch.equal(pa, pb)
