"""Fused flat-buffer Adam (vog_adam_step / optim.FlatAdam) against torch.optim.Adam with the
reference's settings (betas (0.9, 0.99), code/main_dist.py:55; lr 1e-4, configs/anet_srl_cfg.yml:108) run on the CPU."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import vognet_pytorch_b200 as vb                      # noqa: E402
from vognet_pytorch_b200 import synth                  # noqa: E402
from vognet_pytorch_b200.optim import FlatAdam         # noqa: E402

DEV = 'cuda:0'


def test_flat_adam_matches_torch_adam():
    g = torch.Generator().manual_seed(0)
    shapes = [(7,), (1000,), (33, 17), (5, 4, 3), (1,), (256, 130)]
    ref = [torch.randn(s, generator=g).requires_grad_(True) for s in shapes]
    mine = [r.detach().clone().to(DEV).requires_grad_(True) for r in ref]
    opt_ref = torch.optim.Adam(ref, lr=1e-3, betas=(0.9, 0.99))
    opt = FlatAdam(mine, lr=1e-3, betas=(0.9, 0.99))
    for p, r in zip(mine, ref):
        assert torch.equal(p.detach().cpu(), r.detach())          # flattening kept the values
    for step in range(6):
        for p, r in zip(mine, ref):
            gr = torch.randn(r.shape, generator=g) * (10.0 ** (step % 3 - 1))
            r.grad = gr.clone()
            p.grad.copy_(gr)                                        # the .grad views of the flat buffer
        opt_ref.step()
        opt.step()
        torch.cuda.synchronize()
        for p, r in zip(mine, ref):
            # a parameter moves by ~lr = 1e-3 per step; agreement to a few fp32 ulps of the parameter value
            d = (p.detach().cpu() - r.detach()).abs()
            assert (d <= 1e-7 + 1e-6 * r.detach().abs()).all(), (step, tuple(r.shape), float(d.max()))
    st = opt_ref.state[ref[2]]
    o = opt.offsets[2]
    # moments: to a few ulps of their largest element (m cancels, so small entries carry the absolute error)
    m_, v_ = st['exp_avg'], st['exp_avg_sq']
    assert (opt.exp_avg[o:o + m_.numel()].cpu().view_as(m_) - m_).abs().max() <= 2e-6 * m_.abs().max()
    assert (opt.exp_avg_sq[o:o + v_.numel()].cpu().view_as(v_) - v_).abs().max() <= 2e-6 * v_.abs().max()
    opt.zero_grad()
    assert all(float(p.grad.abs().sum()) == 0 for p in mine)


def test_flat_adam_on_the_model_keeps_it_working():
    """All 45 M parameters of VOG_SPAT in one buffer: state_dict unchanged, the forward still reads the (updated)
    parameters through its views, one launch per step."""
    from vognet_pytorch_b200 import _lib
    w, batch = synth.workload('cpu_ref')
    cfg, comm = synth.default_cfg('spat'), synth.default_comm(w['nppf'])
    mdl = vb.get_mdl_loss_eval(cfg)['mdl'](cfg, comm)
    sd = synth.make_state_dict()
    mdl.load_state_dict(sd, strict=True)
    mdl = mdl.to(DEV).eval().set_compute('tf32')
    dbatch = synth.clone_batch(batch, DEV)
    before = mdl(dbatch)['mdl_outs'].clone()
    opt = FlatAdam(mdl.parameters(), lr=1e-4, betas=(0.9, 0.99))
    assert sum(p.numel() for p in mdl.parameters()) == 45250727 and opt.numel >= 45250727
    assert list(mdl.state_dict().keys()) == list(sd.keys())
    assert torch.equal(mdl(dbatch)['mdl_outs'], before)              # flattening alone changes nothing
    opt.flat_grad.fill_(1.0)
    n0 = _lib.lib().vog_launch_count()
    opt.step()
    assert _lib.lib().vog_launch_count() - n0 == 1
    torch.cuda.synchronize()
    # first Adam step with a constant gradient moves every parameter by -lr (m/sqrt(v) = 1 after bias correction)
    k = 'lin2.2.weight'
    assert torch.allclose(mdl.state_dict()[k].cpu(), sd[k] - 1e-4, rtol=0, atol=2e-7)
    after = mdl(dbatch)['mdl_outs']
    assert not torch.equal(after, before) and torch.isfinite(after).all()
    # the step wrote the parameters through a raw pointer: every packed low-precision weight copy must have been
    # rebuilt - the forward has to equal the one of a FRESH model built from the updated state_dict
    fresh = vb.get_mdl_loss_eval(cfg)['mdl'](cfg, comm)
    fresh.load_state_dict({k: v.detach().cpu().clone() for k, v in mdl.state_dict().items()}, strict=True)
    fresh = fresh.to(DEV).eval().set_compute('tf32')
    assert torch.equal(fresh(dbatch)['mdl_outs'], after)


@pytest.mark.parametrize('mode', ['tf32', 'bf16'])
def test_weight_updates_reach_a_captured_graph(mode):
    """A CUDA-graph forward must follow the parameters: after an optimizer step, a load_state_dict and an in-place
    edit, the replay equals a fresh eager model with the same weights (packed copies are refreshed in place, the
    captured addresses stay valid)."""
    w, batch = synth.workload('cpu_ref')
    cfg, comm = synth.default_cfg('spat'), synth.default_comm(w['nppf'])
    sd = synth.make_state_dict()

    def build(state, graph):
        m = vb.get_mdl_loss_eval(cfg)['mdl'](cfg, comm)
        m.load_state_dict(state, strict=True)
        m = m.to(DEV).eval().set_compute(mode)
        m.use_cuda_graph = graph
        return m
    mdl = build(sd, True)
    dbatch = synth.clone_batch(batch, DEV)
    first = mdl(dbatch)['mdl_outs'].clone()
    assert torch.equal(first, build(sd, False)(dbatch)['mdl_outs'])
    opt = FlatAdam(mdl.parameters(), lr=1e-3, betas=(0.9, 0.99))
    assert torch.equal(mdl(dbatch)['mdl_outs'], first)              # re-homing the parameters changes nothing
    g = torch.Generator().manual_seed(3)
    opt.flat_grad.copy_(torch.randn(opt.numel, generator=g).to(DEV))
    opt.step()
    cur = {k: v.detach().cpu().clone() for k, v in mdl.state_dict().items()}
    out = mdl(dbatch)['mdl_outs'].clone()
    assert not torch.equal(out, first)
    assert torch.equal(out, build(cur, False)(dbatch)['mdl_outs'])
    # load_state_dict back to the original weights (in-place copies into the flat buffer)
    mdl.load_state_dict(sd, strict=True)
    assert torch.equal(mdl(dbatch)['mdl_outs'], first)
    # plain in-place edit of one parameter
    with torch.no_grad():
        mdl.lin2[0].weight.mul_(1.5)
    cur = {k: v.detach().cpu().clone() for k, v in mdl.state_dict().items()}
    assert torch.equal(mdl(dbatch)['mdl_outs'], build(cur, False)(dbatch)['mdl_outs'])
    assert len(mdl._graphs) == 1


def test_flat_adam_is_a_torch_optimizer_with_adam_checkpoints():
    """utils/trn_utils.py:807-818 wraps the optimizer in LambdaLR / ReduceLROnPlateau and :610,622 save / load its
    state_dict: FlatAdam takes both, in torch.optim.Adam's layout."""
    g = torch.Generator().manual_seed(0)
    shapes = [(33, 17), (5,), (4, 3, 2)]
    ref = [torch.randn(s, generator=g).requires_grad_(True) for s in shapes]
    mine = [r.detach().clone().to(DEV).requires_grad_(True) for r in ref]
    opt_ref = torch.optim.Adam(ref, lr=1e-3, betas=(0.9, 0.99))
    opt = FlatAdam(mine, lr=1e-3, betas=(0.9, 0.99))
    assert isinstance(opt, torch.optim.Optimizer)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda e: 0.5 ** e)
    sched_ref = torch.optim.lr_scheduler.LambdaLR(opt_ref, lambda e: 0.5 ** e)

    def one_step():
        for p, r in zip(mine, ref):
            gr = torch.randn(r.shape, generator=g)
            r.grad = gr.clone()
            p.grad.copy_(gr)
        opt_ref.step(); opt.step()
        sched_ref.step(); sched.step()
    for _ in range(3):
        one_step()
    assert opt.lr == opt_ref.param_groups[0]['lr'] == 1e-3 * 0.125      # scheduler changes reach the kernel
    # torch Adam -> FlatAdam: resume a "reference checkpoint"
    mine2 = [r.detach().clone().to(DEV).requires_grad_(True) for r in ref]
    opt2 = FlatAdam(mine2, lr=1.0, betas=(0.5, 0.5))
    opt2.load_state_dict(opt_ref.state_dict())
    assert opt2.step_count == 3 and opt2.lr == opt_ref.param_groups[0]['lr'] and opt2.betas == (0.9, 0.99)
    # FlatAdam -> torch Adam
    ref3 = [r.detach().clone().requires_grad_(True) for r in ref]
    opt3 = torch.optim.Adam(ref3, lr=1.0)
    opt3.load_state_dict(opt.state_dict())
    for a, b in zip(opt_ref.state_dict()['state'].values(), opt3.state_dict()['state'].values()):
        assert float(a['step']) == float(b['step'])
        assert (a['exp_avg'] - b['exp_avg']).abs().max() <= 2e-6 * a['exp_avg'].abs().max()
    # both continue identically
    for p, r, q in zip(mine2, ref, ref3):
        gr = torch.randn(r.shape, generator=g)
        r.grad = gr.clone(); q.grad = gr.clone(); p.grad.copy_(gr)
    opt_ref.step(); opt2.step(); opt3.step()
    torch.cuda.synchronize()
    for p, r, q in zip(mine2, ref, ref3):
        assert (p.detach().cpu() - r.detach()).abs().max() <= 1e-7 + 1e-6 * r.detach().abs().max()
        assert (q.detach() - r.detach()).abs().max() <= 1e-7 + 1e-6 * r.detach().abs().max()


def test_flat_adam_rejects_cpu_parameters():
    with pytest.raises(RuntimeError):
        FlatAdam([torch.zeros(3, requires_grad=True)])
