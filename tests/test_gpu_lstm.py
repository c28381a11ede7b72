"""Language side through the persistent LSTM recurrence kernel + tcgen05 projections
(language_encode_tc) against torch/cuDNN packed-sequence LSTM (language_encode) and the golden
language vectors of the reference."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import vognet_pytorch_b200 as vb              # noqa: E402
from vognet_pytorch_b200 import synth          # noqa: E402

DEV = 'cuda:0'


def _model(name):
    w, batch = synth.workload(name)
    cfg, comm = synth.default_cfg(w['conc_type']), synth.default_comm(w['nppf'])
    mdl = vb.get_mdl_loss_eval(cfg)['mdl'](cfg, comm)
    mdl.load_state_dict(synth.make_state_dict(), strict=True)
    return w, batch, mdl.to(DEV).eval()


@pytest.mark.parametrize('name', ['cpu_ref', 'spat_gt5'])
@pytest.mark.parametrize('mode,tol', [('tf32', 2e-3), ('bf16', 2e-2)])
def test_language_side_matches_reference(golden, name, mode, tol):
    g = golden(name)
    w, batch, mdl = _model(name)
    mdl.set_compute(mode)
    db = synth.clone_batch(batch, DEV)
    with torch.no_grad():
        ref_torch = mdl.language_encode(db)                 # torch + cuDNN, packed sequences
        got = mdl.language_encode_tc(db)
    torch.cuda.synchronize()
    B, nsrl = got.shape[:2]
    gold = torch.from_numpy(g['lang']).view(B, nsrl, -1)
    assert (ref_torch.cpu() - gold).abs().max() < 1e-4       # the cuDNN path itself matches the reference
    err = (got.cpu() - gold).abs().max().item()
    print(f'\n[{name}/{mode}] lang max|d| {err:.2e} (values up to {gold.abs().max():.2f})')
    assert err < tol


def _ref_recurrence(gx, whh, lens, T, Bq, H):
    """reference recurrence in float64 on the host"""
    gxh, wh = gx.cpu().double().view(T, Bq, 2, 4 * H), whh.cpu().double()
    ref = torch.zeros(T, Bq, 2 * H, dtype=torch.float64)
    for b in range(Bq):
        n = int(lens[b])
        for d in range(2):
            h = torch.zeros(H, dtype=torch.float64); c = torch.zeros(H, dtype=torch.float64)
            order = range(n) if d == 0 else range(n - 1, -1, -1)
            for t in order:
                gates = gxh[t, b, d] + wh[d] @ h
                i, f, gg, o = gates[:H].sigmoid(), gates[H:2 * H].sigmoid(), gates[2 * H:3 * H].tanh(), gates[3 * H:].sigmoid()
                c = f * c + i * gg
                h = o * c.tanh()
                ref[t, b, d * H:(d + 1) * H] = h
    return ref


@pytest.mark.parametrize('streaming,xmode', [(0, 0), (0, 1), (0, 2), (0, 3), (1, 0)],
                         ids=['resident-tagged', 'resident-flags', 'resident-records', 'resident-pair', 'streaming'])
@pytest.mark.parametrize('lens_list', [[1, 20, 7, 13], [5], [20, 3], [2, 9, 4], [1, 1, 1, 1], [20] * 8,
                                       [3, 17, 20, 1, 8, 12], [4, 9, 20, 1, 13, 7, 2, 18, 5, 11, 16]])
def test_ragged_lengths_and_zero_rows(lens_list, streaming, xmode):
    """lengths 1..20, including full-length and single-token sentences, 1..11 sequences (more than 8 = several
    launches over groups of 8);
    both recurrence kernels (weight-resident: W_hh in registers + shared memory, tagged h exchange;
    weight-streaming: W_hh from L2 every step, counter barrier) against a float64 host recurrence."""
    from vognet_pytorch_b200 import ops, _lib
    w, batch, mdl = _model('spat_gt5')
    mdl.set_compute('tf32')
    T, Bq, H = 20, len(lens_list), 1024
    lens = torch.tensor(lens_list, device=DEV)
    gx = (torch.rand(T * Bq, 8 * H, generator=torch.Generator().manual_seed(3)) - 0.5).to(DEV)
    _, _, whh = mdl._lang_weights(ops.LP_TF32)[0][0]
    _lib.lib().vog_debug_lstm_force_streaming(streaming)
    _lib.lib().vog_debug_lstm_exchange(xmode)
    try:
        outs = [ops.lstm_layer_fwd(gx, whh, lens, T, Bq, ops.LP_TF32).view(T, Bq, 2 * H) for _ in range(3)]
        torch.cuda.synchronize()
    finally:
        _lib.lib().vog_debug_lstm_force_streaming(0)
        _lib.lib().vog_debug_lstm_exchange(4)          # the default (automatic) protocol
    out = outs[0]
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])      # deterministic across launches
    ref = _ref_recurrence(gx, whh, lens, T, Bq, H)
    assert (out.cpu().double() - ref).abs().max() < 1e-3         # tf32 rounding of the stored h
    for b in range(Bq):
        assert out[int(lens[b]):, b].abs().max() == 0 if int(lens[b]) < T else True


@pytest.mark.parametrize('xmode', [0, 2, 3], ids=['tagged', 'records', 'pair'])
def test_resident_kernel_in_cuda_graph_replays(xmode):
    """the exchange buffer is re-zeroed by a kernel inside the captured work, so graph
    replays (same kernel parameters every time) never see stale tags"""
    from vognet_pytorch_b200 import ops, _lib
    _lib.lib().vog_debug_lstm_exchange(xmode)
    try:
        _graph_replay_case(ops)
    finally:
        _lib.lib().vog_debug_lstm_exchange(4)          # the default (automatic) protocol


def _graph_replay_case(ops):
    w, batch, mdl = _model('spat_gt5')
    T, Bq, H = 20, 4, 1024
    lens = torch.tensor([11, 20, 7, 13], device=DEV)
    gx = (torch.rand(T * Bq, 8 * H, generator=torch.Generator().manual_seed(5)) - 0.5).to(DEV)
    _, _, whh = mdl._lang_weights(ops.LP_TF32)[0][0]
    eager = ops.lstm_layer_fwd(gx, whh, lens, T, Bq, ops.LP_TF32).clone()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.graph(g, stream=s):
        out = ops.lstm_layer_fwd(gx, whh, lens, T, Bq, ops.LP_TF32)
    for _ in range(4):
        out.zero_()
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, eager)


@pytest.mark.parametrize('max_ctas', [32, 16, 64])
def test_recurrence_on_a_few_sms(max_ctas):
    """vog_lstm_set_max_ctas: the weight-streaming kernel with 64 / 64 (clamped: at most 64 units per CTA) / 32 hidden
    units per CTA (the SM-partitioned launch used next to a long visual branch) against the float64 host recurrence."""
    from vognet_pytorch_b200 import ops, _lib
    w, batch, mdl = _model('spat_gt5')
    mdl.set_compute('tf32')
    T, H = 20, 1024
    lens_list = [6, 20, 1, 13, 9]
    Bq = len(lens_list)
    lens = torch.tensor(lens_list, device=DEV)
    gx = (torch.rand(T * Bq, 8 * H, generator=torch.Generator().manual_seed(7)) - 0.5).to(DEV)
    _, _, whh = mdl._lang_weights(ops.LP_TF32)[0][0]
    L = _lib.lib()
    L.vog_lstm_set_max_ctas(max_ctas)
    try:
        out = ops.lstm_layer_fwd(gx, whh, lens, T, Bq, ops.LP_TF32).view(T, Bq, 2 * H)
        torch.cuda.synchronize()
    finally:
        L.vog_lstm_set_max_ctas(0)
    ref = _ref_recurrence(gx, whh, lens, T, Bq, H)
    assert (out.cpu().double() - ref).abs().max() < 1e-3
