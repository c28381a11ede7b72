"""End-to-end serving pipeline of bench.py's `e2e` leg: packed pinned batches -> BatchPrefetcher (copy stream, ring of
device buffers) -> captured forward + selection -> PredictionFetcher (side-stream device->host read, one copy for the
three selection results).  With two steps in flight and DIFFERENT batches every step, every step's host-side
predictions must equal what a synchronous call on that batch returns."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import vognet_pytorch_b200 as vb              # noqa: E402
from vognet_pytorch_b200 import synth, runtime          # noqa: E402

DEV = 'cuda:0'


def _ctx(name, mode):
    w, batch = synth.workload(name)
    cfg, comm = synth.default_cfg(w['conc_type']), synth.default_comm(w['nppf'])
    sel = vb.get_mdl_loss_eval(cfg)
    mdl = sel['mdl'](cfg, comm)
    mdl.load_state_dict(synth.make_state_dict(), strict=True)
    mdl = mdl.to(DEV).eval()
    mdl.set_compute(mode)
    return w, batch, mdl, sel['eval'](cfg, comm, DEV)


def test_selection_results_share_one_allocation():
    """ops.select_fwd returns boxes / scores / indexs as views of one buffer (what lets them travel as one copy)"""
    w, batch, mdl, ev = _ctx('cpu_ref', 'tf32')
    db = synth.clone_batch(batch, DEV)
    with torch.no_grad():
        s = ev.get_out_results_boxes(mdl(db), db)
    ptrs = {t.untyped_storage().data_ptr() for t in (s['boxes'], s['scores'], s['indexs'])}
    assert len(ptrs) == 1 and all(t.is_contiguous() for t in s.values())
    assert s['indexs'].dtype == torch.int64 and s['indexs'].data_ptr() % 8 == 0


@pytest.mark.parametrize('name,mode', [('spat_gt5', 'tf32'), ('temp_gt5', 'tf32')])
def test_two_steps_in_flight_return_each_steps_own_predictions(name, mode):
    w, batch, mdl, ev = _ctx(name, mode)
    n = 7
    variants = []
    for i in range(n):                                        # a different batch every step
        b = {k: v.clone() for k, v in batch.items()}
        b['pad_region_feature'] = b['pad_region_feature'] * (1.0 + 0.07 * i)
        b['seg_feature_for_frms'] = b['seg_feature_for_frms'].roll(i, 0)
        variants.append(b)
    want = []
    with torch.no_grad():
        for b in variants:                                    # synchronous reference: plain device batch
            db = synth.clone_batch(b, DEV)
            s = ev.get_out_results_boxes(mdl(db), db)
            torch.cuda.synchronize()
            want.append([s[k].cpu() for k in ('boxes', 'scores', 'indexs')])
    keys = tuple(k for k in mdl._GRAPH_KEYS if k in batch)
    packed = [runtime.pack_host_batch(b, first=keys) for b in variants]
    pre = runtime.BatchPrefetcher(iter(packed), DEV)
    fetch = runtime.PredictionFetcher(DEV)
    got = [None] * n

    def launch(i):
        b = pre.next()
        with torch.no_grad():
            s = ev.get_out_results_boxes(mdl(b), b)
        pre.release(b, fetch.fetch(i, (s['boxes'], s['scores'], s['indexs'])))

    for i in range(n):
        launch(i)
        if i:
            got[i - 1] = [t.clone() for t in fetch.get(i - 1)]
    got[n - 1] = [t.clone() for t in fetch.get(n - 1)]
    assert fetch.nbytes == sum(t.numel() * t.element_size() for t in want[0])
    for i in range(n):
        for a, b_ in zip(got[i], want[i]):
            assert a.dtype == b_.dtype and a.shape == b_.shape
            assert torch.equal(a, b_), f'step {i}'
    # and the steps really differ from each other (the check above is not vacuous)
    assert any(not torch.equal(want[0][1], want[i][1]) for i in range(1, n))


def test_fetcher_with_unrelated_tensors_falls_back_to_one_copy_each():
    f = runtime.PredictionFetcher(DEV)
    a = torch.arange(12, device=DEV, dtype=torch.float32).view(3, 4)
    b = torch.arange(5, device=DEV, dtype=torch.int64)
    f.fetch(0, (a, b))
    ha, hb = f.get(0)
    assert torch.equal(ha, a.cpu()) and torch.equal(hb, b.cpu()) and f.nbytes == 12 * 4 + 5 * 8
