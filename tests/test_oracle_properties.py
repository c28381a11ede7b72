"""Property tests of the CPU oracle against brute-force scalar restatements (small random cases; hypothesis drives the
shapes and seeds): the oracle is what the CUDA kernels are held to, so its vectorised index arithmetic gets its own
independent check."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from oracle import vog_oracle as vo


@settings(max_examples=25, deadline=None)
@given(st.integers(1, 3), st.integers(1, 4), st.integers(1, 4), st.integers(1, 6), st.sampled_from(['spat', 'temp']),
       st.integers(0, 10 ** 6))
def test_select_boxes_matches_loops(B, nsrl, ncmp, nppf, conc, seed):
    g = torch.Generator().manual_seed(seed)
    nfrm = 10
    P = ncmp * nfrm * nppf
    # a coarse grid of values so that ties occur (lowest index must win)
    scores = torch.randint(0, 4, (B, 1, nsrl, P), generator=g).float() / 4
    props = torch.rand(B, P, 7, generator=g)
    out = vo.select_boxes(scores, props, conc, ncmp, nppf)
    for b in range(B):
        for s in range(nsrl):
            for f in range(nfrm):
                best_v, best_vid = None, 0
                for v in range(ncmp):
                    base = (f * ncmp + v) * nppf if conc == 'spat' else (v * nfrm + f) * nppf
                    grp = scores[b, 0, s, base:base + nppf]
                    i = int(np.argmax(grp.numpy()))                       # first maximum
                    assert out['scores'][b, s, v, f] == grp[i]
                    assert torch.equal(out['boxes'][b, s, v, f], props[b, base + i])
                    if best_v is None or grp[i] > best_v:
                        best_v, best_vid = grp[i], v
                assert int(out['indexs'][b, s, f]) == (best_vid if conc == 'spat' else 0)


@settings(max_examples=25, deadline=None)
@given(st.integers(1, 2), st.integers(1, 7), st.integers(1, 5), st.integers(0, 10 ** 6))
def test_bbox_overlaps_matches_scalar_iou(B, N, K, seed):
    g = torch.Generator().manual_seed(seed)
    def boxes(n):
        x1 = torch.rand(B, n, generator=g) * 100
        y1 = torch.rand(B, n, generator=g) * 100
        return torch.stack([x1, y1, x1 + torch.rand(B, n, generator=g) * 50, y1 + torch.rand(B, n, generator=g) * 50,
                            torch.zeros(B, n)], -1)
    a, q = boxes(N), boxes(K)
    a[:, 0, 2:4] = a[:, 0, 0:2]                    # a zero-area anchor: overlap -1 (utils/box_utils.py:114-116)
    q[:, -1, 2:4] = q[:, -1, 0:2]                  # a zero-area gt box: overlap 0 (:112-113)
    msk = (torch.rand(B, N, K, generator=g) > 0.3).to(torch.uint8)
    ov = vo.bbox_overlaps_batch(a, q, msk)
    for b in range(B):
        for i in range(N):
            for k in range(K):
                ax1, ay1, ax2, ay2 = [float(v) for v in a[b, i, :4]]
                gx1, gy1, gx2, gy2 = [float(v) for v in q[b, k, :4]]
                iw = max(min(ax2, gx2) - max(ax1, gx1) + 1, 0.0)
                ih = max(min(ay2, gy2) - max(ay1, gy1) + 1, 0.0)
                ua = (ax2 - ax1 + 1) * (ay2 - ay1 + 1) + (gx2 - gx1 + 1) * (gy2 - gy1 + 1) - iw * ih
                want = iw * ih / ua * float(msk[b, i, k])
                if gx2 - gx1 + 1 == 1 and gy2 - gy1 + 1 == 1:
                    want = 0.0
                if ax2 - ax1 + 1 == 1 and ay2 - ay1 + 1 == 1:
                    want = -1.0
                assert abs(float(ov[b, i, k]) - want) < 1e-5


@settings(max_examples=20, deadline=None)
@given(st.integers(1, 3), st.integers(1, 4), st.integers(1, 5), st.integers(0, 10 ** 6))
def test_concat_videos_is_a_permutation_with_shifts(B, ncmp, nppf, seed):
    g = torch.Generator().manual_seed(seed)
    nfrm, D = 10, 8
    feat = torch.rand(B, ncmp, nfrm * nppf, D, generator=g)
    seg = torch.rand(B, ncmp, nfrm, 12, generator=g)
    props = torch.rand(B, ncmp, nfrm * nppf, 7, generator=g) * 100
    f, s, p = vo.concat_videos(feat, seg, props, 'spat', nfrm, nppf)
    for v in range(ncmp):
        for fr in range(nfrm):
            rows = slice((fr * ncmp + v) * nppf, (fr * ncmp + v + 1) * nppf)
            assert torch.equal(f[:, rows], feat[:, v, fr * nppf:(fr + 1) * nppf])
            assert torch.equal(s[:, fr * ncmp + v], seg[:, v, fr])
            want = props[:, v, fr * nppf:(fr + 1) * nppf].clone()
            want[..., 0] += 720.0 * v
            want[..., 2] += 720.0 * v
            assert torch.equal(p[:, rows], want)
    f2, s2, p2 = vo.concat_videos(feat, seg, props, 'temp', nfrm, nppf)
    assert torch.equal(f2, feat.reshape(B, -1, D)) and torch.equal(s2, seg.reshape(B, -1, 12))
    want = props.clone()
    want[..., 4] += 10.0 * torch.arange(ncmp).view(1, ncmp, 1)
    assert torch.equal(p2, want.reshape(B, -1, 7))
