"""RPN_3D_loss_smp (m3dssd_b200/lib/loss/rpn_3d.py: static shapes, no host synchronisation) against the golden fixture
the UNMODIFIED reference class produced (tests/golden/make_golden_loss.py: loss, stats and gradients of
/root/reference/lib/loss/rpn_3d.py:659-1360 on seeded inputs)."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden_loss as G  # noqa: E402  (case definitions + seeded inputs; imports no reference code at import time)

GOLD = np.load(os.path.join(HERE, "golden", "loss_rpn_3d_smp.npz"))


def _run(name, device):
    from m3dssd_b200.lib.loss.rpn_3d import RPN_3D_loss_smp
    conf, tar, (cls, b2, b3) = G.case_inputs(name)
    crit = RPN_3D_loss_smp(conf).to(device)
    leaves = [t.clone().to(device).requires_grad_(True) for t in (cls, b2, b3)]
    prob = torch.softmax(leaves[0], dim=2)
    loss, stats = crit(leaves[0], prob, leaves[1], leaves[2], tar, torch.tensor(G.FEAT))
    loss.backward()
    return loss, stats, leaves


def _check(name, loss, stats, leaves):
    ref = float(GOLD[name + ".loss"])
    assert abs(float(loss) - ref) <= 2e-6 * abs(ref), (name, float(loss), ref)
    seen = set()
    for s in stats:
        key = "%s.stat.%s.%s" % (name, s["group"], s["name"])
        seen.add(key)
        assert key in GOLD.files, "stat %s is not one the reference reports" % key
        r = float(GOLD[key])
        assert abs(float(s["val"]) - r) <= 1e-5 * max(abs(r), 1e-3), (key, float(s["val"]), r)
    assert seen == {k for k in GOLD.files if k.startswith(name + ".stat.")}
    for k, t in zip(("cls", "bbox_2d", "bbox_3d"), leaves):
        g = GOLD["%s.grad.%s" % (name, k)]
        got = t.grad.detach().cpu().numpy()
        assert (g != 0).any()
        assert np.array_equal(got != 0, g != 0), "%s: gradient support of %s differs (sampling)" % (name, k)
        assert np.abs(got - g).max() <= 2e-6 * np.abs(g).max(), (name, k, np.abs(got - g).max(), np.abs(g).max())


@pytest.mark.parametrize("name", list(G.CASES))
def test_loss_matches_reference_golden_cpu(name):
    _check(name, *_run(name, "cpu"))


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(G.CASES))
def test_loss_matches_reference_golden_gpu(name):
    _check(name, *_run(name, "cuda"))


@pytest.mark.gpu
def test_loss_is_cuda_graph_capturable():
    """No host synchronisation inside: forward + backward of the loss replay as a CUDA graph and follow new inputs."""
    from m3dssd_b200.lib.loss.rpn_3d import RPN_3D_loss_smp
    conf, tar, (cls, b2, b3) = G.case_inputs("shipped")
    crit = RPN_3D_loss_smp(conf).cuda()
    tar = {k: ({kk: vv.cuda() for kk, vv in v.items()} if isinstance(v, dict) else v.cuda()) for k, v in tar.items()}
    st = [t.clone().cuda().requires_grad_(True) for t in (cls, b2, b3)]

    def step():
        for t in st:
            t.grad = None
        loss, _ = crit(st[0], torch.softmax(st[0], dim=2), st[1], st[2], tar)
        loss.backward()
        return loss
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = step()
    with torch.no_grad():
        for t, src in zip(st, G.case_inputs("invalid_image")[2]):  # other predictions, same targets
            t.copy_(src.cuda())
    g.replay()
    torch.cuda.synchronize()
    eager = [t.detach().clone().requires_grad_(True) for t in st]
    ref, _ = crit(eager[0], torch.softmax(eager[0], dim=2), eager[1], eager[2], tar)
    ref.backward()
    assert torch.allclose(out, ref, rtol=1e-6)
    for a, b in zip(st, eager):
        assert torch.allclose(a.grad, b.grad, rtol=1e-5, atol=1e-9)
