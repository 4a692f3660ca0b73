"""`diffpiso_b200.SampleGroups`: sample groups of a batch as independent pipelines on their own CUDA streams (optionally
one CUDA graph per group) reproduce the single-stream batch bit for bit -- state after several fed-back steps, and the
input gradients of every step."""
import numpy as np
import pytest
import torch

from common import ALL_SETUPS, random_fields
from test_gpu_piso_step import build_sim

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _case(name, batch):
    import diffpiso_b200 as dp
    s = ALL_SETUPS[name]()
    ny, nx = s["ny"], s["nx"]
    nf, nc = ny * (nx + 1) + (ny + 1) * nx, ny * nx
    sim = build_sim(s)
    fields = [random_fields(s, 300 + i) for i in range(batch)]
    vel = torch.as_tensor(np.stack([f[0] for f in fields])).to(DEV)
    pres = torch.as_tensor(np.stack([f[1] for f in fields])).to(DEV)
    rng = np.random.RandomState(11)
    w_u = torch.as_tensor(rng.randn(batch, nf).astype(np.float32)).to(DEV)
    w_p = rng.randn(batch, nc).astype(np.float32)
    w_p = torch.as_tensor(w_p - w_p.mean(axis=1, keepdims=True)).to(DEV)
    dvals = torch.as_tensor(s["dirichlet_values"])[None].to(DEV)
    dxy = (s["dy"], s["dx"])

    def fn(v, p, wu, wp):
        nb = v.shape[0]
        v = v.detach().requires_grad_(True)
        p = p.detach().requires_grad_(True)
        velocity = dp.StaggeredGrid(flat=v, resolution=(ny, nx), dx=dxy, extrapolation="periodic")
        pressure = dp.CenteredGrid(p.reshape(nb, ny, nx, 1), dx=dxy, extrapolation="periodic")
        vn, pn, _ = dp.piso_step(velocity, pressure, pressure, pressure, s["dt"], sim, dvals)
        loss = (vn.flat * wu).sum() + (pn.data.reshape(nb, nc) * wp).sum()
        gv, gp = torch.autograd.grad(loss, (v, p))
        return vn.flat.detach(), pn.data.reshape(nb, nc).detach(), gv, gp
    return dp, fn, vel, pres, w_u, w_p


@pytest.mark.parametrize("groups,graph", [(3, False), (4, True), (10, True)])
def test_sample_groups_match_the_single_stream_batch(groups, graph):
    dp, fn, vel, pres, w_u, w_p = _case("periodic32", 10)
    v, p = vel, pres
    want = []
    for _ in range(3):
        v, p, gv, gp = fn(v, p, w_u, w_p)
        want.append((v.clone(), p.clone(), gv.clone(), gp.clone()))
    runner = dp.SampleGroups(fn, (vel, pres, w_u, w_p), groups=groups, graph=graph)
    assert runner.groups == groups and sum(c for _, c in runner.bounds) == 10
    for k in range(3):
        runner.step(feedback={0: 0, 1: 1})
        got = [runner.gather(j).clone() for j in range(4)]
        torch.cuda.synchronize()
        for g_, w_ in zip(got, want[k]):
            assert g_.shape == w_.shape and torch.equal(g_, w_)


def test_sample_groups_host_buffers_round_trip():
    """load() from / fetch() into pinned host buffers (the end-to-end loop of bench.py): two steps through the host equal
    two device-resident steps."""
    dp, fn, vel, pres, w_u, w_p = _case("periodic32", 6)
    v, p = vel, pres
    for _ in range(2):
        v, p, gv, gp = fn(v, p, w_u, w_p)
    runner = dp.SampleGroups(fn, (vel, pres, w_u, w_p), groups=2, graph=True)
    h_in = [[t.cpu().pin_memory() for t in runner.inputs(i)[:2]] for i in range(2)]
    runner.step()
    h_out = [[torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in runner.outputs(i)] for i in range(2)]
    for k in range(2):
        for i in range(2):
            runner.sync(i)
            runner.load(i, h_in[i][0], h_in[i][1], None, None)
            runner.launch(i)
            runner.fetch(i, *h_out[i])
            h_in[i][0], h_out[i][0] = h_out[i][0], h_in[i][0]
            h_in[i][1], h_out[i][1] = h_out[i][1], h_in[i][1]
    for i in range(2):
        runner.sync(i)
    got_v = torch.cat([h_in[i][0] for i in range(2)])          # after the swap the newest state sits in h_in
    got_gv = torch.cat([h_out[i][2] for i in range(2)])
    assert torch.equal(got_v, v.cpu()) and torch.equal(got_gv, gv.cpu())


def test_sample_groups_reject_bad_arguments():
    dp, fn, vel, pres, w_u, w_p = _case("periodic16", 2)
    with pytest.raises(ValueError):
        dp.SampleGroups(fn, (), groups=2)
    with pytest.raises(ValueError):
        dp.SampleGroups(fn, (vel, pres[:1]), groups=2)
    runner = dp.SampleGroups(fn, (vel, pres, w_u, w_p), groups=5, graph=False)
    assert runner.groups == 2
    with pytest.raises(RuntimeError):
        runner.outputs(0)
