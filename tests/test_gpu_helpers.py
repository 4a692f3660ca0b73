"""GPU parity of the stand-alone differentiable operators (diffpiso_b200/helpers.py) against the oracle, forward and
backward (torch.autograd vs oracle/adjoint.py)."""
import numpy as np
import pytest
import torch

from common import SMALL_SETUPS, random_fields, rel_l2
from oracle import adjoint as A
from oracle import oracle as O
from test_gpu_piso_step import DEV, build_sim, extrap

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["ldc8", "periodic16", "tml16x24", "sml16x48"])
def test_helper_operators_forward_and_backward(name):
    import diffpiso_b200 as dp
    s = SMALL_SETUPS[name]()
    sim = build_sim(s)
    ny, nx = s["ny"], s["nx"]
    nf, nc = ny * (nx + 1) + (ny + 1) * nx, ny * nx
    dxy = (s["dy"], s["dx"])
    rng = np.random.RandomState(3)
    per = (s["per_y"], s["per_x"])
    # finite_volume_gradient_tensor
    p = rng.randn(2, nc).astype(np.float32)
    tp = torch.as_tensor(p).to(DEV).requires_grad_(True)
    cg = dp.CenteredGrid(tp.reshape(2, ny, nx, 1), dx=dxy, extrapolation=extrap(s["pbc"]))
    g = dp.finite_volume_gradient_tensor(cg, sim)
    gflat = dp.flatten_staggered_data(g, coord_flip=True)
    w = rng.randn(2, nf).astype(np.float32)
    (gflat * torch.as_tensor(w).to(DEV)).sum().backward()
    for i in range(2):
        assert np.array_equal(gflat[i].detach().cpu().numpy(), O.fv_gradient(ny, nx, s["dy"], s["dx"], s["pbc"], s["access"], p[i]))
        assert rel_l2(tp.grad[i].cpu().numpy(), A.fv_gradient_adj(ny, nx, s["dy"], s["dx"], s["pbc"], s["access"], w[i])) < 1e-6
    # finite_volume_divergence
    vel = rng.randn(2, nf).astype(np.float32)
    tv = torch.as_tensor(vel).to(DEV).requires_grad_(True)
    sg = dp.StaggeredGrid(flat=tv, resolution=(ny, nx), dx=dxy)
    d = dp.finite_volume_divergence(sg, per)
    wc = rng.randn(2, nc).astype(np.float32)
    (d.reshape(2, nc) * torch.as_tensor(wc).to(DEV)).sum().backward()
    for i in range(2):
        assert np.array_equal(d[i].detach().cpu().numpy().ravel(), O.fv_divergence(ny, nx, s["dy"], s["dx"], vel[i]))
        assert np.array_equal(tv.grad[i].cpu().numpy(), A.fv_divergence_adj(ny, nx, s["per_x"], s["per_y"], s["dy"], s["dx"], wc[i]))
    # advection_matrix_cuda + explicit_H_csr
    v0 = np.stack([random_fields(s, 60 + i)[0] for i in range(2)])
    vel0 = dp.StaggeredGrid(flat=torch.as_tensor(v0).to(DEV), resolution=(ny, nx), dx=dxy)
    c = O.step_constants(s["dy"], s["dx"], s["dt"])
    visc = torch.as_tensor(np.atleast_1d(s["visc"])).to(DEV)
    values, rp, ci, a_st, nnz, a_flat = dp.advection_matrix_cuda(vel0, sim, visc, c["beta"])
    orp, oci = O.csr_structure(ny, nx, s["per_x"], s["per_y"])
    assert np.array_equal(rp.cpu().numpy(), orp) and np.array_equal(ci.cpu().numpy(), oci)
    dvec = rng.randn(2, nf).astype(np.float32)
    td = torch.as_tensor(dvec).to(DEV).requires_grad_(True)
    h = dp.explicit_H_csr(values, rp, ci, dp.StaggeredGrid(flat=td, resolution=(ny, nx), dx=dxy), (2, ny + 1, nx + 1, 2),
                          a_st, c["beta"], per)
    hflat = dp.flatten_staggered_data(h, coord_flip=True)
    (hflat * torch.as_tensor(w).to(DEV)).sum().backward()
    n_u, n_v, z_u, z_v = O.sizes(ny, nx, s["per_x"], s["per_y"])
    for i in range(2):
        vi, ai = values[i].cpu().numpy(), a_flat[i].cpu().numpy()
        ref = np.concatenate([O.h_apply(orp[:n_u + 1], oci[:z_u], vi[:z_u], ai[:n_u], c["beta"], dvec[i][:n_u]),
                              O.h_apply(orp[n_u + 1:], oci[z_u:], vi[z_u:], ai[n_u:], c["beta"], dvec[i][n_u:])])
        assert np.array_equal(hflat[i].detach().cpu().numpy(), ref)
        assert rel_l2(td.grad[i].cpu().numpy(), A.h_apply_adj(s, vi, ai, c["beta"], w[i])) < 1e-6
