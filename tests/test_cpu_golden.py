"""CPU suite: the oracle against golden vectors produced by the REFERENCE'S OWN CUDA kernels on a B200
(tests/golden/README.md).  This is what pins the oracle for the assembly, Laplace-matrix and pressure-CG rows of the
hot path; the BiCGStab+ILU0 row has no reference-generated vectors (its launcher needs cuSPARSE entry points that CUDA
12.9 no longer ships) and is pinned only by textbook-determined algorithms checked in test_cpu_structure.py."""
import glob
import os

import numpy as np
import pytest

from common import SMALL_SETUPS, cg_iteration_slack, rel_l2
from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_kernels")
ASSEMBLE = sorted(glob.glob(os.path.join(GOLD, "assemble_*.npz")))
PRESSURE = sorted(glob.glob(os.path.join(GOLD, "pressure_*.npz")))


def test_golden_files_present():
    assert len(ASSEMBLE) >= 7 and len(PRESSURE) >= 14


@pytest.mark.parametrize("path", ASSEMBLE, ids=[os.path.basename(p)[:-4] for p in ASSEMBLE])
def test_oracle_assembly_equals_reference_golden(path):
    """CSR row_ptr / col_ind, matrix values and diagonal: bit-exact with the reference kernels' output."""
    name = os.path.basename(path)[len("assemble_"):-4]
    s = SMALL_SETUPS[name]()
    g = np.load(path)
    ny, nx = s["ny"], s["nx"]
    n_u = ny * (nx + 1)
    rp, ci = O.csr_structure(ny, nx, s["per_x"], s["per_y"])
    assert np.array_equal(rp, g["row_ptr"]) and np.array_equal(ci, g["col_ind"])
    vel = g["vel"]
    up, vp = O.pad_velocity(ny, nx, s["per_x"], s["per_y"], vel[:n_u].reshape(ny, nx + 1), vel[n_u:].reshape(ny + 1, nx))
    areas = (float(g["cell_area"][0]), float(g["cell_area"][1]))
    assert O.cell_areas(s["dy"], s["dx"]) == areas
    assert float(g["grid_spacing"][0]) == float(np.float32(s["dx"])) and float(g["grid_spacing"][1]) == float(np.float32(s["dy"]))
    values, a_diag = O.assemble(ny, nx, s["per_x"], s["per_y"], s["dy"], s["dx"], float(g["beta"]), up, vp, s["dirichlet"],
                                s["active"], s["noslip"], s["visc"], rp, areas=areas)
    assert np.array_equal(values, g["values"])
    assert np.array_equal(a_diag, g["a_diag"])
    assert float(g["beta"]) == O.step_constants(s["dy"], s["dx"], s["dt"])["beta"]


@pytest.mark.parametrize("path", PRESSURE, ids=[os.path.basename(p)[:-4] for p in PRESSURE])
def test_oracle_pressure_solve_equals_reference_golden(path):
    """Laplace matrix bit-exact; CG iteration count within the slack of common.cg_iteration_slack and solution within
    tol-scaled bounds of the reference kernels' (cuBLAS reductions associate differently)."""
    base = os.path.basename(path)[len("pressure_"):-4]
    name, prec = base.rsplit("_", 1)
    s = SMALL_SETUPS[name]()
    g = np.load(path)
    T = np.float64 if prec == "f64" else np.float32
    ny, nx = s["ny"], s["nx"]
    lap = O.laplace(ny, nx, s["active"], s["access"], g["k_vu"], T)
    assert np.array_equal(lap, g["lap"])
    tol = float(g["tol"])
    x, it = O.pressure_cg(ny, nx, s["per_x"], s["per_y"], lap, g["div"].astype(T), tol, s["cg_max_it"], s["cg_reset"],
                          s["rank_deficient"])
    if prec == "f64":
        assert abs(it - int(g["iterations"])) <= cg_iteration_slack(s, it), (it, int(g["iterations"]))
        assert rel_l2(x, g["x"]) < max(2e-5, 1000 * tol)
    else:
        assert rel_l2(x, g["x"]) < 5e-2
