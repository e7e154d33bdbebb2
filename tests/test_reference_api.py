"""Host-side mirror of the reference scripts (eigensolver_b200/reference_api.py): the preset table
covers every solver script the reference ships, with that script's own settings."""
import numpy as np
import pytest

import eigensolver_b200 as esb
from eigensolver_b200.reference_api import SCRIPTS

#: every solver script under /root/reference (SURVEY of the tree: */[Ss]olver*/*.py)
REFERENCE_SOLVER_SCRIPTS = {
    "Cylinder/Non-uniform density/Coronal/solvers/Density_cylinder.py",
    "Cylinder/Non-uniform density/Photospheric/Solvers/Density_cylinder_photospheric.py",
    "Cylinder/Non-uniform flow/Coronal/solvers/Cylinder_method_flow_testing.py",
    "Cylinder/Rotational flow/Photospheric/Solvers/Twisted_photospheric_flow_sausage.py",
    "Cylinder/Rotational flow/Photospheric/Solvers/Twisted_photospheric_flow_sausage_slow.py",
    "Cylinder/Rotational flow/Photospheric/Solvers/Twisted_photospheric_nonlinear_flow_kink_fast.py",
    "Cylinder/Rotational flow/Photospheric/Solvers/Twisted_photospheric_nonlinear_flow_kink_slow.py",
    "Slab/Non uniform density/Coronal/Solvers/multiprocessor_Inhomogeneous_method_coronal.py",
    "Slab/Non uniform density/Photospheric/Solvers/multiprocessor_Inhomogeneous_method.py",
    "Slab/Non uniform flow/Solver/flow_multiprocessor.py",
    "Slab/Non uniform flow/Solver/flow_multiprocessor_coronal.py",
}


def test_every_reference_solver_script_has_a_preset():
    assert {s["ref"] for s in SCRIPTS.values()} == REFERENCE_SOLVER_SCRIPTS


@pytest.mark.parametrize("name", list(SCRIPTS))
def test_preset_is_well_formed(name):
    sp = SCRIPTS[name]
    speeds = sp["speeds"](sp["medium"], sp["profile"])
    assert len(speeds) >= 2 and speeds == sorted(speeds) and np.all(np.isfinite(speeds))
    lo, hi, n = sp["wavenumber"]
    assert 0 < lo < hi and n >= 2 and sp["n_freq"] >= 2 and sp["tol"] > 0
    x = np.array([1.0, 0.5, 0.3]) if sp["kind"] == "cylinder_rotation" else np.array([-1.0, -0.5, -0.3])
    fields = sp["profile"](sp["medium"], x)
    # value, first and second derivative (rotation: v_phi, v_phi', c_i^2)
    assert len(fields) == 3
    flat = getattr(sp["profile"], "width", 1.0) > 100.0     # a uniform profile: differences are rounding
    if sp["kind"] != "cylinder_rotation" and not flat:      # the derivatives are the derivatives (central differences)
        h = 1e-5
        up, dn = sp["profile"](sp["medium"], x + h), sp["profile"](sp["medium"], x - h)
        scale = max(np.abs(fields[0]).max(), 1e-30)
        assert np.allclose((up[0] - dn[0]) / (2 * h), fields[1], atol=1e-8 * scale / h * h + 1e-7 * scale)
        assert np.allclose((up[1] - dn[1]) / (2 * h), fields[2], atol=1e-7 * max(np.abs(fields[1]).max(), scale))


def test_shipped_values():
    """Spot checks against the assignment lines of the scripts (file:line in reference_api.SCRIPTS)."""
    assert SCRIPTS["cylinder_density"]["profile"].width == 0.95            # Density_cylinder.py:125
    assert SCRIPTS["cylinder_flow"]["tol"] == 6.0                          # Cylinder_method_flow_testing.py:530
    assert SCRIPTS["rotation_kink_slow"]["accept"] == "ext"                # ..._kink_slow.py:586
    assert SCRIPTS["slab_density_photospheric"]["solver"]["ext_wavelengths"] == 7.0
    md = SCRIPTS["slab_flow_photospheric"]["medium"]
    assert (md.vA_e, md.U_e, md.U_i0) == (0.0, -0.15, 0.0)                 # flow_multiprocessor.py:65-69
    # the slab density script's speeds list contains cT at the slab boundary (:185, :202)
    sp = SCRIPTS["slab_density"]
    assert any(1.05 < v < 1.15 for v in sp["speeds"](sp["medium"], sp["profile"]))


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(SCRIPTS))
def test_worker_signature_on_gpu(name):
    """sausage()/kink() of every script: the reference's signature, queues receive equal-length lists,
    every reported (k, omega) passes that script's own acceptance test when re-evaluated."""
    sp = SCRIPTS[name]

    class Q:
        def __init__(self):
            self.items = []

        def put(self, x):
            self.items.append(x)

    with esb.ReferenceScript(name) as script:
        speeds = script.default_speeds()
        k = float(np.mean(sp["wavenumber"][:2]))
        found = 0
        for fn in (script.sausage, script.kink):
            for i in range(len(speeds) - 1):
                freq = np.linspace(speeds[i] * k, speeds[i + 1] * k, 4 * sp["n_freq"])
                ws, ks = Q(), Q()
                fn(k, ws, ks, freq)
                assert len(ws.items) == 1 and len(ks.items) == 1 and len(ws.items[0]) == len(ks.items[0])
                assert all(kk == k for kk in ks.items[0])
                if ws.items[0]:
                    mode = 0 if fn == script.sausage else 1
                    w = np.array(ws.items[0])
                    e, q = script.solver.dispersion_grid(mode, [k], w, layout="shared")
                    den = np.abs(e) if script.accept == "ext" else np.maximum(np.abs(e), np.abs(q))
                    assert np.all(np.abs(e - q) * 100 / den < script.tol)
                    assert np.all((w >= freq[0]) & (w <= freq[-1]))
                    found += len(w)
    # the driver form: same format as the pickle
    with esb.ReferenceScript(name) as script:
        out = script.run(wavenumber=np.linspace(*sp["wavenumber"][:2], 6), n_freq=40)
        assert len(out) == 4 and len(out[0]) == len(out[1]) and len(out[2]) == len(out[3])


def test_root_table_files_have_the_scripts_layout(tmp_path):
    """The output file of every solver script: four arrays for the density / flow scripts
    (Density_cylinder.py:1183), two for the rotational ones (Twisted_photospheric_*:782-790), as numpy arrays -
    what the reference's analysis scripts unpack (analysis_compare_coronal_eigenfunctions_coronal.py:364)."""
    import pickle
    from eigensolver_b200 import reference_api as ra
    res = [np.array([3.0, 3.1]), np.array([1.0, 1.0]), np.array([2.5]), np.array([0.7])]
    for script in SCRIPTS:
        path = tmp_path / (script + ".pickle")
        ra.write_root_table(path, res, script)
        with open(path, "rb") as fh:
            raw = pickle.load(fh)
        layout = ra.pickle_layout(script)
        assert len(raw) == (2 if script.startswith("rotation_") else 4) == len(layout)
        assert all(isinstance(a, np.ndarray) and a.dtype == np.float64 for a in raw)
        back = ra.read_root_table(path, script)
        for i in range(4):
            assert np.array_equal(back[i], res[i] if i in layout else np.zeros(0))
    if len(raw) == 4:                     # the unpacking line of the analysis scripts
        sol_omegas, sol_ks, sol_omegas_kink, sol_ks_kink = raw
    with pytest.raises(KeyError):
        ra.pickle_layout("no_such_script")
    with pytest.raises(ValueError):
        ra.read_root_table(path, "rotation_kink" if len(raw) == 4 else "cylinder_density")
