"""CPU: the C-ABI library loads and exports every symbol include/isls_b200.h declares; host-side argument
handling that needs no GPU (no compute calls)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    import __graft_entry__ as G
    G.build()
    from isls_b200 import _lib
    return _lib


def test_header_symbols_exported():
    L = _lib()
    hdr = open(os.path.join(ROOT, "include", "isls_b200.h")).read()
    names = set(re.findall(r"\b(isls_[a-z0-9_]+)\s*\(", hdr))
    assert names, "no declarations found"
    lib = L.lib()
    for nm in sorted(names):
        assert hasattr(lib, nm), "header declares %s but the library does not export it" % nm
    assert names == set(L.EXPORTS), (names ^ set(L.EXPORTS))
    assert lib.isls_version() == 200


def test_model_registry_and_errors():
    L = _lib()
    lib = L.lib()
    assert lib.isls_model_id(b"double_integrator") == 0
    assert lib.isls_model_id(b"car") == 1
    assert lib.isls_model_id(b"arm3") == 2
    assert lib.isls_model_id(b"unicycle") == -2
    assert b"unicycle" in lib.isls_last_error_string()
    assert lib.isls_model_supported(1, 4, 2) == 0
    assert lib.isls_model_supported(1, 5, 2) != 0
    assert lib.isls_model_supported(0, 4, 2) == 0 and lib.isls_model_supported(0, 2, 1) == 0
    assert lib.isls_model_supported(2, 9, 3) == 0


def test_host_mirror_argument_checks():
    """The host mirror refuses what the device path cannot do, loudly (no CPU fallback)."""
    _lib()
    from isls_b200 import iSLS, Bound
    from isls_b200.utils import diag_of
    s = iSLS(4, 2, 100)
    with pytest.raises(TypeError):
        s.forward_model = lambda x, u: x
    s.forward_model = ("car", {"dt": 0.1})
    with pytest.raises(TypeError):
        s._check_get_AB(lambda x, u: None)
    s._check_get_AB("car")
    with pytest.raises(NotImplementedError):
        diag_of(np.ones((2, 2)), "Q")
    Qr, Rr = s.compute_Rr_Qr(None, 10.0)
    assert Qr is None and Rr.shape == (100, 2) and np.all(Rr == 10.0)
    Qr, _ = s.compute_Rr_Qr(np.diag([0.0, 0, 1, 1]), None)
    assert Qr.shape == (100, 4) and np.all(Qr[:, 2:] == 1) and np.all(Qr[:, :2] == 0)
    lo, hi = Bound(-0.5, 0.5).expand(100, 2)
    assert lo.shape == (100, 2) and np.all(hi == 0.5)
    assert np.array_equal(Bound(-1, 1)(np.array([-3.0, 0.2, 5.0])), [-1, 0.2, 1])
    # legacy spellings exist (README.md:24-39)
    assert s.set_cost_variables.__func__ is s.set_quadratic_cost.__func__
    assert hasattr(s, "solve_ilqr")
    with pytest.raises(NotImplementedError):                # the dense batch form is not built: loud, not a silent DP run
        s.solve_ilqr("car", dp=False)


def test_ctypes_struct_layouts_match_the_header():
    """The ctypes mirrors of the POD structs (isls_b200/_lib.py) must have the size and field offsets gcc gives the
    declarations of include/isls_b200.h - a silent mismatch would shift every argument."""
    import ctypes as C
    import subprocess
    import tempfile
    L = _lib()
    structs = {"isls_problem_desc": L.ProblemDesc, "isls_solve_opts": L.SolveOpts, "isls_solve_out": L.SolveOut,
               "isls_sls_admm_opts": L.SlsAdmmOpts, "isls_proj_params": L.ProjParams,
               "isls_proj_set_entry": L.ProjSetEntry, "isls_proj_set_params": L.ProjSetParams}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "isls_b200.h"', 'int main(void) {']
    for cname, ct in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in ct._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ['return 0;', '}']
    with tempfile.TemporaryDirectory() as td:
        src, exe = os.path.join(td, "layout.c"), os.path.join(td, "layout")
        open(src, "w").write("\n".join(lines))
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, src])
        out = subprocess.check_output([exe], text=True)
    got = dict(ln.split() for ln in out.strip().splitlines())
    for cname, ct in structs.items():
        assert int(got[cname]) == C.sizeof(ct), "sizeof(%s): header %s, ctypes %d" % (cname, got[cname], C.sizeof(ct))
        for fname, _ in ct._fields_:
            assert int(got["%s.%s" % (cname, fname)]) == getattr(ct, fname).offset, (cname, fname)
