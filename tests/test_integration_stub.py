"""The reference-side ctypes stub printed in INTEGRATION.md section 3 is EXECUTED here, so it cannot go stale again:
  * CPU: its struct mirrors have the sizeof and field offsets gcc gives include/isls_b200.h; a stale struct is rejected
    by the library's ABI guard (struct_size) instead of being read out of bounds;
  * GPU: `ilqr_admm_b200` bound to a reference-shaped object solves 8 car problems and agrees with the oracle."""
import ctypes as C
import os
import re
import subprocess
import tempfile
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stub():
    import __graft_entry__ as G
    G.build()
    from isls_b200 import _lib
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", md, flags=re.S)
    code = [b for b in blocks if b.startswith("# isls/_b200.py")]
    assert len(code) == 1, "INTEGRATION.md must hold exactly one `# isls/_b200.py` block"
    os.environ["ISLS_B200_LIB"] = _lib.LIB_PATH
    ns = {}
    exec(compile(code[0], "INTEGRATION.md:stub", "exec"), ns)
    return ns


def test_stub_struct_layouts_match_the_header():
    ns = _stub()
    structs = {"isls_problem_desc": ns["_Desc"], "isls_solve_opts": ns["_Opts"], "isls_solve_out": ns["_Out"]}
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
        got = dict(ln.split() for ln in subprocess.check_output([exe], text=True).strip().splitlines())
    for cname, ct in structs.items():
        assert int(got[cname]) == C.sizeof(ct), "sizeof(%s): header %s, stub %d" % (cname, got[cname], C.sizeof(ct))
        for fname, _ in ct._fields_:
            assert int(got["%s.%s" % (cname, fname)]) == getattr(ct, fname).offset, (cname, fname)
        assert ct().struct_size == C.sizeof(ct)


def test_stale_binding_is_rejected_by_the_abi_guard():
    """A caller built against an older header (shorter struct) must get ISLS_E_INVALID, not an out-of-bounds read."""
    ns = _stub()
    L = ns["_L"]
    d = ns["_Desc"](model_id=1, n=4, m=2, N=100, n_via=1, L=20, dt=0.1, u_std=0.01)
    d.struct_size = C.sizeof(d) - 24               # what a stub without the last fields would say
    plan = C.c_void_p()
    assert L.isls_plan_create(C.byref(d), C.byref(plan)) == -1
    assert b"struct_size" in L.isls_last_error_string()
    d.struct_size = 0                              # a pre-guard binding: first field was model_id = 0
    assert L.isls_plan_create(C.byref(d), C.byref(plan)) == -1


@pytest.mark.gpu
def test_stub_solves_car_problems_like_the_oracle():
    ns = _stub()
    from oracle import problems as P, restated as R
    from isls_b200.solver import alphas
    p = P.car_batch(8, I_o=6, I_a=4)
    assert p["zs"].ndim == 2, "the stub broadcasts one via-point set [k, n] over the batch"
    ref_like = types.SimpleNamespace(x_dim=4, u_dim=2, N=p["N"], Qs=[np.diag(q) for q in p["Qdiag"]], seq=p["seq"],
                                     alphas=alphas(50), Rt=np.eye(2) * p["u_std"], zs=p["zs"])
    x, u, cl = ns["ilqr_admm_b200"](ref_like, "car", p["dt"], p["x0"], p["u0"] if p["u0"].ndim == 3 else
                                    np.broadcast_to(p["u0"], (8,) + p["u0"].shape[-2:]), p["lo_u"], p["hi_u"], p["rho_u"],
                                    p["I_o"], p["I_a"], p["L"], p["tol"])
    o = R.ilqr_admm(p)
    m = ~np.isnan(o["cost_log"])
    assert np.array_equal(~np.isnan(cl), m)
    assert np.max(np.abs(cl[m] - o["cost_log"][m]) / np.abs(o["cost_log"][m])) < 1e-9
    assert np.abs(u - o["u"]).max() < 1e-8 and np.abs(x - o["x"]).max() < 1e-8
