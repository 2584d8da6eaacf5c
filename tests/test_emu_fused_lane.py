"""The lane-per-problem whole-solve kernel k_fused_lane (socp.jl_b200/csrc/fused_lane.cuh: tiny problems, BASELINE.json
C3) run on the SIMT emulator (tests/simt_emu/) and compared with the C oracle -- CPU-side coverage of the kernel's
logic: every lanes-per-warp variant, lanes that take several problems one after the other (work queue, cooperative
problem copy, the F / N slot alternation with lanes out of phase), a layout with a positive-orthant block, the
iteration cap, infeasible data.  The emulator is test infrastructure; the product path is the CUDA build."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "simt_emu"))
import emu  # noqa: E402
import socp_b200 as sb  # noqa: E402
from socp_b200 import generators as gen  # noqa: E402
from oracle import c_oracle as co  # noqa: E402

oc = lambda cones: tuple((c.kind, c.offs, c.dim) for c in cones)
rel = lambda a, b: np.abs(a - b) / np.maximum(1.0, np.abs(b))


def check(prob, res, max_iter=40, tol_obj=1e-8, tol_x=1e-6):
    B = prob.c.shape[0]
    ref = co.solve_batch(prob.c, prob.A_cm, prob.b, prob.G_cm, prob.h, oc(prob.cones), sing=np.zeros(B, dtype=np.uint8),
                         nthreads=4, max_iter=max_iter)
    assert (res["status"] == ref["status"]).all(), (res["status"], ref["status"])
    assert np.all(np.abs(res["iters"].astype(int) - ref["iters"].astype(int)) <= 1)
    same = res["iters"] == ref["iters"]
    assert same.sum() >= B - max(1, B // 10), (res["iters"], ref["iters"])
    conv = same & (ref["status"] == sb.STATUS_CONVERGED)
    cmp = conv if conv.any() else same          # iteration cap: compare the iterate the cap stopped at
    d = np.maximum(rel(res["pobj"][cmp], ref["pobj"][cmp]), rel(res["dobj"][cmp], ref["dobj"][cmp]))
    assert d.max() <= tol_obj, d.max()
    assert np.abs(res["x"][cmp] - ref["x"][cmp]).max() <= tol_x
    return ref


@pytest.mark.parametrize("lpw", [8, 16, 32])
def test_c3_vs_c_oracle(lpw):
    prob = gen.make_config("C3", batch=40)
    res = emu.solve_lane(prob.c, prob.G_cm, prob.h, oc(prob.cones), lpw=lpw)
    check(prob, res)


@pytest.mark.parametrize("lpw,order", [(8, 1), (16, 2)])
def test_c3_lanes_take_several_problems(lpw, order):
    """One CTA for 200 problems: every lane works through ~3 problems with different iteration counts, so lanes of a
    warp sit in different phases (initial point next to affine directions, idle N slots); any schedule, same bits."""
    prob = gen.make_config("C3", batch=200)
    res = emu.solve_lane(prob.c, prob.G_cm, prob.h, oc(prob.cones), lpw=lpw, grid_cap=1, order=order)
    check(prob, res)
    ref = emu.solve_lane(prob.c, prob.G_cm, prob.h, oc(prob.cones), lpw=32, grid_cap=0, order=0)
    for key in ("x", "z", "s", "pobj", "dobj", "iters", "status"):
        assert np.array_equal(res[key], ref[key]), key          # a problem's arithmetic does not depend on its lane


def test_mixed_layout_with_orthant_block():
    cones = [sb.POC(0, 5)] + [sb.SOC(5 + 3 * i, 3) for i in range(3)]
    prob = gen.random_feasible(70, 6, 0, cones, 0.3, 0, 11)
    res = emu.solve_lane(prob.c, prob.G_cm, prob.h, oc(prob.cones), lpw=8, grid_cap=1)
    # randomly generated mixed-cone family: the two oracles themselves differ by > 1e-8 on some of these problems
    # (error growth ~1e3 per iteration near the end, see tests/test_emu_fused3.py::check_vs_c_oracle)
    check(prob, res, tol_obj=1e-6, tol_x=1e-4)


def test_iteration_cap_and_unknown_layout():
    prob = gen.make_config("C3", batch=12)
    res = emu.solve_lane(prob.c, prob.G_cm, prob.h, oc(prob.cones), lpw=8, max_iter=3)
    assert (res["status"] == sb.STATUS_MAXITER).all() and (res["iters"] == 3).all()
    check(prob, res, max_iter=3)
    other = gen.random_feasible(4, 12, 0, gen.soc_cones(8, 5), 0.3, 0, 3)
    with pytest.raises(RuntimeError):
        emu.solve_lane(other.c, other.G_cm, other.h, oc(other.cones))


def test_rank_deficient_G_is_reported():
    """G with a zero column: the initial factorisation fails (the reference's cholesky! throws) -> NUMERICAL, zeros."""
    prob = gen.make_config("C3", batch=9)
    G = prob.G_cm.copy().reshape(9, 12, 40)
    G[4, 7, :] = 0.0
    res = emu.solve_lane(prob.c, G.reshape(9, -1), prob.h, oc(prob.cones), lpw=8)
    assert res["status"][4] == sb.STATUS_NUMERICAL and not res["x"][4].any()
    assert (np.delete(res["status"], 4) == sb.STATUS_CONVERGED).all()


def test_c3_deeper_ring_variant(monkeypatch):
    """SOCP_B200_LANE_RS2: two rows per ring stage (64 problems per SM on the device) -- same arithmetic, same bits."""
    prob = gen.make_config("C3", batch=24)
    ref = emu.solve_lane(prob.c, prob.G_cm, prob.h, oc(prob.cones), lpw=32)
    monkeypatch.setenv("SOCP_B200_LANE_RS2", "1")
    res = emu.solve_lane(prob.c, prob.G_cm, prob.h, oc(prob.cones), lpw=16)
    for key in ("x", "z", "s", "pobj", "dobj", "iters", "status"):
        assert np.array_equal(res[key], ref[key]), key


def test_cones_of_different_dimensions():
    """Two SOC(3) and one SOC(5) after two orthant rows: the cone loops run group by group (LaneDimsG); on the device such
    a layout is specialised at run time (csrc/lane_jit.cu), here the one instantiation of the emulator harness."""
    cones = [sb.POC(0, 2), sb.SOC(2, 3), sb.SOC(5, 3), sb.SOC(8, 5)]
    prob = gen.random_feasible(45, 6, 0, cones, 0.3, 0, 5)
    res = emu.solve_lane(prob.c, prob.G_cm, prob.h, oc(prob.cones), lpw=16, grid_cap=1)
    check(prob, res, tol_obj=1e-6, tol_x=1e-4)


def test_plan_family_and_specialisation_choice():
    """fl_plan: C3 has its compile-time instantiation at 96 problems per SM; other layouts of the family are marked for
    run-time specialisation with their cones grouped by dimension; layouts outside the family are declined."""
    c3 = [(1, 4 * j, 4) for j in range(10)]
    pl = emu.plan_lane(12, 0, c3)
    assert pl["fits"] and pl["shape"] == 1 and pl["pps"] == 96 and pl["smem"] == 96 * 300 * 8
    mixed = [(0, 0, 3)] + [(1, 3 + 3 * i, 3) for i in range(4)] + [(1, 15 + 5 * i, 5) for i in range(3)] + [(1, 30, 4)]
    pl = emu.plan_lane(9, 0, mixed)
    assert pl["fits"] and pl["shape"] == 100 and pl["groups"] == [(4, 3), (3, 5), (1, 4)] and pl["pps"] % 32 == 0
    assert pl["smem"] <= 227 * 1024
    assert not emu.plan_lane(9, 1, mixed)["fits"]                                        # equality rows
    assert not emu.plan_lane(9, 0, [(1, 0, 4), (0, 4, 3)])["fits"]                       # orthant rows after a cone
    assert not emu.plan_lane(20, 0, [(1, 4 * j, 4) for j in range(10)])["fits"]          # n > 16
    assert not emu.plan_lane(8, 0, [(1, 0, 12), (1, 12, 12)])["fits"]                    # cone dimension > 8
    assert not emu.plan_lane(8, 0, [(1, 1, 4), (1, 5, 4)])["fits"]                       # a gap before the first cone
