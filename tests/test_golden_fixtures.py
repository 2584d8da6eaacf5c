"""Committed golden fixtures (tests/golden/solve_golden.json, written by tests/golden/make_golden.py from the numpy
oracle): the C oracle must reproduce them on the CPU, the CUDA path (through the C ABI) on a B200."""
import json
import os

import numpy as np
import pytest

from socp_b200 import generators as gen
import refcases as rc

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "solve_golden.json")))
REF_ITERS = {"socp1": 5, "socp2": 12, "socp3": 10, "control": 4}


def test_fixture_pins():
    for name, it in REF_ITERS.items():
        g = GOLD["reference_instances"][name]
        assert g["status"] == 0 and g["iters"] == it
    assert abs(GOLD["reference_instances"]["control"]["pobj"] - 0.2901440) < 1e-6
    for name in ("socp1", "socp2", "socp3"):                              # test/runtests.jl:142,166,187
        assert np.linalg.norm(np.array(GOLD["reference_instances"][name]["x"]) - rc.ALL_C1[name]()["xstar"]) < 1e-3


def _check(res_status, res_iters, res_pobj, res_dobj, gold, otol):
    for q, g in enumerate(gold):
        assert int(res_status[q]) == g["status"], q
        assert abs(int(res_iters[q]) - g["iters"]) <= 1, q
        if int(res_iters[q]) == g["iters"]:
            assert abs(res_pobj[q] - g["pobj"]) <= otol * max(1.0, abs(g["pobj"])), q
            assert abs(res_dobj[q] - g["dobj"]) <= otol * max(1.0, abs(g["dobj"])), q


@pytest.mark.parametrize("cfg", ["C2", "C3"])
def test_c_oracle_reproduces_fixtures(cfg):
    from oracle import c_oracle as co
    gold = GOLD["seeded"][cfg]
    prob = gen.make_config(cfg, batch=len(gold))
    cones = tuple((c.kind, c.offs, c.dim) for c in prob.cones)
    r = co.solve_batch(prob.c, prob.A_cm, prob.b, prob.G_cm, prob.h, cones, sing=prob.sing, nthreads=2)
    _check(r["status"], r["iters"], r["pobj"], r["dobj"], gold, 1e-8)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", ["C2", "C3"])
def test_gpu_reproduces_fixtures(cfg):
    import socp_b200 as sb
    gold = GOLD["seeded"][cfg]
    prob = gen.make_config(cfg, batch=len(gold))
    res = sb.solve_socp_batch(prob, sb.SolverState(prob))
    assert res.timings["path_used"] == sb.PATH_FUSED
    _check(res.status, res.iters, res.pobj, res.dobj, gold, 1e-8)
    for q, g in enumerate(gold):
        if int(res.iters[q]) == g["iters"]:
            assert np.max(np.abs(res.x[q] - np.array(g["x"]))) < 1e-4      # late-iteration amplification, SURVEY.md 7.3


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["socp1", "socp2", "socp3", "control"])
def test_gpu_reference_instances_vs_fixtures(name):
    import socp_b200 as sb
    d = rc.ALL_C1[name]()
    g = GOLD["reference_instances"][name]
    cones = tuple(sb.POC(o, dm) if kd == 0 else sb.SOC(o, dm) for kd, o, dm in d["cones"])
    prob = sb.Problem(d["c"], d["A"], d["b"], d["G"], d["h"], cones)
    st = sb.solve_socp(prob, sb.SolverState(prob, sb.B200Solver(prob)))
    assert st.status == g["status"] and st.iters == g["iters"]
    otol = 1e-8 if name in ("socp1", "control") else 1e-6                  # see tests/test_gpu_parity.py
    assert abs(st.pobj - g["pobj"]) <= otol * max(1.0, abs(g["pobj"]))
    assert np.linalg.norm(st.x - np.array(g["x"])) <= 1e-3 * max(1.0, np.linalg.norm(g["x"]))
