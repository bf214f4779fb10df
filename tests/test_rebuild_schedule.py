"""Library-side rebuild schedule (callers that pass ago < 0): `Neighbor::decide` + `check_distance`
(src/neighbor.cpp:1923-2001) against `neighbor->ago` dumped from the reference binary on a hot 256-atom fluid
(tests/golden/ago_*.npz, generator oracle/make_golden_ago.py): rebuilds at irregular steps.

  * CPU: a numpy restatement of the schedule reproduces the reference's rebuild steps (pins the rule the device follows);
  * GPU (through the C ABI): the library rebuilds on exactly those steps, and its dipoles / forces at the last step equal
    the reference's (1e-10), i.e. the lists it kept in between were the right ones.
"""
import re

import numpy as np
import pytest

import polhelpers as H

CASES = ["ago_every1", "ago_delay4_every2", "ago_nocheck"]


def schedule(fx):
    w = str(fx["neigh_modify"]).split()
    kv = {w[i]: w[i + 1] for i in range(1, len(w), 2)}
    return int(kv["delay"]), int(kv["every"]), kv["check"] == "yes"


def restated_rebuild_steps(fx):
    """src/neighbor.cpp:1923-1937 (decide) and :1989-2001 (check_distance: any owned atom moved more than skin/2 since the
    last build; positions are compared as stored, the reference wraps them at the rebuild itself)"""
    delay, every, check = schedule(fx)
    x, trig = fx["x"], (0.5 * float(fx["skin"])) ** 2
    steps, ago, xhold = [0], 0, x[0]
    for k in range(1, x.shape[0]):
        ago += 1
        rebuild = False
        if ago >= delay and ago % every == 0:
            d = x[k] - xhold
            rebuild = (not check) or bool(((d * d).sum(1) > trig).any())
        if rebuild:
            steps.append(k)
            ago, xhold = 0, x[k]
    return steps


@pytest.mark.parametrize("case", CASES)
def test_restated_schedule_reproduces_the_reference(case):
    fx = H.load_fixture(case)
    ref = np.nonzero(fx["ago"] == 0)[0].tolist()
    assert restated_rebuild_steps(fx) == ref
    # and ago itself counts the steps since the last rebuild
    ago = 0
    for k in range(1, len(fx["ago"])):
        ago = 0 if k in ref else ago + 1
        assert fx["ago"][k] == ago


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_library_rebuilds_on_the_reference_steps(case):
    from gpu_common import pb, c
    fx = H.load_fixture(case)
    delay, every, check = schedule(fx)
    words = str(fx["pair_style"]).split()
    s = pb.PairStyle(device=0)
    s.set_ntypes(2)
    s.command(str(fx["pair_style"]))
    s.command("pair_coeff 1 1 0.1 3.0")
    s.command("pair_coeff 2 2 0.1 3.0")
    s.init(g_ewald=float(fx["g_ewald"]), molecular=0, skin=float(fx["skin"]), neigh_every=every, neigh_delay=delay,
           neigh_check=int(check))
    s.set_box(fx["boxlo"], fx["boxhi"])
    n = fx["x"].shape[1]
    q, ty, al = c(fx["q"], np.float64), c(fx["type"], np.int32), c(fx["alpha"], np.float64)
    rebuilt = []
    for k in range(fx["x"].shape[0]):
        mu, f = np.zeros((n, 3)), np.zeros((n, 3))
        res = s.compute(c(fx["x"][k], np.float64), q, ty, al, mu, f, ago=-1)
        if res.status & pb.STATUS_REBUILT:
            rebuilt.append(k)
        if k == 0:
            assert H.rel_err(mu, fx["mu_first"]) < 1e-10 and H.rel_err(f, fx["f_first"]) < 1e-10
    assert rebuilt == np.nonzero(fx["ago"] == 0)[0].tolist()
    assert res.status & pb.STATUS_EXACT
    assert H.rel_err(mu, fx["mu_last"]) < 1e-10 and H.rel_err(f, fx["f_last"]) < 1e-10
    assert abs(res.eng_pol - float(fx["eng_pol_last"])) < 1e-10 * abs(float(fx["eng_pol_last"]))
    assert abs(res.eng_coul - float(fx["eng_coul_last"])) < 1e-10 * abs(float(fx["eng_coul_last"]))
    s.close()
