"""The CPU oracle against the committed golden vectors (tests/golden/*.npz,
generated from the unmodified reference sources by tests/golden/make_golden.py).
Teacher-forced: every step restarts from the recorded state."""
import glob
import os

import numpy as np
import pytest

from helpers import make_oracle

# the control() fixtures (make_golden.py); c3_phik_8192 / models_entropy have their own tests
GOLDEN = sorted(f for f in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                if os.path.basename(f).startswith(("c1_", "c2_", "c4_", "c5_", "omni_")))
TIGHT = 1e-12


def test_fixtures_present():
    assert len(GOLDEN) >= 8


@pytest.mark.parametrize("path", [p for p in GOLDEN if os.path.basename(p).startswith("c1_")], ids=os.path.basename)
def test_closed_loop_cases(path):
    g = np.load(path)
    bounds, model = tuple(g["bounds"]), int(g["model"])
    o = make_oracle(model, mu=g["mu"])
    for s in range(len(g["x"])):
        o.set_ut(g["ut_before"][s])
        u0 = o.control(bounds, g["x"][s])
        np.testing.assert_allclose(u0, g["u0"][s], rtol=0, atol=TIGHT)
        np.testing.assert_allclose(o.get_ut(), g["ut_after"][s], rtol=0, atol=TIGHT)
        np.testing.assert_allclose(o.last()["ck"], g["ck"][s], rtol=0, atol=TIGHT)
        o.add_state_memory(g["mem_final"][s])
    np.testing.assert_allclose(o.get_phik(), g["phik"], rtol=0, atol=TIGHT)


@pytest.mark.parametrize("path", [p for p in GOLDEN if not os.path.basename(p).startswith("c1_")],
                         ids=os.path.basename)
def test_batch_cases(path):
    g = np.load(path)
    model, nb, horizon = int(g["model"]), int(g["nb"]), float(g["horizon"])
    for i in range(len(g["x"])):
        o = make_oracle(model, nb=nb, horizon=horizon)
        o.set_ut(g["ut_before"][i])
        for m in g["mem"]:
            o.add_state_memory(m[i])
        u0 = o.control(tuple(g["bounds"]), g["x"][i])
        np.testing.assert_allclose(u0, g["u0"][i], rtol=0, atol=TIGHT)
        np.testing.assert_allclose(o.get_ut(), g["ut_after"][i], rtol=0, atol=TIGHT)
        np.testing.assert_allclose(o.last()["ck"], g["ck"][i], rtol=0, atol=TIGHT)
    np.testing.assert_allclose(o.get_phik(), g["phik"], rtol=0, atol=TIGHT)
