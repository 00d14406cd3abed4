"""CPU: the oracle and the host Init against the reference binary on random configurations
(a short run of tests/golden/fuzz_oracle_vs_reference.py; skipped where oracle/_ref is absent)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, str(ROOT / "tests" / "golden"))
import fuzz_oracle_vs_reference as F  # noqa: E402


@pytest.mark.skipif(not F.REF.exists(), reason="oracle/_ref/fv2d_ref not built (needs /root/reference)")
@pytest.mark.parametrize("seed", [101, 102])
def test_oracle_matches_reference_on_random_configurations(seed):
    env = dict(os.environ, OMP_NUM_THREADS="1", OMP_PROC_BIND="false")
    rng = np.random.default_rng(seed)
    tally = {}
    for _ in range(60):
        base, ov = F.draw(rng)
        res, msg = F.one(base, ov, 5, env)
        tally[res] = tally.get(res, 0) + 1
        assert not res.startswith("MISMATCH") and res != "ref-failed", (base, ov, msg)
    assert tally.get("ok", 0) >= 40
