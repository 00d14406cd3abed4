import sys
import tempfile
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

GOLDEN_DIR = ROOT / "tests" / "golden"
GOLDEN_NAMES = sorted(p.stem for p in GOLDEN_DIR.glob("*.npz"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


class Golden:
    """One fixture dumped by the reference itself (tests/golden/make_goldens.py)."""

    def __init__(self, name):
        self.name = name
        z = np.load(GOLDEN_DIR / f"{name}.npz")
        self.ini_text = str(z["ini_text"])
        self.params_line = str(z["params_line"])
        self.warnings = str(z["warnings"])
        self.nsteps = int(z["nsteps"])
        self.dts = z["dts"]
        self.Q0, self.QN, self.UN = z["Q0"], z["QN"], z["UN"]
        self.mass, self.energy, self.t = float(z["mass"]), float(z["energy"]), float(z["t"])
        self._tmp = None

    def ini_path(self):
        if self._tmp is None:
            self._tmp = tempfile.NamedTemporaryFile("w", suffix=f"_{self.name}.ini", delete=False)
            self._tmp.write(self.ini_text)
            self._tmp.close()
        return self._tmp.name

    def ref_params(self):
        """key -> float from the reference's own printout of its DeviceParams."""
        out = {}
        for tok in self.params_line.split()[1:]:
            k, v = tok.split("=")
            out[k] = float(v)
        return out


@pytest.fixture(params=GOLDEN_NAMES)
def golden(request):
    return Golden(request.param)


def load_golden(name):
    return Golden(name)


def rel_l1(a, b):
    """relative L1 distance, the norm BASELINE.json's parity bar is stated in"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.sum(np.abs(b))
    return float(np.sum(np.abs(a - b)) / den) if den > 0 else float(np.sum(np.abs(a - b)))
