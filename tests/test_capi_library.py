"""The C-ABI library itself: loads without a GPU, exports every symbol the header declares,
and fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import subprocess

import numpy as np
import pytest

from conftest import load_golden
from fv2d_b200 import capi


def test_library_loads_and_exports_every_declared_symbol():
    L = capi.lib()
    names = capi.exported_symbols_in_header()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert L.fv2d_abi_version() == 1


def test_symbols_are_plain_c_abi():
    out = subprocess.run(["nm", "-D", "--defined-only", str(capi.LIB_PATH)], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    for n in capi.exported_symbols_in_header():
        assert n in exported, n  # unmangled


def test_struct_layouts_match_the_header():
    # sizes as laid out by the C compiler for include/fv2d_params.h
    src = '#include "include/fv2d_params.h"\n#include <stdio.h>\nint main(){printf("%zu %zu\\n", sizeof(fv2d_device_params), sizeof(fv2d_run_params));}'
    import tempfile, os
    from pathlib import Path
    root = Path(capi.__file__).resolve().parents[1]
    with tempfile.TemporaryDirectory() as td:
        (Path(td) / "s.c").write_text(src)
        subprocess.run(["gcc", "-I", str(root), "-o", f"{td}/s", f"{td}/s.c"], check=True, cwd=root)
        a, b = subprocess.run([f"{td}/s"], capture_output=True, text=True).stdout.split()
    assert int(a) == C.sizeof(capi.DeviceParams) and int(b) == C.sizeof(capi.RunParams)


def test_no_cpu_fallback():
    """Without a CUDA device the compute entry points must fail with FV2D_ERR_CUDA."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    dev, run = capi.params_from_ini(load_golden("sod_x").ini_path())
    with pytest.raises(capi.Fv2dError) as e:
        capi.Context(dev)
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)


def test_argument_validation_needs_no_gpu():
    dev, run = capi.params_from_ini(load_golden("sod_x").ini_path())
    bad = dev.copy()
    bad.Ng = 1
    with pytest.raises(capi.Fv2dError) as e:
        capi.Context(bad)
    assert e.value.code == 1
    bad = dev.copy()
    bad.thermal_conductivity_active, bad.thermal_conductivity_mode = 1, capi.TCM_B02
    with pytest.raises(capi.Fv2dError) as e:
        capi.Context(bad)
    assert e.value.code == 1 and "B02" in str(e.value)
    with pytest.raises(capi.Fv2dError) as e:  # 16 rows over 8 slabs: thinner than the two ghost layers
        capi.Context(dev, nranks=8, rank=0)
    assert e.value.code == 1 and "thinner" in str(e.value)
    with pytest.raises(capi.Fv2dError) as e:
        capi.Context(dev, nranks=9, rank=0)
    assert e.value.code == 1
