"""The C-ABI library loads and exports every symbol include/t2fit.h declares; without a GPU it fails loudly."""
import ctypes as C
import os
import re

import pytest

from tests.conftest import ROOT
from fetal_t2mapping_b200 import _abi, build


@pytest.fixture(scope="module")
def lib():
    build.build_lib()
    return _abi.load_library()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "t2fit.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(t2fit_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    names = declared_symbols()
    assert len(names) >= 10
    bound = {n for n, _, _ in _abi.SYMBOLS}
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/t2fit.h but not exported"
        assert n in bound, f"{n} has no ctypes signature in _abi.SYMBOLS"
    assert lib.t2fit_abi_version() == _abi.ABI_VERSION


def test_struct_layout_matches_header(tmp_path):
    """sizeof / offsetof of every field, computed by gcc from include/t2fit.h, equal the ctypes mirror."""
    import subprocess
    fields = {"t2fit_problem": [f for f, _ in _abi.Problem._fields_], "t2fit_outputs": [f for f, _ in _abi.Outputs._fields_]}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "t2fit.h"', 'int main(void) {']
    for st, fs in fields.items():
        lines.append(f'  printf("{st} %zu\\n", sizeof({st}));')
        for f in fs:
            lines.append(f'  printf("{st}.{f} %zu\\n", offsetof({st}, {f}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for st, cls in (("t2fit_problem", _abi.Problem), ("t2fit_outputs", _abi.Outputs)):
        assert int(got[st]) == C.sizeof(cls), st
        for f, _ in cls._fields_:
            assert int(got[f"{st}.{f}"]) == getattr(cls, f).offset, f"{st}.{f}"


def test_dtype_codes_match_header():
    """The element-type codes of the ctypes mirror are the T2FIT_DT_* values of the header; host echo arrays map to them
    (float32 = 0 = default) and anything else is cast to float32 by the mirror."""
    import numpy as np
    from fetal_t2mapping_b200.api import _host_echoes
    hdr = open(os.path.join(ROOT, "include", "t2fit.h")).read()
    codes = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define T2FIT_DT_(\w+) (\d+)", hdr)}
    names = {"U8": "uint8", "I16": "int16", "U16": "uint16", "I32": "int32", "F32": "float32", "F64": "float64"}
    assert {names[k]: v for k, v in codes.items()} == {k: v for k, v in _abi.DTYPES.items() if k != "bool"}
    for name, code in _abi.ECHO_DTYPES.items():
        assert code == (0 if name == "float32" else _abi.DTYPES[name])
    for dt in ("float32", "float64", "int16", "uint16", "int32"):
        p = _abi.Problem()
        a = np.arange(12, dtype=dt).reshape(4, 3)
        b = _host_echoes(a, p)
        assert b is a and p.echo_dtype == _abi.ECHO_DTYPES[dt]
    for a in (np.arange(12, dtype=np.uint8).reshape(4, 3), np.arange(24, dtype=np.float64).reshape(4, 6)[:, ::2]):
        p = _abi.Problem()
        b = _host_echoes(a, p)                                  # unsupported type / not contiguous: cast as the reference does
        assert b.dtype == np.float32 and b.flags.c_contiguous and p.echo_dtype == 0 and np.array_equal(b, a)


def test_library_contains_sm100a_sass_only():
    out = os.popen(f"cuobjdump -lelf {_abi.LIB_PATH} 2>/dev/null").read()
    if not out.strip():
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out)


def test_no_gpu_means_loud_failure(lib):
    """There is no CPU fallback: without a device init reports ENODEVICE and compute calls ENOTINIT."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    assert lib.t2fit_init(0) == -2
    assert b"no CPU implementation" in lib.t2fit_last_error()
    p, o = _abi.Problem(), _abi.Outputs()
    assert lib.t2fit_run(C.byref(p), C.byref(o), None) == -3
    import numpy as np
    import fetal_t2mapping_b200 as t2
    with pytest.raises(_abi.T2FitError):
        t2.fit_voxels_batch(np.ones((4, 3), np.float32), None, [1.0, 2.0, 3.0], "gaussian", t2.preset("gaussian")[1])


def test_work_model_is_pure_and_consistent(lib):
    import fetal_t2mapping_b200 as t2
    a, b = t2.work_model("gaussian", 5), t2.work_model("gaussian", 10)
    assert a["bytes_per_voxel"] == 5 * 4 + 12 + 1
    assert b["flop_per_pass"] > a["flop_per_pass"] and t2.work_model("gaussian_rician", 5)["flop_per_pass"] > a["flop_per_pass"]
