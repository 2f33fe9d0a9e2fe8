"""Host logic without a GPU: the C-ABI library loads and exports every symbol include/dfd_b200.h declares, the
ctypes signature table agrees with the header, and the engine's call plan (arenas, pointer arithmetic, argument
lists) builds for the BASELINE configurations."""
import os
import re

import pytest

from deepfake_detection_b200 import _lib
from deepfake_detection_b200.arch import get_spec, param_entries

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _protos():
    hdr = open(os.path.join(ROOT, "include", "dfd_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return re.findall(r"\b(?:int|const char\*)\s+(dfd_\w+)\s*\(([^)]*)\)\s*;", hdr)


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    names = [n for n, _ in _protos()]
    assert len(names) >= 30
    for n in names:
        assert hasattr(L.cdll, n), n
    assert L.cdll.dfd_abi_version() == 1
    assert L.stat_slots == 8


def test_ctypes_signatures_match_header():
    for name, params in _protos():
        if name == "dfd_last_error":
            continue
        codes = ""
        if params.strip() != "void":
            for p in params.split(","):
                p = p.strip()
                codes += "p" if "*" in p else "l" if "long long" in p else "f" if p.startswith("float") else \
                    "d" if p.startswith("double") else "i"
        assert _lib.SIGNATURES[name] == codes, name


@pytest.mark.parametrize("arch,batch,res", [("efficientnet_b0", 2, 64), ("efficientnet_b4", 1, 76),
                                             ("efficientnet_b0", 4, 224)])
def test_engine_plan_builds(arch, batch, res):
    from deepfake_detection_b200.engine import Engine
    eng = Engine(arch, batch, res, res, device="plan-only")
    spec = get_spec(arch)
    n = sum(int(__import__("math").prod(s)) for _, s, _ in param_entries(spec))
    assert eng.n_params >= n
    assert set(eng.p_off) == {e[0] for e in param_entries(spec)}
    assert eng.n_launch["fwd"] > 100 and eng.n_launch["bwd"] > 150
    # arena offsets never overlap
    spans = sorted((o, o + k) for o, _, k in eng.p_off.values())
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))


def test_no_oracle_on_product_path():
    """The product package must never import oracle/ (or fall back to torch compute)."""
    pkg = os.path.join(ROOT, "deepfake_detection_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
