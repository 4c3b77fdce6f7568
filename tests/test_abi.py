"""The C-ABI library loads and exports every symbol include/gse.h declares; argument validation of
the entry points that can be exercised without a GPU.  CPU only (no compute calls)."""
import ctypes
import os
import re

import numpy
import pytest

from conftest import ROOT


def declared_functions():
    text = open(os.path.join(ROOT, "include", "gse.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gse_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from gpu_se_b200 import _lib
    names = declared_functions()
    assert len(names) >= 18
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in names:
        assert hasattr(raw, name), "libgse_b200.so does not export %s" % name
    # the ctypes signature table covers exactly the header
    assert sorted(_lib.SIGNATURES) == names
    assert _lib.lib.gse_abi_version() == _lib.GSE_ABI_VERSION


def test_struct_layout_matches_header():
    from gpu_se_b200 import _lib
    assert ctypes.sizeof(_lib.gse_mixture) == 8 + 8 * (8 + 40 + 200)
    m = _lib.make_mixture(numpy.zeros((2, 5)), numpy.stack([numpy.eye(5)] * 2), [0.75, 0.25])
    assert (m.nd, m.nx) == (2, 5) and m.weights[1] == 0.25 and m.covs[25 + 6] == 1.0
    with pytest.raises(ValueError):
        _lib.make_mixture(numpy.zeros((9, 5)), numpy.stack([numpy.eye(5)] * 9), numpy.ones(9))


def test_ctypes_structs_match_the_c_header(tmp_path):
    """Compile include/gse.h with gcc and compare every struct's size and field offsets with the ctypes
    mirrors (the header is plain C: any FFI can consume it)."""
    import json
    import shutil
    import subprocess
    from gpu_se_b200 import _lib
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    exe = str(tmp_path / "abi_probe")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "abi_probe.c"), "-o", exe], check=True)
    c = json.loads(subprocess.run([exe], check=True, capture_output=True, text=True).stdout)
    for name in ("gse_mixture", "gse_shards", "gse_step_params"):
        cls = getattr(_lib, name)
        assert ctypes.sizeof(cls) == c["sizeof." + name], name
        for field, _ in cls._fields_:
            off, size = c["%s.%s" % (name, field)]
            f = getattr(cls, field)
            assert (f.offset, f.size) == (off, size), (name, field)
    assert c["GSE_ABI_VERSION"] == _lib.GSE_ABI_VERSION
    assert c["GSE_MAX_SHARDS"] == _lib.GSE_MAX_SHARDS and c["GSE_MAX_ND"] == _lib.GSE_MAX_ND
    assert c["GSE_MAILBOX_BYTES"] == _lib.GSE_MAILBOX_BYTES and c["GSE_IPC_HANDLE_BYTES"] == _lib.GSE_IPC_HANDLE_BYTES


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    from gpu_se_b200 import _lib
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.GseError, match="no CPU fallback"):
        _lib._load()


def test_unknown_model_is_rejected_without_fallback():
    from gpu_se_b200.model.BioreactorModel import Bioreactor, model_id_for
    assert model_id_for(Bioreactor.homeostatic_DEs, Bioreactor.static_outputs) == 1
    with pytest.raises(NotImplementedError, match="no CPU fallback"):
        model_id_for(lambda x, u, dt: x, lambda x, u: x)


def test_no_gpu_means_error_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import gpu_se_b200 as g
    sp = g.MultivariateGaussianSum(numpy.zeros((1, 5)), numpy.eye(5)[None], numpy.ones(1))
    mp = g.MultivariateGaussianSum(numpy.zeros((1, 2)), numpy.eye(2)[None], numpy.ones(1))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        g.ParticleFilter(g.Bioreactor.homeostatic_DEs, g.Bioreactor.static_outputs, 16, sp, sp, mp)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "gpu_se_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_docs_quote_the_current_abi():
    """DESIGN.md states the ABI version and the number of entry points: keep both in step with include/gse.h."""
    from gpu_se_b200 import _lib
    design = open(os.path.join(ROOT, "DESIGN.md")).read()
    m = re.search(r"ABI v(\d+), (\d+) `extern \"C\"` entry points", design)
    assert m, "DESIGN.md section 0 no longer states the ABI version / entry-point count"
    assert int(m.group(1)) == _lib.GSE_ABI_VERSION
    assert int(m.group(2)) == len(declared_functions())
    header = open(os.path.join(ROOT, "include", "gse.h")).read()
    assert re.search(r"#define GSE_ABI_VERSION %d\b" % _lib.GSE_ABI_VERSION, header)
    # every entry point is attributed to the reference interface it replaces (or marked new) in INTEGRATION.md
    integration = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    groups = {"gse_peer_alloc": "gse_peer_alloc/open/close/free", "gse_peer_open": "gse_peer_alloc/open/close/free",
              "gse_peer_close": "gse_peer_alloc/open/close/free", "gse_peer_free": "gse_peer_alloc/open/close/free",
              "gse_peer_allgather_stats": "gse_peer_allgather_stats/_totals/_moments",
              "gse_peer_allgather_totals": "gse_peer_allgather_stats/_totals/_moments",
              "gse_peer_allgather_moments": "gse_peer_allgather_stats/_totals/_moments"}
    missing = [n for n in declared_functions() if n not in integration and groups.get(n, "\0") not in integration]
    assert not missing, "INTEGRATION.md does not mention: %s" % ", ".join(missing)
