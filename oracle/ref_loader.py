"""Import the reference's own CPU classes: from /root/reference (build container), else from the unmodified copy
``oracle/_ref`` that ``oracle/make_ref.py`` makes there and that travels to the GPU box (git-ignored).

TEST INFRASTRUCTURE ONLY.  The reference imports ``cupy`` (filter/particle.py:6,
filter/gs_ukf.py:2, gaussian_sum_dist/MultivariateGaussianSum.py:3) and, through sim_base,
``osqp`` (controller.py:6); neither is installed.  Empty stub modules are registered in
``sys.modules`` so that the numpy code paths import unmodified.  Always pass ``library=numpy``.

``/root/reference`` does not exist on the GPU box: callers must check :func:`available` (true there only when
``oracle/_ref`` was shipped).
"""
import os
import sys
import types
import warnings

_COPY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REFERENCE_ROOT = os.environ.get("GSE_REFERENCE_ROOT", "/root/reference")
if not os.path.isdir(os.path.join(REFERENCE_ROOT, "filter")) and os.path.isdir(os.path.join(_COPY, "filter")):
    REFERENCE_ROOT = _COPY


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "filter"))


def is_copy():
    return os.path.abspath(REFERENCE_ROOT) == os.path.abspath(_COPY)


def load():
    """Returns a namespace with the reference modules ``filter``, ``gaussian_sum_dist``, ``model``."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for name in ("cupy", "osqp"):
        if name not in sys.modules:
            stub = types.ModuleType(name)
            stub.float32 = "stub"
            stub.__gse_stub__ = True
            sys.modules[name] = stub
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import filter as ref_filter            # noqa: A004  (the reference's package name)
        import gaussian_sum_dist as ref_gsd
        import model as ref_model
    return types.SimpleNamespace(filter=ref_filter, gaussian_sum_dist=ref_gsd, model=ref_model)
