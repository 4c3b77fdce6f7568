"""``resample_from_cumsum``: systematic resampling fed a caller's own float64 cumulative sum
(contract (ii)).

The reference computes ``cumsum = numpy.cumsum(weights); cumsum /= cumsum[-1]`` (particle.py:89-90; the
GPU class: ``torch.cumsum``, :301-304) and then compares ``cumsum[k] < (i + r) / N`` (CPU loop, :96-100)
or walks with ``cumsum[k] > u`` (GPU kernel ``_parallel_resample``, :223-263).  Handing that very array to
``gse_resample_search_f64`` reproduces the reference's ``sample_index`` bit for bit: the only arithmetic
left is ``(i + r) / N`` in float64, evaluated as the reference evaluates it.
"""
import numpy
import torch

from gpu_se_b200 import _device, _lib
from gpu_se_b200.filter._base import Context


def resample_from_cumsum(cumsum, r, n_out=None, normalise=False, side="left", device=None, ctx=None):
    """Ancestor indices (int64 torch tensor on the device) of a systematic resample.

    cumsum    : (n,) float64 cumulative weights, numpy / torch / anything array-like; already normalised
                unless ``normalise`` (then every element is divided by ``cumsum[-1]`` on the device, as numpy does)
    r         : offset in [0, 1)   (``numpy.random.rand()``, particle.py:93)
    n_out     : number of output rows N in ``u_i = (i + r) / N`` (default: n)
    side      : 'left' = the reference's CPU loop (``cumsum[k] < u``), 'right' = its GPU kernel (``cumsum[k] <= u``)
    """
    if side not in ("left", "right"):
        raise ValueError("side must be 'left' or 'right'")
    dev = _device.resolve_device(device)
    if isinstance(cumsum, torch.Tensor):
        c = cumsum.detach().as_subclass(torch.Tensor).to(device=dev, dtype=torch.float64).reshape(-1).contiguous()
    else:
        c = torch.as_tensor(numpy.ascontiguousarray(_device.to_numpy(cumsum), dtype=numpy.float64).reshape(-1),
                            device=dev)
    n = c.numel()
    n_out = n if n_out is None else int(n_out)
    if c.data_ptr() % 32:
        c = c.clone()
    own = ctx is None
    if own:
        ctx = Context(dev, max(n, n_out), None, None)
    try:
        idx = torch.empty(_device.round_up(n_out, 64), dtype=torch.int32, device=dev)
        _lib.check(_lib.lib.gse_resample_search_f64(ctx.handle, c.data_ptr(), n, int(bool(normalise)),
                                                    int(side == "right"), float(r), n_out, 0, n_out, idx.data_ptr(),
                                                    _device.stream_ptr(dev)))
        out = idx[:n_out].to(torch.int64)
        torch.cuda.current_stream(dev).synchronize()
        ctx.check_device_errors()
    finally:
        if own:
            ctx.close()
    return out
