"""Small helpers around torch buffers (torch is used for device memory and streams only)."""
import numpy
import torch


class DeviceArray(torch.Tensor):
    """A torch tensor with the one cupy method the reference's callers use on filter attributes:
    ``.get()`` -> numpy (e.g. ``pgf.means.get()``, tests/GSUKF_test.py:77-79;
    ``state_pdf.draw().get()``, sim_base.py:281-284)."""

    def get(self):
        return self.detach().as_subclass(torch.Tensor).cpu().numpy()


def wrap(t):
    return t.as_subclass(DeviceArray)


def to_numpy(a):
    """numpy view/copy of anything array-like a caller may hand in (numpy, list, torch, cupy-like)."""
    if isinstance(a, torch.Tensor):
        return a.detach().as_subclass(torch.Tensor).cpu().numpy()
    if hasattr(a, "get") and not isinstance(a, numpy.ndarray):
        return numpy.asarray(a.get())
    return numpy.asarray(a)


def resolve_device(device=None):
    if not torch.cuda.is_available():
        raise RuntimeError("gpu_se_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("gpu_se_b200 runs on CUDA devices only, got %r" % (device,))
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


def stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream


def round_up(n, m):
    return (n + m - 1) // m * m
