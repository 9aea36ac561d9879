"""Device plumbing: CUDA tensors, streams and the device-resident render cache.

PyTorch is used only for device memory, streams and (in bench/range rendering) process plumbing;
all pixel work goes through the C-ABI library (``_native``).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Hashable

import numpy as np
import torch

from . import _native


class NoDeviceError(RuntimeError):
    pass


def require_cuda() -> None:
    """The product path has no CPU fallback: fail loudly without a usable GPU + extension."""
    _native.lib()
    if not torch.cuda.is_available():
        raise NoDeviceError("movis_b200 needs a CUDA device (B200 / sm_100a); there is no CPU fallback")


def gpu_numa_node(device_index: int) -> int:
    """NUMA node of the GPU's PCIe slot from /sys/bus/pci/devices/<bdf>/numa_node (-1: unknown or not
    exposed, e.g. on a single-node virtual machine).  NVML's CPU affinity is useless on such hosts
    (every CPU is reported local to every GPU); sysfs is the kernel's own view."""
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as fh:
            return int(fh.read().strip())
    except Exception:
        return -1


def _parse_cpulist(text: str) -> list[int]:
    cpus: list[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_host_to_gpu(device_index: int) -> dict:
    """Pin the calling process to the CPUs of the GPU's NUMA node and prefer that node for its memory,
    so that pinned staging buffers allocated afterwards live next to that GPU's PCIe root.  Matters when
    several ranks deliver frames to host memory at once.  Returns what was done:
    ``{"numa_node": n, "cpus": [...], "mempolicy": bool}``; with an unknown node (``-1``) or a node
    without usable CPUs nothing is changed."""
    import ctypes
    import os
    info = {"numa_node": gpu_numa_node(device_index), "cpus": [], "mempolicy": False}
    node = info["numa_node"]
    if node < 0:
        return info
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            usable = sorted(set(_parse_cpulist(fh.read())) & os.sched_getaffinity(0))
        if usable:
            os.sched_setaffinity(0, usable)
            info["cpus"] = usable
        # set_mempolicy(MPOL_PREFERRED, {node}): new pages (incl. the driver's pinned allocations made on
        # behalf of this thread) come from the GPU's node when it has room
        mask = ctypes.c_ulong(1 << node)
        libc = ctypes.CDLL(None, use_errno=True)
        if node < 64 and libc.syscall(238, 1, ctypes.byref(mask), 65) == 0:   # x86-64: __NR_set_mempolicy
            info["mempolicy"] = True
    except Exception:
        pass
    return info


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def is_device_image(x) -> bool:
    return isinstance(x, torch.Tensor)


def check_image_array(a: np.ndarray, what: str = "image") -> None:
    """The output contract the reference asserts on layer images (composition.py:783-786)."""
    assert isinstance(a, np.ndarray), f"{what} must be a numpy array"
    assert a.dtype == np.uint8, f"{what} must have dtype=np.uint8"
    assert a.ndim == 3, f"{what} must have 3 dimensions (H, W, C)"
    assert a.shape[2] == 4, f"{what} must have 4 channels (RGBA)"


def to_device(image, what: str = "image") -> torch.Tensor:
    """uint8 (H, W, 4) numpy array or tensor -> CUDA tensor whose pixels are 4 contiguous bytes.

    Non-contiguous host views (e.g. crops, ops.py:296) are compacted on the host first."""
    require_cuda()
    if isinstance(image, torch.Tensor):
        t = image
        assert t.dtype == torch.uint8 and t.ndim == 3 and t.shape[2] == 4, f"{what} must be uint8 (H, W, 4)"
        if not t.is_cuda:
            t = t.cuda(non_blocking=True)
        if t.stride(2) != 1 or t.stride(1) != 4 or t.stride(0) % 4 != 0 or t.data_ptr() % 4 != 0:
            t = t.contiguous()
        return t
    check_image_array(image, what)
    host = np.ascontiguousarray(image)
    if not host.flags.writeable:  # torch refuses read-only views; the upload never writes to it
        host = host.copy()
    return torch.from_numpy(host).cuda()


def upload_source(image: np.ndarray) -> torch.Tensor:
    """``to_device`` for a layer's own pixels, remembering whether their alpha is 255 throughout (one
    reduction, once per upload): ``fill_descriptors`` passes that on as
    MVB_LAYER_OPAQUE_SOURCE so the kernel can skip what such a layer hides.  The mark lives on this
    tensor object only: views, crops and effect outputs do not inherit it."""
    t = to_device(image)
    # on the device, right behind the copy: 50 us, where numpy needs milliseconds per 4K image
    if t.numel():
        lowest = int(t[..., 3].min().item())
        if lowest == 255:
            t.mvb_opaque = True
        elif lowest == 0:
            attach_alpha_map(t)
    return t


def attach_alpha_map(t: torch.Tensor) -> torch.Tensor:
    """Compute the alpha map of a device image (the largest alpha of every 32 x 32 block,
    ``mvb_alpha_map_rgba8``) and keep it on the tensor object: ``fill_descriptors`` hands it to the fused
    kernel, which then skips the layer in tiles that only see fully transparent blocks (a keyed-out green
    screen, the margins of a sprite).  Work elimination only; views and effect outputs do not inherit it."""
    from . import _native
    ptr, w, h, stride = image_args(t)
    grid = torch.empty(((h + 31) // 32, (w + 31) // 32), dtype=torch.uint8, device=t.device)
    _native.check(_native.lib().mvb_alpha_map_rgba8(ptr, w, h, stride, grid.data_ptr(), stream_ptr()))
    t.mvb_alpha_map = grid
    return t


def to_host(t: torch.Tensor) -> np.ndarray:
    return t.contiguous().cpu().numpy()


def empty_image(h: int, w: int) -> torch.Tensor:
    return torch.empty((h, w, 4), dtype=torch.uint8, device="cuda")


def image_args(t: torch.Tensor) -> tuple[int, int, int, int]:
    """(ptr, width, height, row stride in bytes) for the C ABI."""
    return t.data_ptr(), t.shape[1], t.shape[0], t.stride(0)


class DeviceCache:
    """Device-resident render cache: hashable key -> CUDA tensor, LRU within a byte budget.

    Stands in for the reference's ``diskcache.Cache(size_limit=1 GiB)`` (composition.py:65), which
    pickles frames to disk; entries here never leave HBM."""

    def __init__(self, size_limit: int = 1024 * 1024 * 1024) -> None:
        self.size_limit = int(size_limit)
        self._entries: "OrderedDict[Hashable, torch.Tensor]" = OrderedDict()
        self._bytes = 0
        self.hits = 0
        self.misses = 0

    @staticmethod
    def _nbytes(t: torch.Tensor) -> int:
        return t.numel() * t.element_size()

    def __contains__(self, key: Hashable) -> bool:
        return key in self._entries

    def __len__(self) -> int:
        return len(self._entries)

    @property
    def nbytes(self) -> int:
        return self._bytes

    def get(self, key: Hashable):
        t = self._entries.get(key)
        if t is None:
            self.misses += 1
            return None
        self._entries.move_to_end(key)
        self.hits += 1
        return t

    def __getitem__(self, key: Hashable) -> torch.Tensor:
        t = self.get(key)
        if t is None:
            raise KeyError(key)
        return t

    def __setitem__(self, key: Hashable, value: torch.Tensor) -> None:
        if key in self._entries:
            self._bytes -= self._nbytes(self._entries.pop(key))
        n = self._nbytes(value)
        if n > self.size_limit:
            return
        self._entries[key] = value
        self._bytes += n
        while self._bytes > self.size_limit and self._entries:
            _, old = self._entries.popitem(last=False)
            self._bytes -= self._nbytes(old)

    def __delitem__(self, key: Hashable) -> None:
        self._bytes -= self._nbytes(self._entries.pop(key))

    def clear(self) -> None:
        self._entries.clear()
        self._bytes = 0
