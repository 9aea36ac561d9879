import os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from movis_b200 import _native
L = _native.lib(); H, W = 2160, 3840
g = torch.Generator(device="cuda").manual_seed(0)
src = torch.randint(0, 256, (H, W, 4), dtype=torch.uint8, device="cuda", generator=g)
src[: H // 2, :, 0] = 0; src[: H // 2, :, 1] = 255; src[: H // 2, :, 2] = 0
dst = torch.empty_like(src)
lo = np.array([50, 178, 178], np.uint8); hi = np.array([70, 255, 255], np.uint8)
for i in range(3):
    _native.check(L.mvb_effect_chroma_key_rgba8(src.data_ptr(), W, H, W * 4, dst.data_ptr(), W * 4, lo.ctypes.data, hi.ctypes.data, torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
