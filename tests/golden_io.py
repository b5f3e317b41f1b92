"""Loader for tests/golden/*.npz (written by oracle/gen_golden.py from the reference's own Python code).
bf16 tensors are stored as raw uint16 bit patterns; `f32()` turns them into their exact fp32 values."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def f32(a: np.ndarray) -> np.ndarray:
    if a.dtype == np.uint16:
        return (a.astype(np.uint32) << 16).view(np.float32)
    return a


def is_bf16(a: np.ndarray) -> bool:
    return a.dtype == np.uint16


class Golden:
    def __init__(self, name: str):
        self.z = np.load(os.path.join(GOLDEN_DIR, name))
        self.files = set(self.z.files)

    def raw(self, key):
        return self.z[key]

    def __getitem__(self, key):
        return f32(self.z[key])

    def __contains__(self, key):
        return key in self.files

    def torch(self, key, device="cpu"):
        """bf16-stored arrays come back as torch.bfloat16, everything else in its own dtype."""
        import torch
        a = self.z[key]
        if a.dtype == np.uint16:
            return torch.from_numpy(a.view(np.int16).copy()).view(torch.bfloat16).to(device)
        return torch.from_numpy(np.ascontiguousarray(a)).to(device)

    def weights(self, prefix="w."):
        return {k[len(prefix):]: k for k in self.files if k.startswith(prefix)}
