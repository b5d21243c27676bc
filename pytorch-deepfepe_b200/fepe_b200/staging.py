"""Host <-> device staging for the end-to-end path.

A step's inputs (matches, weights, intrinsics, GT pose, virtual points) live in ONE pinned host
buffer with a fixed layout and are moved with ONE cudaMemcpyAsync; the results a caller reads back
(F, pose/loss rows) come back in ONE copy.  Eight small copies per step cost more in launch overhead
than the 5.8 MB payload costs in PCIe time.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch

from . import _lib, ops

_IN_FIELDS = (("matches_xy_ori", 4), ("weights", 1), ("Ks", None), ("q_cam", None), ("t_cam", None),
              ("delta_Rtijs_4_4", None), ("pts1_virt", None), ("pts2_virt", None))


class StagedStep:
    """Pinned host + device mirrors of one batch of B pairs x N correspondences (V virtual points)."""

    def __init__(self, B: int, N: int, V: int, device):
        self.B, self.N, self.V, self.device = B, N, V, device
        shapes = {"matches_xy_ori": (B, N, 4), "weights": (B, N), "Ks": (B, 3, 3), "q_cam": (B, 4), "t_cam": (B, 3),
                  "delta_Rtijs_4_4": (B, 4, 4), "pts1_virt": (B, V, 3), "pts2_virt": (B, V, 3)}
        self._slices: Dict[str, Tuple[int, int, tuple]] = {}
        off = 0
        for name, shp in shapes.items():
            n = int(np.prod(shp))
            self._slices[name] = (off, n, shp)
            off += (n + 3) // 4 * 4                      # keep every field 16-byte aligned
        self.h_in = torch.empty(off, dtype=torch.float32).pin_memory()
        self.d_in = torch.empty(off, dtype=torch.float32, device=device)
        n_out = B * 9 + B * _lib.POSE_OUT_FLOATS
        self.h_out = torch.empty(n_out, dtype=torch.float32).pin_memory()
        self.d_out = torch.empty(n_out, dtype=torch.float32, device=device)
        self.d_res = torch.empty(B, N, dtype=torch.float32, device=device)
        self.d_epi = torch.empty(B, N, dtype=torch.float32, device=device)
        self.in_bytes = off * 4
        self.out_bytes = n_out * 4

    def _view(self, buf, name):
        off, n, shp = self._slices[name]
        return buf[off:off + n].view(shp)

    def pack(self, batch: dict, out: torch.Tensor = None) -> torch.Tensor:
        """Lay a synth/dataset style dict of numpy arrays out in a pinned host buffer (what a DataLoader
        collate_fn with pin_memory would produce); returns the buffer (`out` or a new one)."""
        buf = out if out is not None else torch.empty_like(self.h_in).pin_memory()
        for name, _ in _IN_FIELDS:
            self._view(buf, name).copy_(torch.from_numpy(np.ascontiguousarray(batch[name])).reshape(
                self._slices[name][2]))
        return buf

    def run(self, stream, affine, clamp_epi: float = 0.5, clamp_loss: float = 0.02, host: torch.Tensor = None) -> None:
        """H2D (1 copy) -> fepe_fit_pose_fwd (fit + pose head) -> D2H (1 copy), all on `stream`.
        `host` is a buffer made by pack(); default: this object's own pinned buffer."""
        with torch.cuda.stream(stream):
            self.d_in.copy_(host if host is not None else self.h_in, non_blocking=True)
            v = lambda k: self._view(self.d_in, k)
            F = self.d_out[:self.B * 9].view(self.B, 3, 3)
            pose = self.d_out[self.B * 9:].view(1, self.B, _lib.POSE_OUT_FLOATS)
            ops.fit_pose_forward(v("matches_xy_ori"), v("weights"), affine, v("Ks"), v("q_cam"), v("t_cam"),
                                 v("delta_Rtijs_4_4"), v("pts1_virt"), v("pts2_virt"), clamp_at=clamp_epi,
                                 virt_clamp_at=clamp_loss, out=(F, self.d_res, self.d_epi, None, pose[0]))
            self.h_out.copy_(self.d_out, non_blocking=True)

    def capture(self, stream, affine, clamp_epi: float = 0.5, clamp_loss: float = 0.02) -> None:
        """Record run() from this object's own pinned buffer into a CUDA graph (H2D memcpy node -> kernel -> D2H
        memcpy node).  Afterwards replay() costs one cudaGraphLaunch on the host instead of ~15 Python / driver calls
        (at 256 pairs per step the device needs ~110 us of PCIe time; the eager host path alone takes longer)."""
        self.run(stream, affine, clamp_epi, clamp_loss)          # warm-up outside capture (lazy init, attributes)
        stream.synchronize()
        self._graph = torch.cuda.CUDAGraph()
        self._stream = stream
        with torch.cuda.graph(self._graph, stream=stream):
            self.run(stream, affine, clamp_epi, clamp_loss)

    def replay(self) -> None:
        """One end-to-end step from self.h_in (fill it with pack(batch, out=self.h_in)): results in self.h_out once
        the stream is synchronised."""
        with torch.cuda.stream(self._stream):
            self._graph.replay()

    def results(self):
        """Host views of the last run's F [B,3,3] and pose rows [B,32] (valid after a stream sync)."""
        return (self.h_out[:self.B * 9].view(self.B, 3, 3),
                self.h_out[self.B * 9:].view(self.B, _lib.POSE_OUT_FLOATS))
