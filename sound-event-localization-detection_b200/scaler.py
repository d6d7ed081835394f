"""Dataset normalisation scaler (not in the reference — parity unpinned).

Per feature (channel, mel): count, sum and sum of squares over the kept frames of the training split are
accumulated on each GPU by the feature call (``seld_features(..., d_stats)``); ranks combine them with ONE
``all_reduce`` of ``2*C*M + 1`` float64 values (about 7 KB for the 7-channel FOA feature) — the only collective
of the whole front-end — and every rank derives mean / std locally.  ``apply`` runs the in-place
``(x - mean) / std`` kernel."""
from __future__ import annotations

import torch

from . import _lib


class FeatureScaler:
    def __init__(self, n_features: int, device=None):
        self.n_features = int(n_features)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        # [sum (F) | sum of squares (F) | count (1)] in one float64 buffer so that a single all_reduce suffices
        self.buf = torch.zeros(2 * self.n_features + 1, dtype=torch.float64, device=self.device)
        self.mean = None
        self.std = None
        self._mean32 = self._inv32 = None

    @property
    def stats(self) -> torch.Tensor:
        """View handed to ``FeaturePlan.run(stats=...)`` / ``seld_features`` (2*F float64)."""
        return self.buf[: 2 * self.n_features]

    def add_count(self, frames: int) -> None:
        self.buf[-1] += float(frames)

    def merge(self, stats: torch.Tensor, frames: int) -> None:
        self.buf[: 2 * self.n_features] += stats.to(self.device)
        self.add_count(frames)

    def sync(self, group=None) -> None:
        """Sum the partials over all ranks (NCCL all-reduce over NVLink when the backend is nccl)."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.buf, op=dist.ReduceOp.SUM, group=group)

    def finalize(self, eps: float = 0.0):
        n = self.buf[-1]
        if float(n) <= 0:
            raise ValueError("FeatureScaler.finalize: no frames accumulated")
        F = self.n_features
        self.mean = self.buf[:F] / n
        var = torch.clamp(self.buf[F:2 * F] / n - self.mean * self.mean, min=0.0)
        self.std = torch.sqrt(var + eps)
        safe = torch.where(self.std > 0, self.std, torch.ones_like(self.std))
        self._mean32 = self.mean.to(torch.float32).contiguous()
        self._inv32 = (1.0 / safe).to(torch.float32).contiguous()
        return self.mean, self.std

    def apply(self, x: torch.Tensor) -> torch.Tensor:
        """In-place (x - mean) / std over the trailing feature axes (numel of them == n_features)."""
        if self._mean32 is None:
            self.finalize()
        if not x.is_cuda or x.dtype != torch.float32 or not x.is_contiguous():
            raise ValueError("x must be a contiguous float32 CUDA tensor")
        if x.numel() % self.n_features:
            raise ValueError("trailing size does not match n_features")
        rows = x.numel() // self.n_features
        stream = torch.cuda.current_stream(x.device).cuda_stream
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().seld_scaler_apply(x.data_ptr(), rows, self.n_features, self._mean32.data_ptr(),
                                                    self._inv32.data_ptr(), stream), "seld_scaler_apply")
        return x

    def state_dict(self):
        return {"buf": self.buf.cpu(), "n_features": self.n_features}

    def load_state_dict(self, sd):
        self.buf.copy_(sd["buf"])
        self.finalize()
