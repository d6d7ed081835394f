"""Class loss from compact targets: drop-in for the active part of the reference's ``SMRSELDLoss``
(loss.py:149-172 — ``w_class *`` softmax-MSE (loss.py:43-54) or weighted cross entropy (loss.py:27-41); the AIUR and
converging-localisation terms are commented out of its ``forward``) that takes the batch's class-SET MASK instead of the
dense ``(B, T, I*J, M)`` float32 target tensor.

``DeviceLoader(..., targets="mask")`` yields ``(spectrograms, mask)`` with ``mask`` an int16 tensor ``(B, T, I*J)``
(bit c = an event of class c covers the cell, 0 = background): 5 MB per batch instead of 145 MB, painted by
``seld_batch_class_mask``; ``CompactSMRSELDLoss`` runs ``seld_class_loss`` forward and backward (autograd.Function), no
host synchronisation besides the ``.item()`` the reference itself does for its breakdown dict."""
from __future__ import annotations

import torch

from . import _lib

LOSS_TYPES = {"mse": 0, "ce": 1}


class _ClassLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, mask, loss_type, weight):
        if not logits.is_cuda or logits.dtype != torch.float32:
            raise ValueError("logits must be a float32 CUDA tensor (B, T, G, M)")
        z = logits.contiguous()
        M = z.shape[-1]
        n_cells = z.numel() // M
        if mask.dtype != torch.int16 or mask.numel() != n_cells or mask.device != z.device or not mask.is_contiguous():
            raise ValueError("mask must be a contiguous int16 tensor with one entry per (batch, frame, cell) on the logits' device")
        sums = torch.zeros(2, dtype=torch.float64, device=z.device)
        stream = torch.cuda.current_stream(z.device).cuda_stream
        _lib.check(_lib.lib().seld_class_loss(loss_type, z.data_ptr(), mask.data_ptr(), n_cells, M, _lib.ptr(weight),
                                              sums.data_ptr(), None, None, stream), "seld_class_loss")
        denom = sums[1] if loss_type == 1 else sums.new_tensor(float(n_cells * M))
        ctx.save_for_backward(z, mask, denom)
        ctx.loss_type, ctx.weight = loss_type, weight
        return (sums[0] / denom).to(torch.float32)

    @staticmethod
    def backward(ctx, grad_out):
        z, mask, denom = ctx.saved_tensors
        M = z.shape[-1]
        n_cells = z.numel() // M
        gscale = (grad_out.to(torch.float64) / denom).to(torch.float32).contiguous()  # device scalar: no synchronisation
        grad = torch.empty_like(z)
        stream = torch.cuda.current_stream(z.device).cuda_stream
        _lib.check(_lib.lib().seld_class_loss(ctx.loss_type, z.data_ptr(), mask.data_ptr(), n_cells, M, _lib.ptr(ctx.weight),
                                              None, grad.data_ptr(), gscale.data_ptr(), stream), "seld_class_loss")
        return grad, None, None, None


class _AuxLossFn(torch.autograd.Function):
    """(converging-localisation loss, AIUR loss) of reference loss.py:90-146 / :56-88 from logits and the class-set mask."""

    @staticmethod
    def forward(ctx, logits, mask, I, J):
        if not logits.is_cuda or logits.dtype != torch.float32 or logits.dim() != 4:
            raise ValueError("logits must be a float32 CUDA tensor (B, T, G, M)")
        z = logits.contiguous()
        B, T, G, M = z.shape
        if G != I * J:
            raise ValueError(f"grid {I} x {J} does not match {G} cells")
        if mask.dtype != torch.int16 or mask.numel() != B * T * G or mask.device != z.device or not mask.is_contiguous():
            raise ValueError("mask must be a contiguous int16 tensor with one entry per (batch, frame, cell) on the logits' device")
        sums = torch.zeros(3, dtype=torch.float64, device=z.device)
        stream = torch.cuda.current_stream(z.device).cuda_stream
        _lib.check(_lib.lib().seld_aux_losses(z.data_ptr(), mask.data_ptr(), B * T, I, J, M, sums.data_ptr(), None, None, stream),
                   "seld_aux_losses")
        denom = sums[1] * float(I * J) + 1e-10       # loss.py:143
        ctx.save_for_backward(z, mask, denom)
        ctx.grid = (I, J)
        cl = (sums[0] / denom).to(torch.float32)
        aiur = (1.0 - sums[2] / float(B * T)).to(torch.float32)  # loss.py:85-86; an argmax statistic: no gradient
        ctx.mark_non_differentiable(aiur)
        return cl, aiur

    @staticmethod
    def backward(ctx, grad_cl, _grad_aiur):
        z, mask, denom = ctx.saved_tensors
        B, T, G, M = z.shape
        gscale = (grad_cl.to(torch.float64) / denom).to(torch.float32).contiguous()  # device scalar: no synchronisation
        grad = torch.empty_like(z)
        stream = torch.cuda.current_stream(z.device).cuda_stream
        _lib.check(_lib.lib().seld_aux_losses(z.data_ptr(), mask.data_ptr(), B * T, ctx.grid[0], ctx.grid[1], M, None, grad.data_ptr(),
                                              gscale.data_ptr(), stream), "seld_aux_losses")
        return grad, None, None, None


class CompactSMRSELDLoss(torch.nn.Module):
    """``SMRSELDLoss(loss_type, w_class, ..., class_weights)`` of the reference (loss.py:9-25, :149-172) on compact targets:
    ``forward(y_pred (B, T, G, M) logits, mask (B, T, G) int16) -> (total_loss, {'class_<type>': float})``."""

    def __init__(self, loss_type="ce", w_class=1.0, w_aiur=0.5, w_cl=0.5, grid_size=None, class_weights=None,
                 use_aux_terms: bool = False):
        """``use_aux_terms=True`` adds ``w_aiur * aiur + w_cl * cl``, the combination the reference keeps commented out
        (loss.py:158-165); the default is the reference's live behaviour, ``w_class * class loss``."""
        super().__init__()
        self.use_aux_terms = use_aux_terms
        if loss_type not in LOSS_TYPES:
            raise ValueError("loss_type must be 'mse' or 'ce'")
        self.loss_type, self.w_class, self.w_aiur, self.w_cl = loss_type, w_class, w_aiur, w_cl
        self.I, self.J = grid_size if grid_size is not None else (None, None)
        self.register_buffer("class_weights", None if class_weights is None else
                             torch.as_tensor(class_weights, dtype=torch.float32).contiguous(), persistent=False)

    def forward(self, y_pred: torch.Tensor, mask: torch.Tensor):
        w = self.class_weights
        if w is not None and w.device != y_pred.device:
            w = self.class_weights = w.to(y_pred.device)
        loss_class = _ClassLossFn.apply(y_pred, mask, LOSS_TYPES[self.loss_type], w if self.loss_type == "ce" else None)
        total_loss = self.w_class * loss_class
        breakdown = {f"class_{self.loss_type}": float(loss_class.item())}
        if self.use_aux_terms:
            cl, aiur = self._aux(y_pred, mask)
            total_loss = total_loss + self.w_aiur * aiur + self.w_cl * cl
            breakdown["aiur"], breakdown["cl"] = float(aiur.item()), float(cl.item())
        return total_loss, breakdown

    def _aux(self, y_pred, mask):
        if self.I is None or self.J is None:
            raise ValueError("the AIUR / converging-localisation terms need grid_size=(I, J)")
        return _AuxLossFn.apply(y_pred, mask, int(self.I), int(self.J))

    def aiur_loss(self, y_pred: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        """``SMRSELDLoss.aiur_loss`` (loss.py:56-88) of softmax(y_pred) against the targets the mask stands for."""
        return self._aux(y_pred, mask)[1]

    def converging_localization_loss(self, y_pred: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        """``SMRSELDLoss.converging_localization_loss`` (loss.py:90-146) of softmax(y_pred); differentiable in y_pred."""
        return self._aux(y_pred, mask)[0]
