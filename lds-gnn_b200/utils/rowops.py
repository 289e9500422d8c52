"""Fused, twice-differentiable row-local ops of the unrolled inner steps (kernels: csrc/lds_rowops.cu).

`masked_nll(logits, info)` = F.nll_loss(F.log_softmax(logits, 1)[mask], y[mask]) and the accuracy on the same rows
(reference: src/models/gcn.py:34 + src/trainers/inner.py:63-66 / src/trainers/outer.py:65-67) as ONE launch; the gradient
w.r.t. the logits is one launch, and so is the gradient of that gradient, which the hyper step takes when it differentiates
through the inner optimiser's `autograd.grad(loss, params, create_graph=True)` (src/trainers/inner.py:71 via higher).
As ATen ops the three are ~7 + 6 + 15 kernels of 1-2 us each per inner step.
"""
import torch


class MaskInfo:
    """Index structures of one boolean node mask: rows (int64 [m]), slot (int32 [n]: position among the rows or -1), labels y [n]."""

    def __init__(self, mask: torch.Tensor, y: torch.Tensor):
        rows = mask.nonzero().flatten()
        self.rows = rows
        self.m = int(rows.numel())
        slot = torch.full((mask.numel(),), -1, dtype=torch.int32, device=mask.device)
        slot[rows] = torch.arange(self.m, dtype=torch.int32, device=mask.device)
        self.slot = slot
        self.y = y.to(torch.int64).contiguous()
        self.selected_labels = self.y[rows]


class _MaskedNLL(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, info):
        from .. import kernels
        z = z.contiguous()
        ctx.save_for_backward(z)
        ctx.info = info
        out = kernels.masked_nll_forward(z, info.rows, info.y)
        loss, acc = out[0], out[1]
        ctx.mark_non_differentiable(acc)
        return loss, acc

    @staticmethod
    def backward(ctx, g_loss, _g_acc):
        (z,) = ctx.saved_tensors
        return _MaskedNLLGrad.apply(z, g_loss, ctx.info), None


class _MaskedNLLGrad(torch.autograd.Function):
    """dz = g (softmax(z) - onehot(y)) / m on the mask rows, 0 elsewhere; differentiable once more in z and g."""

    @staticmethod
    def forward(ctx, z, g_loss, info):
        from .. import kernels
        g_loss = g_loss.reshape(1).to(torch.float32).contiguous()
        ctx.save_for_backward(z, g_loss)
        ctx.info = info
        return kernels.masked_nll_grad(z, info.slot, info.y, info.m, g_loss)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, u):
        from .. import kernels
        z, g_loss = ctx.saved_tensors
        info = ctx.info
        want_z, want_g = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        out_z, out_g = kernels.masked_nll_grad_grad(u.contiguous(), z, info.slot, info.y, info.m, g_loss, want_z=want_z, want_g=want_g)
        return out_z, (out_g.sum().reshape(()) if want_g else None), None


FUSED_NLL = [True]


def fused_nll_ok(z: torch.Tensor) -> bool:
    return FUSED_NLL[0] and z.is_cuda and z.dtype == torch.float32 and z.dim() == 2 and 0 < z.shape[1] <= 128


def masked_nll(logits: torch.Tensor, info: MaskInfo):
    """(loss, accuracy) as 0-dim device tensors; loss is differentiable to second order in the logits."""
    return _MaskedNLL.apply(logits, info)
