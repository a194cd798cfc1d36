"""fp32 parity arithmetic for the library convolutions that remain on the fp32 path.

cuDNN reads ``torch.backends.cudnn.allow_tf32`` when a kernel is LAUNCHED, so a ``with cudnn.flags(...)`` around the
forward does not cover dgrad / wgrad, which autograd launches later.  ``fence_losses`` puts an identity node on every loss:
its backward (the first node of the backward pass) switches TF32 off and queues the restore as an engine callback that
runs when the whole backward pass has finished."""
import torch


class _Fp32Fence(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        prev_c, prev_m = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
        if prev_c or prev_m:
            torch.backends.cudnn.allow_tf32 = False
            torch.backends.cuda.matmul.allow_tf32 = False

            def restore():
                torch.backends.cudnn.allow_tf32 = prev_c
                torch.backends.cuda.matmul.allow_tf32 = prev_m
            torch.autograd.Variable._execution_engine.queue_callback(restore)
        return g


def fence_losses(out):
    """Wrap every differentiable tensor of a loss dict (or a tensor) in the fp32 fence."""
    if isinstance(out, dict):
        return {k: fence_losses(v) for k, v in out.items()}
    if isinstance(out, (list, tuple)):
        return type(out)(fence_losses(v) for v in out)
    if torch.is_tensor(out) and out.requires_grad:
        return _Fp32Fence.apply(out)
    return out


class fp32_exact:
    """Context for the forward: TF32 off for cuDNN and cuBLAS."""

    def __enter__(self):
        self.prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False

    def __exit__(self, *exc):
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = self.prev
