"""nn.Linear whose fp32 CUDA inference forward runs on the tensor cores at fp32 accuracy (``tome_linear_f32``: exact
three-way bf16 split, the plane products down to 2^-32 of the term, fp32 accumulation) instead of the library's CUDA-core SGEMM -- the reference benchmark
runs the models in fp32 with TF32 off (slowfast/utils/model_benchmark.py:21-45), and those SGEMMs are 83 % of the patched
VideoMAE step.  Same parameters, same state-dict keys; every other case (training, bf16, odd shapes, CPU) is nn.Linear."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class TomeLinear(nn.Linear):
    def forward(self, x):
        if not torch.is_tensor(x):          # _native.Planes: the previous exact-split kernel's output, already split
            from tome import _native
            return _native.linear_f32(x, self.weight, self.bias)
        if x.is_cuda and x.dtype == torch.float32 and not self.training:
            from tome import _native
            if _native.linear_f32_usable(x, self.weight, self.bias):
                return _native.linear_f32(x, self.weight, self.bias)
        return F.linear(x, self.weight, self.bias)


def install(model: nn.Module) -> nn.Module:
    """Class-swap every plain nn.Linear of a host model (no parameter is touched)."""
    for m in model.modules():
        if type(m) is nn.Linear:
            m.__class__ = TomeLinear
    return model


def linear(x, weight, bias=None, out="fp32"):
    """F.linear for the host models' explicit GEMMs (tubelet embeddings, q/v-biased QKV).  ``out="both"`` also returns the
    result's split planes (None when the library GEMM ran)."""
    if x.is_cuda and x.dtype == torch.float32:
        from tome import _native
        if _native.linear_f32_usable(x, weight, bias):
            return _native.linear_f32(x, weight, bias, out=out)
    y = F.linear(x, weight, bias)
    return (y, None) if out == "both" else y
