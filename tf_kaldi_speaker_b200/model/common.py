"""Small helpers of the operator layer (reference model/common.py)."""
from .. import _lib as L


def activation_id(params):
    """network_relu_type -> kernel activation id (model/tdnn.py:25-30, model/common.py:27-42)."""
    t = params.dict.get("network_relu_type", "relu")
    if t == "prelu":
        return L.ACT_PRELU
    if t == "lrelu":
        return L.ACT_LRELU
    return L.ACT_RELU


def l2_scaling(x, scaling_factor, epsilon=1e-12, name="l2_norm"):
    """model/common.py:45-58 on a torch tensor (inspection only: inside the training / extraction path the
    scaling is fused into the head's feature-preparation kernel, see Engine.margin_head)."""
    import torch
    sq = (x * x).sum(-1, keepdim=True)
    return x * (torch.rsqrt(torch.clamp(sq, min=epsilon)) * scaling_factor)
