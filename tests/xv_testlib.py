"""Shared helpers of the GPU parity tests: configs, oracle <-> engine plumbing, error metrics."""
import numpy as np
import torch

from oracle import xvector_oracle as O


def base_params(**kw):
    d = dict(seed=0, network_type="tdnn", last_layer_no_bn=False, last_layer_linear=True, feature_norm=False,
             pooling_type="statistics_pooling", embedding_node="tdnn6_dense", weight_l2_regularizer=1e-2,
             batchnorm_momentum=0.99, clip_gradient=False, clip_gradient_norm=3, use_nesterov=False,
             num_nodes_pooling_layer=1500)
    d.update(kw)
    return d


def head_params(loss_type, m=None):
    d = {}
    for pre, lt in (("asoftmax", "asoftmax"), ("amsoftmax", "additive_margin_softmax"),
                    ("arcsoftmax", "additive_angular_margin_softmax")):
        d[pre + "_lambda_min"] = 0
        d[pre + "_lambda_base"] = 1000
        d[pre + "_lambda_gamma"] = 1e-5
        d[pre + "_lambda_power"] = 5
    d["asoftmax_m"] = 4
    d["amsoftmax_m"] = 0.2
    d["arcsoftmax_m"] = 0.2
    if m is not None:
        key = {"asoftmax": "asoftmax_m", "additive_margin_softmax": "amsoftmax_m",
               "additive_angular_margin_softmax": "arcsoftmax_m"}.get(loss_type)
        if key:
            d[key] = m
    return d


def make_batch(B, T, D, C, seed=0):
    """Synthetic segments with per-utterance offset/scale (speaker-like variability).  Plain iid noise makes every
    utterance's pooled statistics nearly identical, so the utterance-level batch-norms divide by a vanishing
    between-utterance variance and amplify bf16 rounding noise by orders of magnitude (ill-conditioned test)."""
    g = torch.Generator().manual_seed(seed)
    m = torch.randn(B, 1, D, generator=g, dtype=torch.float32)
    s = 0.5 + torch.rand(B, 1, D, generator=g, dtype=torch.float32)
    x = m + s * torch.randn(B, T, D, generator=g, dtype=torch.float32)
    y = torch.randint(0, C, (B,), generator=g, dtype=torch.int32)
    return x, y


def rel_fro(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def min_cosine(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return float(torch.nn.functional.cosine_similarity(a, b, dim=-1).min())
