"""CPU restatement of the tf-kaldi-speaker x-vector hot path.  TEST INFRASTRUCTURE ONLY.

This module is the *checker* for the CUDA path, never the thing shipped or measured
(except as the labelled ``cpu_baseline`` / ``--impl reference`` arm of ``bench.py``).
It restates, in plain PyTorch-CPU (fp64 by default, fp32 for the timed CPU baseline),
the arithmetic of these reference files (paths relative to /root/reference):

  model/tdnn.py:33-191               network topology, variable names, defaults
  model/pooling.py:22-32             statistics_pooling
  model/multitask_v1/pooling.py:22-38  length-masked statistics pooling (statistics_pooling_v2)
  model/pooling.py:78-189            self_attention (multi-head attentive statistics + penalty)
  model/common.py:27-58,113-265      prelu, l2_scaling, dense_* helpers, split_heads
  model/loss.py:29-35,97-159,207-247,293-345   softmax / asoftmax / AM / AAM heads
  model/loss.py:985-1037             auxiliary ring loss / MHE (pinned: model/test_utils.py:855-884, tests/golden/aux.npz)
  model/trainer.py:168-188           entire_network (feature_norm -> l2_scaling)
  model/trainer.py:261-303           validation-time margin neutralisation
  model/trainer.py:328-358,403-436   optimizer, total loss, gradient clipping, BN update deps
  egs/voxceleb/v1/nnet/lib/extract.py:65-94    chunk-and-average extraction rule

Parity pinning status (see DESIGN.md "Oracle"):
  * margin heads (asoftmax m=1,2,4 / AM / AAM, feature_norm on/off): PINNED against the reference's own
    NumPy known-answer functions model/test_utils.py:157-318 on the reference's adversarial inputs
    (model/tdnn.py:254-343) -- tests/golden/make_golden.py imports them from /root/reference (build container only)
    and commits tests/golden/heads.npz; tests/test_oracle_golden.py checks this module against it.
  * length-masked statistics pooling: PINNED against model/multitask_v1/pooling.py:68-80 (the reference's inline
    ``compute_stat_pooling``, exec'd from the reference source text by make_golden.py; golden vectors committed).
  * ring loss / MHE: PINNED against model/test_utils.py:855-884 (tests/golden/aux.npz).
  * self-attention pooling: weakly pinned against model/test_utils.py:321-376 (py2 integer division fixed).
  * TDNN conv/dense/BN forward, unmasked statistics_pooling, all backward passes, optimizer and BN
    moving-stat updates, plain softmax head, extraction averaging: "parity unpinned" -- the reference
    holds no known-answer test for them and TensorFlow 1.x (un-vendored, unpinned; README.md:31-34) cannot
    run in this image.  They are restated here from the cited lines plus the TF-1.12 documented defaults
    (BN eps 1e-3, valid padding, glorot-uniform init, mean-reduced sparse softmax xent, l2_normalize eps
    1e-12, l2_regularizer = s*sum(w^2)/2, leaky_relu alpha 0.2).

Gradients: the reference never writes a backward pass (trainer.py:403 uses TF autodiff and its tests only
assert "not NaN"), so PyTorch autograd on this fp64 restatement is the gradient oracle.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

VAR2STD_EPSILON = 1e-12   # model/pooling.py:6
BN_EPSILON = 1e-3         # tf.layers.batch_normalization default


# --------------------------------------------------------------------------------------
# Params: stand-in for misc/utils.py:13-61 (the original imports TensorFlow at module top)
# --------------------------------------------------------------------------------------
class ParamsPlain(object):
    """Attribute bag with a ``.dict`` view, as misc/utils.py:44-61."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    @property
    def dict(self):
        return self.__dict__


# --------------------------------------------------------------------------------------
# Parameter schema / initialisation (TF variable names are the keys)
# --------------------------------------------------------------------------------------
def _glorot_uniform(shape, gen):
    """TF glorot_uniform / xavier_initializer(uniform=True): limit = sqrt(6/(fan_in+fan_out))."""
    if len(shape) == 2:
        fan_in, fan_out = shape
    else:  # conv kernel [1, k, Cin, Cout]
        rf = shape[0] * shape[1]
        fan_in, fan_out = rf * shape[2], rf * shape[3]
    limit = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * limit


def frame_layer_specs(dim, params):
    """[(name, kind, k, Cin, Cout)] for tdnn1..5 (model/tdnn.py:39-127)."""
    p = params.dict.get("num_nodes_pooling_layer", 1500)
    return [("tdnn1", "conv", 5, dim, 512), ("tdnn2", "conv", 5, 512, 512), ("tdnn3", "conv", 7, 512, 512),
            ("tdnn4", "dense", 1, 512, 512), ("tdnn5", "dense", 1, 512, p)]


def pooling_output_dim(params):
    p = params.dict.get("num_nodes_pooling_layer", 1500)
    if params.pooling_type == "statistics_pooling":
        return 2 * p
    if params.pooling_type == "self_attention":
        nodes = list(params.att_value_num_nodes)
        dv = nodes[-1] if len(nodes) > 0 else _endpoint_dim(params.att_value_input, params)
        return 2 * dv
    if params.pooling_type == "ghost_vlad":
        nodes = list(params.vlad_value_num_nodes)
        dv = nodes[-1] if len(nodes) > 0 else _endpoint_dim(params.vlad_value_input, params)
        return params.vlad_num_centers * dv
    raise NotImplementedError("Not implement %s pooling" % params.pooling_type)


def _endpoint_dim(name, params):
    p = params.dict.get("num_nodes_pooling_layer", 1500)
    return p if name.startswith("tdnn5") else 512


def init_params(dim, params, num_speakers=None, loss_type=None, seed=0, dtype=torch.float64):
    """Create every variable of the graph with TF's default initialisers (Appendix B)."""
    gen = torch.Generator().manual_seed(seed)
    P = OrderedDict()
    relu_type = params.dict.get("network_relu_type", "relu")

    def bn(prefix, c):
        P[prefix + "/gamma"] = torch.ones(c, dtype=torch.float64)
        P[prefix + "/beta"] = torch.zeros(c, dtype=torch.float64)
        P[prefix + "/moving_mean"] = torch.zeros(c, dtype=torch.float64)
        P[prefix + "/moving_variance"] = torch.ones(c, dtype=torch.float64)

    def alpha(prefix, c):
        if relu_type == "prelu":
            P[prefix + "/alpha"] = torch.full((c,), 0.01, dtype=torch.float64)

    for name, kind, k, cin, cout in frame_layer_specs(dim, params):
        if kind == "conv":
            P["tdnn/%s_conv/kernel" % name] = _glorot_uniform((1, k, cin, cout), gen)
            P["tdnn/%s_conv/bias" % name] = torch.zeros(cout, dtype=torch.float64)
        else:
            P["tdnn/%s_dense/kernel" % name] = _glorot_uniform((cin, cout), gen)
            P["tdnn/%s_dense/bias" % name] = torch.zeros(cout, dtype=torch.float64)
        bn("tdnn/%s_bn" % name, cout)
        alpha("tdnn/%s_relu" % name, cout)

    if params.pooling_type == "self_attention":
        def net(kind, in_dim, nodes, last_type):
            d = in_dim
            for i, n in enumerate(nodes):
                scope = "tdnn/attention/att_%s%d" % (kind, i)
                P["%s/att_%s%d_dense/kernel" % (scope, kind, i)] = _glorot_uniform((d, n), gen)
                P["%s/att_%s%d_dense/bias" % (scope, kind, i)] = torch.zeros(n, dtype=torch.float64)
                is_last = (i == len(nodes) - 1)
                if (not is_last) or last_type == 2:
                    bn("%s/att_%s%d_bn" % (scope, kind, i), n)
                if ((not is_last) or last_type in (1, 2)):
                    alpha("%s/att_%s%d_relu" % (scope, kind, i), n)
                d = n
            return d
        dk = net("key", _endpoint_dim(params.att_key_input, params), list(params.att_key_num_nodes),
                 params.att_key_network_type)
        dv = _endpoint_dim(params.att_value_input, params)
        if len(params.att_value_num_nodes) > 0:
            dv = net("value", dv, list(params.att_value_num_nodes), params.att_value_network_type)
        h = params.att_num_heads
        qd = dk // h if params.att_split_key else dk
        q = torch.empty(h, qd, dtype=torch.float64)
        torch.nn.init.trunc_normal_(q, mean=0.0, std=0.1, a=-0.2, b=0.2, generator=gen)
        P["tdnn/attention/query"] = q
        if params.dict.get("att_apply_nonlinear", False):
            bn("tdnn/attention/att_post_bn", 2 * dv)
            alpha("tdnn/attention/att_post_relu", 2 * dv)

    if params.pooling_type == "ghost_vlad":       # model/pooling.py:225-258
        def vnet(kind, in_dim, nodes):
            d = in_dim
            for i, n in enumerate(nodes):
                scope = "tdnn/vlad/vlad_%s%d" % (kind, i)
                P["%s/vlad_%s%d_dense/kernel" % (scope, kind, i)] = _glorot_uniform((d, n), gen)
                P["%s/vlad_%s%d_dense/bias" % (scope, kind, i)] = torch.zeros(n, dtype=torch.float64)
                bn("%s/vlad_%s%d_bn" % (scope, kind, i), n)
                alpha("%s/vlad_%s%d_relu" % (scope, kind, i), n)
                d = n
            return d
        dv = vnet("value", _endpoint_dim(params.vlad_value_input, params), list(params.vlad_value_num_nodes))
        dk = vnet("key", _endpoint_dim(params.vlad_key_input, params), list(params.vlad_key_num_nodes))
        kg = params.vlad_num_centers + params.vlad_num_ghosts
        P["tdnn/vlad/vlad_weight_affine/kernel"] = _glorot_uniform((dk, kg), gen)
        P["tdnn/vlad/vlad_weight_affine/bias"] = torch.zeros(kg, dtype=torch.float64)
        P["tdnn/vlad/vlad_centers"] = _glorot_uniform((kg, dv), gen)

    pool_dim = pooling_output_dim(params)
    P["tdnn/tdnn6_dense/kernel"] = _glorot_uniform((pool_dim, 512), gen)
    P["tdnn/tdnn6_dense/bias"] = torch.zeros(512, dtype=torch.float64)
    bn("tdnn/tdnn6_bn", 512)
    alpha("tdnn/tdnn6_relu", 512)
    e = params.dict.get("num_nodes_last_layer", 512)
    P["tdnn/tdnn7_dense/kernel"] = _glorot_uniform((512, e), gen)
    P["tdnn/tdnn7_dense/bias"] = torch.zeros(e, dtype=torch.float64)
    if not params.dict.get("last_layer_no_bn", False):
        bn("tdnn/tdnn7_bn", e)
    if not params.dict.get("last_layer_linear", False):
        alpha("tdnn/tdnn7_relu", e)

    if num_speakers is not None:
        P["softmax/output/kernel"] = _glorot_uniform((e, num_speakers), gen)
        if loss_type == "softmax":
            P["softmax/output/bias"] = torch.zeros(num_speakers, dtype=torch.float64)
        if "ring_loss" in params.dict.get("aux_loss_func", []):
            P["softmax_ringloss/r"] = torch.tensor(float(params.ring_loss_init), dtype=torch.float64)   # loss.py:1008-1011
    return OrderedDict((k, v.to(dtype)) for k, v in P.items())


def _average_centres(params, loss_type):
    """generalized_angular_triplet_loss with triplet_center = "average": the class centres are a non-trainable variable
    updated like batch-norm statistics, without a regulariser (model/loss.py:755-760)."""
    return (loss_type == "generalized_angular_triplet_loss" and params is not None
            and params.dict.get("triplet_center") == "average")


def trainable_names(P, params=None, loss_type=None):
    skip_w = _average_centres(params, loss_type)
    return [k for k in P if not (k.endswith("moving_mean") or k.endswith("moving_variance")
                                 or (skip_w and k == "softmax/output/kernel"))]


def l2_regularised(name):
    """L2 applies to kernels (model/tdnn.py:43, model/common.py:136, model/loss.py:33,102) and the VLAD centres
    (model/pooling.py:256-258)."""
    return name.endswith("/kernel") or name.endswith("/vlad_centers")


# --------------------------------------------------------------------------------------
# Building blocks
# --------------------------------------------------------------------------------------
def _activation(x, P, prefix, relu_type):
    """relu / prelu (model/common.py:27-42) / leaky_relu alpha=0.2 (model/tdnn.py:25-30)."""
    if relu_type == "prelu":
        a = P[prefix + "/alpha"]
        return F.relu(x) + a * (x - x.abs()) * 0.5
    if relu_type == "lrelu":
        return F.leaky_relu(x, 0.2)
    return F.relu(x)


def bf16_ste(x):
    """Round to bfloat16 storage with a straight-through gradient.  Used only by the ``emulate_bf16`` mode, which
    mimics WHERE the CUDA path stores bf16 (inputs, trunk kernels, pre-BN outputs, activations) so that tests can
    separate kernel bugs from the intrinsic rounding noise of a bf16 pipeline (ReLU-mask flips etc.)."""
    return x + (x.detach().to(torch.bfloat16).to(x.dtype) - x.detach())


def batch_norm(x, P, prefix, momentum, is_training, updates, unbiased_moving_var=False, store=None,
               stats_from_stored=False):
    """tf.layers.batch_normalization, axis=-1, eps=1e-3.  Training: biased batch statistics over all
    leading axes; moving <- moving*m + batch*(1-m).  The fused rank-4 TF path (tdnn1-3) updates
    moving_variance with the unbiased batch variance; the unfused rank-2/3 path with the biased one."""
    g, b = P[prefix + "/gamma"], P[prefix + "/beta"]
    if store is not None and stats_from_stored:   # short-K layers: statistics are taken from the stored bf16 tensor
        x = store(x)
        store = None
    if is_training:
        xr = x.reshape(-1, x.shape[-1])
        n = xr.shape[0]
        mean = xr.mean(0)
        var = ((xr - mean) ** 2).mean(0)
        if updates is not None:
            with torch.no_grad():
                mv = var * (n / max(n - 1, 1)) if unbiased_moving_var else var
                updates[prefix + "/moving_mean"] = P[prefix + "/moving_mean"] * momentum + mean * (1 - momentum)
                updates[prefix + "/moving_variance"] = P[prefix + "/moving_variance"] * momentum + mv * (1 - momentum)
    else:
        mean, var = P[prefix + "/moving_mean"], P[prefix + "/moving_variance"]
    if store is not None:      # statistics come from the fp32 accumulators, the normalised tensor from bf16 storage
        x = store(x)
    return (x - mean) * torch.rsqrt(var + BN_EPSILON) * g + b


def temporal_conv(x, kernel, bias):
    """tf.layers.conv2d(features[b,1,l,d], Cout, (1,k)) with 'valid' padding, stride 1 (tdnn.py:39-44)."""
    k, cin, cout = kernel.shape[1], kernel.shape[2], kernel.shape[3]
    w = kernel[0].permute(2, 1, 0)              # [Cout, Cin, k]
    y = F.conv1d(x.transpose(1, 2), w, bias)    # [B, Cout, T-k+1]
    return y.transpose(1, 2)


def statistics_pooling(x, lengths=None):
    """model/pooling.py:22-32; with ``lengths`` the masked form of multitask_v1/pooling.py:22-38."""
    if lengths is None:
        mean = x.mean(1, keepdim=True)
        var = ((x - mean) ** 2).mean(1)
        mean = mean.squeeze(1)
    else:
        t = torch.arange(x.shape[1]).unsqueeze(0)
        mask = (t < lengths.unsqueeze(1)).to(x.dtype).unsqueeze(2)
        flen = lengths.to(x.dtype).reshape(-1, 1, 1)
        mean = (x * mask).sum(1, keepdim=True) / (flen + 1e-16)
        var = (((x - mean) ** 2) * mask).sum(1) / (flen.squeeze(2) + 1e-16)
        mean = mean.squeeze(1)
    floor = (var <= VAR2STD_EPSILON).to(x.dtype)
    var = (1.0 - floor) * var + floor * VAR2STD_EPSILON
    return torch.cat([mean, torch.sqrt(var)], 1)


def _dense_stack(x, P, kind, nodes, last_type, params, is_training, endpoints, updates):
    """Key / value nets of self_attention (pooling.py:83-118; common.py:113-223)."""
    relu_type = params.dict.get("network_relu_type", "relu")
    for i, _ in enumerate(nodes):
        name = "att_%s%d" % (kind, i)
        scope = "tdnn/attention/" + name
        x = x @ P["%s/%s_dense/kernel" % (scope, name)] + P["%s/%s_dense/bias" % (scope, name)]
        endpoints["%s_dense" % name] = x
        last = (i == len(nodes) - 1)
        if (not last) or last_type == 2:
            x = batch_norm(x, P, "%s/%s_bn" % (scope, name), params.batchnorm_momentum, is_training, updates)
            endpoints["%s_bn" % name] = x
            x = _activation(x, P, "%s/%s_relu" % (scope, name), relu_type)
            endpoints["%s_relu" % name] = x
        elif last_type == 1:
            x = _activation(x, P, "%s/%s_relu" % (scope, name), relu_type)
            endpoints["%s_relu" % name] = x
        elif last_type == 3:
            x = torch.tanh(x)
            endpoints["%s_tanh" % name] = x
    return x


def self_attention(endpoints, P, params, is_training, updates, lengths=None):
    """model/pooling.py:78-189.  Returns (att [B, 2*dv], penalty scalar)."""
    relu_type = params.dict.get("network_relu_type", "relu")
    value = endpoints[params.att_value_input]
    key = endpoints[params.att_key_input]
    key = _dense_stack(key, P, "key", list(params.att_key_num_nodes), params.att_key_network_type,
                       params, is_training, endpoints, updates)
    if len(params.att_value_num_nodes) > 0:
        value = _dense_stack(value, P, "value", list(params.att_value_num_nodes),
                             params.att_value_network_type, params, is_training, endpoints, updates)
    h = params.att_num_heads
    b, l, dv = value.shape
    assert dv % h == 0, "The dim of the value must be divided by the num of heads."
    v = value.reshape(b, l, h, dv // h).permute(0, 2, 1, 3)            # split_heads: [B,H,L,dv/H]
    if params.att_split_key:
        assert key.shape[2] % h == 0
        k = key.reshape(b, l, h, key.shape[2] // h).permute(0, 2, 1, 3)
        e = torch.einsum("bhld,hd->blh", k, P["tdnn/attention/query"])
    else:
        k = key.unsqueeze(1)
        e = torch.einsum("bmld,hd->blh", k, P["tdnn/attention/query"])
    if params.att_use_scale:
        e = e * (float(k.shape[-1]) ** -0.5)
    e = e.permute(0, 2, 1)                                             # [B,H,L]
    if lengths is not None:   # no reference definition; extrapolated from the masked pooling rule
        t = torch.arange(l).reshape(1, 1, l)
        e = e.masked_fill(t >= lengths.reshape(-1, 1, 1), float("-inf"))
    w = torch.softmax(e, dim=-1)
    endpoints["attention_weights"] = w
    mean = torch.einsum("bhld,bhl->bhd", v, w)
    var = torch.einsum("bhld,bhl->bhd", (v - mean.unsqueeze(2)) ** 2, w)
    mean = mean.reshape(b, dv)
    var = var.reshape(b, dv)
    floor = (var <= VAR2STD_EPSILON).to(var.dtype)
    var = (1.0 - floor) * var + floor * VAR2STD_EPSILON
    att = torch.cat([mean, torch.sqrt(var)], 1)
    endpoints["att_output_before_nonlinear"] = att
    if params.dict.get("att_apply_nonlinear", False):
        att = batch_norm(att, P, "tdnn/attention/att_post_bn", params.batchnorm_momentum, is_training, updates)
        endpoints["att_post_bn"] = att
        att = _activation(att, P, "tdnn/attention/att_post_relu", relu_type)
        endpoints["att_post_relu"] = att
    pen = torch.einsum("ijk,ilk->ijl", w, w) - torch.eye(h, dtype=w.dtype).unsqueeze(0)
    penalty = params.att_penalty_term * (pen ** 2).sum() / float(b)
    return att, penalty


def _l2_normalize(x, dim, epsilon=1e-12):
    """tf.nn.l2_normalize: x * rsqrt(max(sum(x^2), epsilon))."""
    return x * torch.rsqrt(torch.clamp((x * x).sum(dim, keepdim=True), min=epsilon))


def ghost_vlad(endpoints, P, params, is_training, updates, lengths=None, store=None):
    """model/pooling.py:195-277 (NetVLAD / GhostVLAD).  ``lengths``: frames >= length carry no posterior mass (no reference
    definition; the masked-pooling rule of multitask_v1/pooling.py).  ``store``: storage rounding of the frame tensors
    (bf16 emulation)."""
    relu_type = params.dict.get("network_relu_type", "relu")
    q = store if store is not None else (lambda t: t)

    def stack(x, kind, nodes):
        for i, _ in enumerate(nodes):             # dense_bn_relu, model/common.py:113-146
            name = "vlad_%s%d" % (kind, i)
            scope = "tdnn/vlad/" + name
            x = x @ q(P["%s/%s_dense/kernel" % (scope, name)]) + P["%s/%s_dense/bias" % (scope, name)]
            endpoints["%s_dense" % name] = x
            x = batch_norm(x, P, "%s/%s_bn" % (scope, name), params.batchnorm_momentum, is_training, updates, store=store,
                           stats_from_stored=store is not None)
            endpoints["%s_bn" % name] = x
            x = q(_activation(x, P, "%s/%s_relu" % (scope, name), relu_type))
            endpoints["%s_relu" % name] = x
        return x

    value = stack(endpoints[params.vlad_value_input], "value", list(params.vlad_value_num_nodes))
    key = stack(endpoints[params.vlad_key_input], "key", list(params.vlad_key_num_nodes))
    key = q(key @ q(P["tdnn/vlad/vlad_weight_affine/kernel"]) + P["tdnn/vlad/vlad_weight_affine/bias"])
    A = torch.softmax(key, dim=-1)                                    # [B, L, K+G]
    if lengths is not None:
        t = torch.arange(A.shape[1]).reshape(1, -1, 1)
        A = A * (t < lengths.reshape(-1, 1, 1)).to(A.dtype)
    endpoints["vlad_weights"] = A
    centers = P["tdnn/vlad/vlad_centers"]
    res = torch.einsum("blk,bld->bkd", A, value) - A.sum(1).unsqueeze(2) * centers.unsqueeze(0)
    res = res[:, :params.vlad_num_centers, :]
    res = _l2_normalize(res, -1)
    out = res.reshape(res.shape[0], -1)
    if params.vlad_final_l2_norm:
        out = _l2_normalize(out, -1)
    endpoints["vlad_value"] = value
    endpoints["vlad_key"] = key
    endpoints["vlad_centers"] = centers
    return out


# --------------------------------------------------------------------------------------
# Network (model/tdnn.py:8-191) and entire_network (model/trainer.py:168-188)
# --------------------------------------------------------------------------------------
def tdnn(features, P, params, is_training=False, updates=None, lengths=None, mirror_tf_fused_bn=True,
         emulate_bf16=False):
    """Returns (features, endpoints, penalty).  ``lengths`` (input-domain frames per row) enables the
    masked pooling used for batched variable-length extraction; frame layers are unaffected because
    BN in inference mode is element-wise and valid output frames never read padded input frames."""
    relu_type = params.dict.get("network_relu_type", "relu")
    mom = params.batchnorm_momentum
    ep = OrderedDict()
    q = bf16_ste if emulate_bf16 else (lambda t: t)
    x = q(features)
    for name, kind, k, cin, cout in frame_layer_specs(features.shape[-1], params):
        if kind == "conv":
            x = temporal_conv(x, q(P["tdnn/%s_conv/kernel" % name]), P["tdnn/%s_conv/bias" % name])
            ep["%s_conv" % name] = x
        else:
            x = x @ q(P["tdnn/%s_dense/kernel" % name]) + P["tdnn/%s_dense/bias" % name]
            ep["%s_dense" % name] = x
        x = batch_norm(x, P, "tdnn/%s_bn" % name, mom, is_training, updates,
                       unbiased_moving_var=(mirror_tf_fused_bn and kind == "conv"),
                       store=bf16_ste if emulate_bf16 else None,
                       stats_from_stored=True)    # the GEMM epilogue takes the statistics of the stored bf16 tensor
        ep["%s_bn" % name] = x
        x = _activation(x, P, "tdnn/%s_relu" % name, relu_type)
        if not (name == "tdnn5" and params.pooling_type == "statistics_pooling"):
            x = q(x)      # the fused tdnn5 BN+ReLU+pooling kernel pools the fp32 activation, never storing it
        ep["%s_relu" % name] = x
    plen = None if lengths is None else (lengths - 14)
    penalty = None
    if params.pooling_type == "statistics_pooling":
        x = statistics_pooling(x, plen)
    elif params.pooling_type == "self_attention":
        x, penalty = self_attention(ep, P, params, is_training, updates, plen)
    elif params.pooling_type == "ghost_vlad":
        x = ghost_vlad(ep, P, params, is_training, updates, plen, store=bf16_ste if emulate_bf16 else None)
    else:
        raise NotImplementedError("Not implement %s pooling" % params.pooling_type)
    ep["pooling"] = x
    x = x @ P["tdnn/tdnn6_dense/kernel"] + P["tdnn/tdnn6_dense/bias"]
    ep["tdnn6_dense"] = x
    x = batch_norm(x, P, "tdnn/tdnn6_bn", mom, is_training, updates)
    ep["tdnn6_bn"] = x
    x = _activation(x, P, "tdnn/tdnn6_relu", relu_type)
    ep["tdnn6_relu"] = x
    x = x @ P["tdnn/tdnn7_dense/kernel"] + P["tdnn/tdnn7_dense/bias"]
    ep["tdnn7_dense"] = x
    if not params.dict.get("last_layer_no_bn", False):
        x = batch_norm(x, P, "tdnn/tdnn7_bn", mom, is_training, updates)
        ep["tdnn7_bn"] = x
    if not params.dict.get("last_layer_linear", False):
        x = _activation(x, P, "tdnn/tdnn7_relu", relu_type)
        ep["tdnn7_relu"] = x
    return x, ep, penalty


def l2_scaling(x, scaling_factor, epsilon=1e-12):
    """model/common.py:45-58."""
    sq = (x ** 2).sum(-1, keepdim=True)
    return x * (torch.rsqrt(torch.clamp(sq, min=epsilon)) * scaling_factor)


def entire_network(features, P, params, is_training=False, updates=None, lengths=None, emulate_bf16=False):
    """model/trainer.py:168-188."""
    x, ep, penalty = tdnn(features, P, params, is_training, updates, lengths, emulate_bf16=emulate_bf16)
    ep["output"] = x
    if params.dict.get("feature_norm", False):
        assert "feature_scaling_factor" in params.dict
        x = l2_scaling(x, params.feature_scaling_factor)
        ep["output"] = x
    return x, ep, penalty


# --------------------------------------------------------------------------------------
# Heads (model/loss.py)
# --------------------------------------------------------------------------------------
def margin_lambda(lmin, base, gamma, power, global_step):
    """loss.py:144-147 / 235-240 / 333-337."""
    lam = max(float(lmin), float(base) * (1.0 + float(gamma) * float(global_step)) ** (-float(power)))
    fa = 1.0 / (1.0 + lam)
    return lam, fa, 1.0 - fa


def _xent(logits, labels):
    """tf.losses.sparse_softmax_cross_entropy: mean over the batch of -log softmax(logits)[label]."""
    return F.cross_entropy(logits, labels.long(), reduction="mean")


def softmax_head(x, labels, P, params=None):
    """loss.py:29-35 (dense with bias + xent)."""
    logits = x @ P["softmax/output/kernel"] + P["softmax/output/bias"]
    return _xent(logits, labels), logits


def _margin_head(x, labels, w, phi_fn, fa, fs):
    eps = 1e-12
    wn = w * torch.rsqrt(torch.clamp((w ** 2).sum(0, keepdim=True), min=eps))     # l2_normalize(w, dim=0)
    logits = x @ wn
    idx = torch.arange(x.shape[0])
    sel = logits[idx, labels.long()]
    xnorm = torch.clamp(torch.linalg.norm(x, dim=1), min=eps)
    cos = torch.clamp(sel / xnorm, -1 + eps, 1 - eps)
    phi = phi_fn(cos)
    scaled = phi * xnorm
    delta = torch.zeros_like(logits)
    delta[idx, labels.long()] = scaled - sel
    logits_m = logits + delta
    updated = fs * logits + fa * logits_m
    return _xent(updated, labels), logits


def asoftmax_head(x, labels, P, params, global_step=None):
    """loss.py:97-159.  m=1 returns plain xent on ||x||cos(theta) (loss.py:110-115)."""
    w = P["softmax/output/kernel"]
    m = int(params.asoftmax_m)
    if m == 1:
        wn = w * torch.rsqrt(torch.clamp((w ** 2).sum(0, keepdim=True), min=1e-12))
        logits = x @ wn
        return _xent(logits, labels), logits
    if m == 2:
        def phi(c):
            return 2 * torch.sign(c) * c ** 2 - 1
    elif m == 4:
        def phi(c):
            c2, c4 = c ** 2, c ** 4
            s0 = torch.sign(c)
            s3 = torch.sign(2 * c2 - 1) * s0
            s4 = 2 * s0 + s3 - 3
            return s3 * (8 * c4 - 8 * c2 + 1) + s4
    else:
        raise NotImplementedError("[ERROR] m=%d is not unsupported." % m)
    gs = params.dict["global_step"] if global_step is None else global_step
    _, fa, fs = margin_lambda(params.asoftmax_lambda_min, params.asoftmax_lambda_base,
                              params.asoftmax_lambda_gamma, params.asoftmax_lambda_power, gs)
    return _margin_head(x, labels, w, phi, fa, fs)


def additive_margin_softmax_head(x, labels, P, params, global_step=None):
    """loss.py:207-247."""
    m = float(params.amsoftmax_m)
    gs = params.dict["global_step"] if global_step is None else global_step
    _, fa, fs = margin_lambda(params.amsoftmax_lambda_min, params.amsoftmax_lambda_base,
                              params.amsoftmax_lambda_gamma, params.amsoftmax_lambda_power, gs)
    return _margin_head(x, labels, P["softmax/output/kernel"], lambda c: c - m, fa, fs)


def additive_angular_margin_softmax_head(x, labels, P, params, global_step=None):
    """loss.py:293-345."""
    m = float(params.arcsoftmax_m)
    gs = params.dict["global_step"] if global_step is None else global_step
    _, fa, fs = margin_lambda(params.arcsoftmax_lambda_min, params.arcsoftmax_lambda_base,
                              params.arcsoftmax_lambda_gamma, params.arcsoftmax_lambda_power, gs)

    def phi(c):
        sin = torch.sqrt(torch.clamp(1 - c ** 2, min=1e-12))
        u = c * math.cos(m) - sin * math.sin(m)
        return torch.where(c > math.cos(math.pi - m), u, -u - 2)
    return _margin_head(x, labels, P["softmax/output/kernel"], phi, fa, fs)


HEADS = {
    "softmax": softmax_head,
    "asoftmax": asoftmax_head,
    "additive_margin_softmax": additive_margin_softmax_head,
    "additive_angular_margin_softmax": additive_angular_margin_softmax_head,
}


def aux_loss(x, labels, P, params):
    """model/loss.py:985-1037: ring loss (trainable radius ``softmax_ringloss/r``) and minimum hyperspherical energy of the
    normalised speaker matrix; pinned against model/test_utils.py:855-884 (tests/golden/aux.npz)."""
    total = 0.0
    for name in params.aux_loss_func:
        if name == "ring_loss":
            r = P["softmax_ringloss/r"]
            total = total + float(params.ring_loss_lambda) * ((torch.linalg.norm(x, dim=1) - r) ** 2).mean()
        elif name == "mhe_loss":
            w = P["softmax/output/kernel"]
            wn = w * torch.rsqrt(torch.clamp((w ** 2).sum(0, keepdim=True), min=1e-12))
            sel = wn.t()[labels.long()]
            total = total + float(params.mhe_lambda) * (1.0 / ((2.0 - 2.0 * (sel @ wn)).mean() + 1e-6))
        else:
            raise NotImplementedError("Unsupported loss function %s" % name)
    return total


# --------------------------------------------------------------------------------------
# Metric-learning losses on the embeddings (model/loss.py:358-705, model/common.py:61-110)
# --------------------------------------------------------------------------------------
def pairwise_euc_distances(x, squared=False):
    """model/common.py:61-93: ||a||^2 - 2<a,b> + ||b||^2 from the Gram matrix (its diagonal gives the norms), clamped at 0;
    the non-squared form masks exact zeros around the sqrt."""
    dot = x @ x.t()
    sq = torch.diagonal(dot)
    d = torch.clamp(sq.unsqueeze(1) - 2.0 * dot + sq.unsqueeze(0), min=0.0)
    if not squared:
        mask = (d == 0.0).to(d.dtype)
        d = torch.sqrt(d + mask * 1e-16) * (1.0 - mask)
    return d


def pairwise_cos_similarity(x, epsilon=1e-12):
    """model/common.py:96-110."""
    dot = x @ x.t()
    inv = torch.rsqrt(torch.clamp((x * x).sum(-1, keepdim=True), min=epsilon))
    return torch.clamp(dot * (inv @ inv.t()), -1.0, 1.0)


def _masked_maximum(data, mask, dim=1):
    m = data.amin(dim, keepdim=True)
    return ((data - m) * mask).amax(dim, keepdim=True) + m


def _masked_minimum(data, mask, dim=1):
    m = data.amax(dim, keepdim=True)
    return ((data - m) * mask).amin(dim, keepdim=True) + m


def semihard_triplet_loss(x, labels, params):
    """model/loss.py:358-498 (TF metric_learning triplet_semihard_loss on pairwise_euc_distances).  [x, i] indexes the
    (anchor, positive) pair; amin / amax split the gradient evenly among ties like tf.reduce_min / reduce_max."""
    b = x.shape[0]
    lab = labels.reshape(-1, 1)
    margin = float(params.margin)
    d = pairwise_euc_distances(x, bool(params.triplet_loss_squared))
    adjacency = lab == lab.t()
    adj_not = ~adjacency
    d_tile = d.repeat(b, 1)                                            # row i*b + x = d[x, :]
    mask = adj_not.repeat(b, 1) & (d_tile > d.t().reshape(-1, 1))     # d[x, y] > d[x, i] for a negative y
    mask_final = (mask.to(d.dtype).sum(1, keepdim=True) > 0.0).reshape(b, b).t()
    maskf = mask.to(d.dtype)
    negatives_outside = _masked_minimum(d_tile, maskf).reshape(b, b).t()
    negatives_inside = _masked_maximum(d, adj_not.to(d.dtype)).repeat(1, b)
    semi_hard = torch.where(mask_final, negatives_outside, negatives_inside)
    loss_mat = margin + d - semi_hard
    mask_pos = adjacency.to(d.dtype) - torch.eye(b, dtype=d.dtype)
    num_pos = torch.clamp(mask_pos.sum(), min=1e-16)
    return torch.clamp(loss_mat * mask_pos, min=0.0).sum() / num_pos


def _angular_positive(c, loss_type, margin):
    """d_p of model/loss.py:535-560.  The sqrt of the arc form is floored at 1e-12 so that masked-out entries with
    |cos| = 1 (the diagonal) give a finite derivative; the reference graph yields 0 * inf there."""
    if loss_type == "asoftmax":
        m = int(margin)
        if m == 1:
            return c
        if m == 2:
            return 2.0 * torch.sign(c) * c * c - 1.0
        if m == 4:
            c2, c4 = c * c, c ** 4
            s0 = torch.sign(c)
            s3 = torch.sign(2.0 * c2 - 1.0) * s0
            s4 = 2.0 * s0 + s3 - 3.0
            return s3 * (8.0 * c4 - 8.0 * c2 + 1.0) + s4
        raise NotImplementedError("[ERROR] m=%d is not unsupported if asoftmax is selected." % m)
    if loss_type == "additive_margin_softmax":
        return c - margin
    new = c * math.cos(margin) - torch.sqrt(torch.clamp(1.0 - c * c, min=1e-12)) * math.sin(margin)
    return torch.where(c <= math.cos(math.pi - margin), -new - 2.0, new)


def angular_triplet_loss(x, labels, params):
    """model/loss.py:501-634: online-mined triplet loss on pairwise cosines, d_p margin-transformed like the softmax
    heads; triplet_type "all" (mean over the violating triplets) or "hard" (hardest positive / negative per anchor)."""
    assert params.triplet_type in ("all", "hard")
    assert params.loss_type in ("asoftmax", "additive_margin_softmax", "additive_angular_margin_softmax")
    margin = float(params.margin)
    eps = 1e-12
    b = x.shape[0]
    c = pairwise_cos_similarity(x)
    pos = _angular_positive(c, params.loss_type, margin)
    neg = c
    eye = torch.eye(b, dtype=torch.bool)
    lab_eq = labels.reshape(1, -1) == labels.reshape(-1, 1)
    if params.triplet_type == "all":
        ne = ~eye
        distinct = ne.unsqueeze(2) & ne.unsqueeze(1) & ne.unsqueeze(0)
        valid = lab_eq.unsqueeze(2) & (~lab_eq).unsqueeze(1)
        mask = (distinct & valid).to(c.dtype)
        tl = torch.clamp(mask * (neg.unsqueeze(1) - pos.unsqueeze(2)), min=0.0)
        num_pos = (tl > eps).to(c.dtype).sum()
        return tl.sum() / (num_pos + 1e-16)
    map_ = (lab_eq & ~eye).to(c.dtype)
    ap = pos * map_ + pos.amax(1, keepdim=True) * (1.0 - map_)
    hardest_pos = ap.amin(1, keepdim=True)
    man = (~lab_eq).to(c.dtype)
    an = neg * man + pos.amin(1, keepdim=True) * (1.0 - man)          # (sic) the fill is the row minimum of d_p
    hardest_neg = an.amax(1, keepdim=True)
    return torch.clamp(hardest_neg - hardest_pos, min=0.0).mean()


def e2e_valid_loss(x, labels, params):
    """model/loss.py:637-705: softmax GE2E loss (scale 20, no bias) on speaker-ordered batches; a sample's own speaker is
    scored against the centre of the OTHER segments of that speaker."""
    n, m = int(params.num_valid_speakers_per_batch), int(params.num_valid_segments_per_speaker)
    f = l2_scaling(x, 1.0)
    dim = f.shape[1]
    fr = f.reshape(n, m, dim)
    center = l2_scaling(fr.mean(1), 1.0)
    center_ex = l2_scaling((fr.sum(1, keepdim=True) - fr).reshape(n * m, dim), 1.0)
    sim = f @ center.t()
    sim_ex = (f * center_ex).sum(1)
    own = torch.arange(n).repeat_interleave(m)
    mask = torch.zeros(n * m, n, dtype=f.dtype)
    mask[torch.arange(n * m), own] = 1.0
    sim = 20.0 * (sim * (1.0 - mask) + sim_ex.unsqueeze(1) * mask)
    return _xent(sim, own)


def generalized_angular_triplet_loss(x, labels, P, params, is_training=True, updates=None):
    """model/loss.py:708-901 (loss_compute = "raw"): triplet loss against class centres ``softmax/output/kernel`` [E, C] --
    trainable ("learnable") or a moving average of the class members updated before the distances are taken ("average",
    tf.assign: no gradient through the update).  Returns (loss, parts)."""
    assert params.triplet_center in ("learnable", "average")
    if params.loss_compute != "raw":
        raise NotImplementedError("Not implemented.")           # loss.py:826
    assert float(params.l2_loss_weight) == 0.0, "The weight decay is applied by regularization term, not the loss!"
    margin, target_margin = float(params.margin), float(params.target_margin)
    topn = int(params.triplet_topn)
    w = P["softmax/output/kernel"]
    lab = labels.long()
    fn = _l2_normalize(x, 1)
    w_upd = w
    if params.triplet_center == "average" and is_training:
        decay = 1.0 - float(params.triplet_center_momentum)
        wf = w.detach().t()
        delta = (wf[lab] - x.detach()) * decay
        w_upd = (wf - torch.zeros_like(wf).index_add_(0, lab, delta)).t()      # scatter_nd sums repeated labels
        if updates is not None:
            updates["softmax/output/kernel"] = w_upd
    wn = _l2_normalize(w_upd, 0)
    eps = 1e-12
    dist = (fn * fn).sum(1, keepdim=True) + (wn * wn).sum(0, keepdim=True) - 2.0 * (fn @ wn)     # sum((f - w)^2)
    b, c = dist.shape
    mask = torch.zeros_like(dist)
    mask[torch.arange(b), lab] = 1.0
    target = dist[torch.arange(b), lab]
    new_dist = dist * (1.0 - mask) + (dist.amax(1, keepdim=True) + dist) * mask
    tmask = (target > target_margin).to(dist.dtype)
    if topn == 1:
        tl = tmask * torch.clamp(margin + target - new_dist.amin(1), min=1e-16)
    elif topn == 0:
        tl = tmask.unsqueeze(1) * (torch.clamp(margin + target.unsqueeze(1) - new_dist, min=1e-16) * (1.0 - mask))
    else:
        nt = -torch.topk(-new_dist, topn, dim=1, sorted=False)[0]
        tl = tmask.unsqueeze(1) * torch.clamp(margin + target.unsqueeze(1) - nt, min=1e-16)
    triplet = tl.sum() / ((tl > eps).to(dist.dtype).sum() + eps)
    center = (tmask * target).sum() / (tmask.sum() + eps)
    between = -((1.0 - torch.eye(c, dtype=dist.dtype)) * (2.0 - 2.0 * (wn.t() @ wn))).sum() / (c * (c - 1))
    loss = (float(params.triplet_loss_weight) * triplet + float(params.center_loss_weight) * center
            + float(params.between_loss_weight) * between)
    return loss, {"triplet_loss": triplet, "center_loss": center, "between_loss": between, "average_centers": w_upd}


METRIC_LOSSES = {"semihard_triplet_loss": semihard_triplet_loss, "angular_triplet_loss": angular_triplet_loss}


def loss_network(loss_type, x, labels, P, params, global_step=None, is_training=True, updates=None):
    if loss_type == "generalized_angular_triplet_loss":
        return generalized_angular_triplet_loss(x, labels, P, params, is_training, updates)[0], None
    if loss_type in METRIC_LOSSES:
        return METRIC_LOSSES[loss_type](x, labels, params), None
    if loss_type == "e2e_valid_loss":
        return e2e_valid_loss(x, labels, params), None
    if loss_type not in HEADS:
        raise NotImplementedError("Not implement %s loss" % loss_type)
    if loss_type == "softmax":
        loss, logits = softmax_head(x, labels, P, params)
    else:
        loss, logits = HEADS[loss_type](x, labels, P, params, global_step)
        if loss_type == "asoftmax" and int(params.asoftmax_m) == 1:
            return loss, logits          # loss.py:110-115 returns before the auxiliary losses
    if "aux_loss_func" in params.dict and len(params.dict["aux_loss_func"]) > 0:
        loss = loss + aux_loss(x, labels, P, params)
    return loss, logits


# --------------------------------------------------------------------------------------
# Training step (model/trainer.py:328-358, 403-436)
# --------------------------------------------------------------------------------------
def regularization_loss(P, params):
    """tf.losses.get_regularization_loss(): sum over kernels of s*sum(w^2)/2."""
    s = float(params.weight_l2_regularizer)
    s_out = float(params.dict.get("output_weight_l2_regularizer", s))
    total = 0.0
    for k, v in P.items():
        if l2_regularised(k) and not (k == "softmax/output/kernel" and params.dict.get("triplet_center") == "average"
                                      and "triplet_topn" in params.dict):
            total = total + (s_out if k.startswith("softmax/") else s) * (v ** 2).sum() / 2
    return total


def forward_loss(P, features, labels, params, loss_type, global_step, is_training=True, updates=None,
                 emulate_bf16=False):
    x, ep, penalty = entire_network(features, P, params, is_training=is_training, updates=updates,
                                    emulate_bf16=emulate_bf16)
    loss, logits = loss_network(loss_type, x, labels, P, params, global_step, is_training, updates)
    total = loss + regularization_loss(P, params)
    if penalty is not None:
        total = total + penalty
    ep["logits"] = logits
    return loss, total, ep


def train_step(P, opt_state, features, labels, params, loss_type, learning_rate, global_step, emulate_bf16=False):
    """One sess.run(train_op).  Returns (raw_loss, total_loss, grads, new_P, new_opt_state, endpoints);
    endpoints["__raw_grads"] holds the gradients before clip_by_global_norm."""
    names = trainable_names(P, params, loss_type)
    Pg = OrderedDict((k, v.detach().clone().requires_grad_(k in names)) for k, v in P.items())
    updates = OrderedDict()
    loss, total, ep = forward_loss(Pg, features, labels, params, loss_type, global_step, True, updates, emulate_bf16)
    gl = torch.autograd.grad(total, [Pg[n] for n in names], allow_unused=True)
    grads = OrderedDict((n, (g if g is not None else torch.zeros_like(P[n]))) for n, g in zip(names, gl))
    ep["__raw_grads"] = grads
    if params.dict.get("clip_gradient", False):
        gn = torch.sqrt(sum((g ** 2).sum() for g in grads.values()))
        c = float(params.clip_gradient_norm)
        scale = c / torch.clamp(gn, min=c)
        grads = OrderedDict((n, g * scale) for n, g in grads.items())
    newP, new_state = apply_optimizer(P, grads, opt_state, params, learning_rate)
    for k, v in updates.items():
        newP[k] = v.detach()
    return loss.detach(), total.detach(), grads, newP, new_state, ep


def apply_optimizer(P, grads, state, params, lr):
    """tf.train.GradientDescentOptimizer / MomentumOptimizer / AdamOptimizer (trainer.py:328-347)."""
    opt = params.dict.get("optimizer", "sgd")
    state = dict(state) if state else {}
    newP = OrderedDict((k, v.detach().clone()) for k, v in P.items())
    if opt == "sgd":
        for n, g in grads.items():
            newP[n] = P[n].detach() - lr * g
    elif opt == "momentum":
        mu = float(params.momentum)
        nest = bool(params.dict.get("use_nesterov", False))
        for n, g in grads.items():
            acc = state.get(n, torch.zeros_like(g)) * mu + g
            state[n] = acc
            newP[n] = P[n].detach() - lr * ((g + mu * acc) if nest else acc)
    elif opt == "adam":
        b1, b2, eps = 0.9, 0.999, 1e-8
        t = state.get("__t", 0) + 1
        state["__t"] = t
        lr_t = lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t)
        for n, g in grads.items():
            m = state.get(n + "/m", torch.zeros_like(g)) * b1 + (1 - b1) * g
            v = state.get(n + "/v", torch.zeros_like(g)) * b2 + (1 - b2) * g * g
            state[n + "/m"], state[n + "/v"] = m, v
            newP[n] = P[n].detach() - lr_t * m / (torch.sqrt(v) + eps)
    else:
        raise SystemExit("Optimizer %s is not supported." % opt)
    return newP, state


def valid_params(params, loss_type):
    """Margin neutralisation for the validation graph (trainer.py:261-303)."""
    vp = ParamsPlain(**dict(params.dict))
    if loss_type == "asoftmax":
        vp.asoftmax_m = 1
    elif loss_type == "additive_margin_softmax":
        vp.amsoftmax_m = 0
    elif loss_type == "additive_angular_margin_softmax":
        vp.arcsoftmax_m = 0
    if "aux_loss_func" in vp.dict:
        vp.aux_loss_func = []          # trainer.py:279-282
    return vp


# --------------------------------------------------------------------------------------
# Extraction (egs/voxceleb/v1/nnet/lib/extract.py:65-94, model/trainer.py:708-726)
# --------------------------------------------------------------------------------------
def predict(features, P, params):
    """Trainer.predict: [T,D] -> [E] or [N,T,D] -> [N,E]; BN in inference mode."""
    with torch.no_grad():
        single = features.dim() == 2
        x = features.unsqueeze(0) if single else features
        _, ep, _ = entire_network(x, P, params, is_training=False)
        e = ep[params.embedding_node]
        return e[0] if single else e


def extract_embedding(feature, P, params, chunk_size=10000, min_chunk_size=25, normalize=False):
    """One utterance of extract.py's loop.  Returns None if skipped (too short)."""
    t = feature.shape[0]
    if t < min_chunk_size:
        return None
    if t > chunk_size:
        half = chunk_size // 2        # py2 integer division at extract.py:73,75
        n = int(np.ceil(float(t - chunk_size) / half)) + 1
        chunks, lens = [], []
        for i in range(n):
            start = i * half
            this = chunk_size if t - start > chunk_size else t - start
            lens.append(this)
            chunks.append(feature[start:start + this])
        embs = predict(torch.stack(chunks[:-1]), P, params)
        last = predict(chunks[-1], P, params)
        embs = torch.cat([embs, last.unsqueeze(0)], 0)
        if normalize:
            embs = embs / torch.sqrt((embs ** 2).sum(1, keepdim=True))
        ln = torch.tensor(lens, dtype=embs.dtype).unsqueeze(1)
        emb = (embs * ln).sum(0) / ln.sum()
    else:
        emb = predict(feature, P, params)
    if normalize:
        emb = emb / torch.sqrt((emb ** 2).sum())
    return emb


# --------------------------------------------------------------------------------------
# Algorithmic work
# --------------------------------------------------------------------------------------
def flops_fwd(T, D, C, pool_nodes=1500):
    return 2 * 512 * (5 * D * (T - 4) + 2560 * (T - 8) + (3584 + 512 + pool_nodes) * (T - 14)) \
        + 2 * (2 * pool_nodes * 512 + 512 * 512 + 512 * C)


def flops_train(T, D, C, pool_nodes=1500):
    return 3 * flops_fwd(T, D, C, pool_nodes) - 2 * 512 * 5 * D * (T - 4)
