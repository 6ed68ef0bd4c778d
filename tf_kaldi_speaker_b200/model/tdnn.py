"""x-vector network builder with the reference's operator surface (model/tdnn.py:8-191).

``tdnn(features, params, is_training, reuse_variables, aux_features)`` keeps the reference's signature, its
``params`` keys/defaults (it mutates ``params.dict`` exactly like the original) and the names of the ``endpoints``
it returns; instead of TF graph nodes it enqueues the sm_100a kernels on the current CUDA stream through the
Engine and returns device-tensor handles (``.dense()`` gives a plain fp32 torch tensor).

Every layer is affine -> BN -> ReLU ("bn+relu" order, tdnn.py:10); the three temporal convolutions of width
5/5/7 (tdnn.py:39,57,75) are implicit GEMMs over the flat-time activation matrix, the two per-frame dense layers
(tdnn.py:96,115) are the k=1 case of the same kernel.
"""
from collections import OrderedDict

import numpy as np

from .. import _lib as L
from ..runtime import VarSpec, get_engine, _pad_to
from .common import activation_id
from .pooling import (statistics_pooling, self_attention, ghost_vlad, declare_attention_variables, declare_vlad_variables,
                      pooling_output_dim)


def _bn_specs(store, prefix, c_real, c_pad):
    store.declare(VarSpec(prefix + "/gamma", (c_real,), (c_pad,), init="ones"))
    store.declare(VarSpec(prefix + "/beta", (c_real,), (c_pad,)))
    store.declare(VarSpec(prefix + "/moving_mean", (c_real,), (c_pad,), trainable=False))
    store.declare(VarSpec(prefix + "/moving_variance", (c_real,), (c_pad,), trainable=False, init="ones", pad_value=1.0))
    return (prefix + "/gamma", prefix + "/beta", prefix + "/moving_mean", prefix + "/moving_variance")


def input_padding(dim):
    """(dpad, ldo): per-tap channel padding and the K extent of the packed tdnn1 operand."""
    dpad = 32 if dim <= 32 else _pad_to(dim, 8)
    return dpad, _pad_to(5 * dpad, 64)


def declare_variables(engine, dim, params):
    """Create the variable schema of SURVEY Appendix B (TF names are the keys) in the engine's ParamStore."""
    st = engine.store
    l2 = float(params.weight_l2_regularizer)
    prelu = params.dict.get("network_relu_type", "relu") == "prelu"
    if "num_nodes_pooling_layer" not in params.dict:
        params.dict["num_nodes_pooling_layer"] = 1500          # tdnn.py:111-113
    if "num_nodes_last_layer" not in params.dict:
        params.dict["num_nodes_last_layer"] = 512               # tdnn.py:162-164
    P = int(params.num_nodes_pooling_layer)
    Ppad = _pad_to(P, 64)
    E = int(params.num_nodes_last_layer)
    dpad, ldo = input_padding(dim)

    def alpha(prefix, c_pad, c_real):
        if prelu:
            st.declare(VarSpec(prefix + "/alpha", (c_real,), (c_pad,), init=0.01))

    # tdnn1: TF kernel [1,5,D,512] -> rows j*dpad + c of the packed operand
    rm = (np.arange(5)[:, None] * dpad + np.arange(dim)[None, :]).reshape(-1)
    st.declare(VarSpec("tdnn/tdnn1_conv/kernel", (1, 5, dim, 512), (ldo, 512), row_map=rm, l2=l2, shadow="plain",
                       init="glorot", fans=(5 * dim, 5 * 512)))
    st.declare(VarSpec("tdnn/tdnn1_conv/bias", (512,), (512,)))
    for n, k in ((2, 5), (3, 7)):
        st.declare(VarSpec("tdnn/tdnn%d_conv/kernel" % n, (1, k, 512, 512), (k * 512, 512), l2=l2, shadow="plain",
                           init="glorot", fans=(k * 512, k * 512)))
        st.declare(VarSpec("tdnn/tdnn%d_conv/bias" % n, (512,), (512,)))
    st.declare(VarSpec("tdnn/tdnn4_dense/kernel", (512, 512), (512, 512), l2=l2, shadow="plain", init="glorot",
                       fans=(512, 512)))
    st.declare(VarSpec("tdnn/tdnn4_dense/bias", (512,), (512,)))
    st.declare(VarSpec("tdnn/tdnn5_dense/kernel", (512, P), (512, Ppad), l2=l2, shadow="plain", init="glorot",
                       fans=(512, P)))
    st.declare(VarSpec("tdnn/tdnn5_dense/bias", (P,), (Ppad,)))
    for n, (cr, cp) in zip(range(1, 6), ((512, 512),) * 4 + ((P, Ppad),)):
        _bn_specs(st, "tdnn/tdnn%d_bn" % n, cr, cp)
        alpha("tdnn/tdnn%d_relu" % n, cp, cr)

    if params.pooling_type == "self_attention":
        declare_attention_variables(engine, params)
    elif params.pooling_type == "ghost_vlad":
        declare_vlad_variables(engine, params)
    pool_real, pool_pad, pool_rows = pooling_output_dim(params)
    st.declare(VarSpec("tdnn/tdnn6_dense/kernel", (pool_real, 512), (pool_pad, 512), row_map=pool_rows, l2=l2,
                       shadow="split", init="glorot", fans=(pool_real, 512)))
    st.declare(VarSpec("tdnn/tdnn6_dense/bias", (512,), (512,)))
    _bn_specs(st, "tdnn/tdnn6_bn", 512, 512)
    alpha("tdnn/tdnn6_relu", 512, 512)
    st.declare(VarSpec("tdnn/tdnn7_dense/kernel", (512, E), (512, E), l2=l2, shadow="split", init="glorot",
                       fans=(512, E)))
    st.declare(VarSpec("tdnn/tdnn7_dense/bias", (E,), (E,)))
    if not params.dict.get("last_layer_no_bn", False):
        _bn_specs(st, "tdnn/tdnn7_bn", E, E)
    if not params.dict.get("last_layer_linear", False):
        alpha("tdnn/tdnn7_relu", E, E)


def _bn_names(prefix):
    return (prefix + "/gamma", prefix + "/beta", prefix + "/moving_mean", prefix + "/moving_variance")


def tdnn(features, params, is_training=None, reuse_variables=None, aux_features=None, lengths=None, ragged=None):
    """Build (= run) the TDNN.

    Args:
        features: fp32 CUDA tensor [batch, length, dim].
        params: configuration (nnet_conf JSON keys).
        is_training: True -> BN uses batch statistics and the backward tape is recorded.
        reuse_variables: kept for signature compatibility; variables live in the engine's name-keyed store.
        aux_features: unused by the TDNN (tdnn.py:8), accepted for compatibility.
        lengths: optional int tensor [batch] of valid input frames per row (batched variable-length extraction;
                 an extension -- the reference runs one utterance per call, extract.py:65-90).
        ragged: optional (starts, lengths) int32 device tensors: ``features`` is then ONE row [1, sum of lengths, dim]
                holding the utterances back to back (no padding); statistics pooling only, inference only.
    :return: (features, endpoints) -- output of the last layer and an OrderedDict of every component's output.
    """
    eng = get_engine()
    st = eng.store
    training = bool(is_training)
    act = activation_id(params)
    prelu = act == L.ACT_PRELU
    mom = float(params.batchnorm_momentum)
    endpoints = OrderedDict()
    if "num_nodes_pooling_layer" not in params.dict:
        params.dict["num_nodes_pooling_layer"] = 1500
    dim = features.shape[-1]
    dpad, _ = input_padding(dim)

    x = eng.pack_input(features, lengths=lengths, k=5, dpad=dpad)
    # Layers 1-3: temporal convolutions (tdnn.py:39-93); layers 4-5: per-frame dense (tdnn.py:96-127)
    for n, kind, k, cout in ((1, "conv", 1, 512), (2, "conv", 5, 512), (3, "conv", 7, 512), (4, "dense", 1, 512),
                             (5, "dense", 1, int(params.num_nodes_pooling_layer))):
        name = "tdnn%d" % n
        y, a = eng.frame_affine(x, "tdnn/%s_%s/kernel" % (name, kind), "tdnn/%s_%s/bias" % (name, kind), k, cout,
                                name, training, bn=_bn_names("tdnn/%s_bn" % name), act=act,
                                alpha=("tdnn/%s_relu/alpha" % name) if prelu else None,
                                unbiased_moving_var=(kind == "conv"), momentum=mom,
                                # tdnn5's BN + ReLU is fused into the statistics pooling (tdnn5_relu stays lazy)
                                defer_apply=(n == 5 and params.pooling_type == "statistics_pooling"))
        endpoints["%s_%s" % (name, kind)] = y
        endpoints["%s_bn" % name] = y            # y.affine = (scale, shift): BN output = y*scale + shift (lazy)
        endpoints["%s_relu" % name] = a
        x = a

    # Pooling layer (tdnn.py:133-143)
    if params.pooling_type == "statistics_pooling" and ragged is not None:
        u = eng.stats_pool(x, training, ragged=(ragged[0], ragged[1] - 14))      # pooled-domain lengths: T - 4 - 4 - 6
    elif params.pooling_type == "statistics_pooling":
        u = statistics_pooling(x, aux_features, endpoints, params, is_training)
    elif params.pooling_type == "self_attention":
        u = self_attention(x, aux_features, endpoints, params, is_training)
    elif params.pooling_type == "ghost_vlad":
        u = ghost_vlad(x, aux_features, endpoints, params, is_training)
    else:
        raise NotImplementedError("Not implement %s pooling" % params.pooling_type)
    endpoints["pooling"] = u
    eng.mark_utterance_level()

    # Utterance-level network (tdnn.py:147-189)
    y6, bn6, a6 = eng.utt_affine(u, "tdnn/tdnn6_dense/kernel", "tdnn/tdnn6_dense/bias", "tdnn6", training,
                                 bn=_bn_names("tdnn/tdnn6_bn"), act=act,
                                 alpha="tdnn/tdnn6_relu/alpha" if prelu else None, momentum=mom)
    endpoints["tdnn6_dense"] = y6
    endpoints["tdnn6_bn"] = bn6
    endpoints["tdnn6_relu"] = a6

    if "num_nodes_last_layer" not in params.dict:
        params.dict["num_nodes_last_layer"] = 512
    if "last_layer_no_bn" not in params.dict:
        params.last_layer_no_bn = False
    if "last_layer_linear" not in params.dict:
        params.last_layer_linear = False
    y7, bn7, a7 = eng.utt_affine(a6, "tdnn/tdnn7_dense/kernel", "tdnn/tdnn7_dense/bias", "tdnn7", training,
                                 bn=None if params.last_layer_no_bn else _bn_names("tdnn/tdnn7_bn"),
                                 act=L.ACT_NONE if params.last_layer_linear else act,
                                 alpha="tdnn/tdnn7_relu/alpha" if (prelu and not params.last_layer_linear) else None,
                                 momentum=mom)
    endpoints["tdnn7_dense"] = y7
    if not params.last_layer_no_bn:
        endpoints["tdnn7_bn"] = bn7
    if not params.last_layer_linear:
        endpoints["tdnn7_relu"] = a7
    return a7, endpoints
