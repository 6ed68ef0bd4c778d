"""NetVLAD / GhostVLAD pooling with the reference's operator surface (model/pooling.py:195-277).

``ghost_vlad(features, aux_features, endpoints, params, is_training)`` reads its key / value inputs from ``endpoints``
like the original.  The key / value networks (``dense_bn_relu`` stacks, model/common.py:113-146) and the
``vlad_weight_affine`` layer are frame layers on the tcgen05 GEMM; posteriors, residual aggregation, the two L2
normalisations and all their gradients are the kernels of csrc/xv_vlad.cu (Engine.vlad_pool).
"""
import numpy as np

from .. import _lib as L
from ..runtime import VarSpec, get_engine, _pad_to
from .attention import _endpoint_dim, _run_stack

MAX_CLUSTERS = 64


def vlad_value_dim(params):
    nodes = list(params.vlad_value_num_nodes)
    return int(nodes[-1]) if nodes else _endpoint_dim(params.vlad_value_input, params)


def vlad_key_dim(params):
    nodes = list(params.vlad_key_num_nodes)
    return int(nodes[-1]) if nodes else _endpoint_dim(params.vlad_key_input, params)


def _layers(kind, in_dim, nodes):
    out, d = [], in_dim
    for i, n in enumerate(nodes):
        out.append(("vlad_%s%d" % (kind, i), d, int(n), True, "relu"))
        d = int(n)
    return out


def _value_layers(params):
    return _layers("value", _endpoint_dim(params.vlad_value_input, params), list(params.vlad_value_num_nodes))


def _key_layers(params):
    return _layers("key", _endpoint_dim(params.vlad_key_input, params), list(params.vlad_key_num_nodes))


def vlad_output_dim(params):
    """(real dim, padded dim, row map real -> padded) of the pooled vector = rows of tdnn6_dense/kernel: cluster k owns
    columns [k*dv, (k+1)*dv), each cluster padded to the value tensor's channel padding."""
    k, dv = int(params.vlad_num_centers), vlad_value_dim(params)
    dvp = _pad_to(dv, 64)
    rows = (np.arange(k)[:, None] * dvp + np.arange(dv)[None, :]).reshape(-1)
    return k * dv, k * dvp, rows


def declare_vlad_variables(engine, params):
    """Variables of the ``tdnn/vlad`` scope (pooling.py:225-258): the value / key dense_bn_relu stacks, the cluster
    assignment layer ``vlad_weight_affine`` and the centres ``vlad_centers`` [centers + ghosts, value dim] (xavier, L2)."""
    st = engine.store
    l2 = float(params.weight_l2_regularizer)
    prelu = params.dict.get("network_relu_type", "relu") == "prelu"
    kg = int(params.vlad_num_centers) + int(params.vlad_num_ghosts)
    if not (1 <= int(params.vlad_num_centers) <= kg <= MAX_CLUSTERS):
        raise NotImplementedError("ghost_vlad: 1 <= vlad_num_centers <= centers + ghosts <= %d" % MAX_CLUSTERS)
    for name, cin, cout, bn, act in _value_layers(params) + _key_layers(params):
        scope = "tdnn/vlad/%s/%s" % (name, name)
        cin_p, cout_p = _pad_to(cin, 64), _pad_to(cout, 64)
        st.declare(VarSpec(scope + "_dense/kernel", (cin, cout), (cin_p, cout_p), l2=l2, shadow="plain", init="glorot",
                           fans=(cin, cout)))
        st.declare(VarSpec(scope + "_dense/bias", (cout,), (cout_p,)))
        st.declare(VarSpec(scope + "_bn/gamma", (cout,), (cout_p,), init="ones"))
        st.declare(VarSpec(scope + "_bn/beta", (cout,), (cout_p,)))
        st.declare(VarSpec(scope + "_bn/moving_mean", (cout,), (cout_p,), trainable=False))
        st.declare(VarSpec(scope + "_bn/moving_variance", (cout,), (cout_p,), trainable=False, init="ones", pad_value=1.0))
        if prelu:
            st.declare(VarSpec(scope + "_relu/alpha", (cout,), (cout_p,), init=0.01))
    dk, dv = vlad_key_dim(params), vlad_value_dim(params)
    dkp, dvp, kgp = _pad_to(dk, 64), _pad_to(dv, 64), _pad_to(kg, 64)
    st.declare(VarSpec("tdnn/vlad/vlad_weight_affine/kernel", (dk, kg), (dkp, kgp), l2=l2, shadow="plain", init="glorot",
                       fans=(dk, kg)))
    st.declare(VarSpec("tdnn/vlad/vlad_weight_affine/bias", (kg,), (kgp,)))
    st.declare(VarSpec("tdnn/vlad/vlad_centers", (kg, dv), (kg, dvp), l2=l2, init="glorot", fans=(kg, dv)))


def ghost_vlad(features, aux_features, endpoints, params, is_training):
    """NetVLAD and GhostVLAD (model/pooling.py:195-277).

    Args:
        features: unused (the reference reads ``endpoints`` instead, pooling.py:226-227).
        aux_features: unused.
        endpoints: outputs of the frame layers; ``endpoints[params.vlad_key_input]`` / ``[params.vlad_value_input]`` are
                   the key / value sources.  Gains ``vlad_weights``, ``vlad_value``, ``vlad_key``, ``vlad_centers``.
        params: vlad_num_centers, vlad_num_ghosts, vlad_key_input, vlad_key_num_nodes, vlad_value_input,
                vlad_value_num_nodes, vlad_final_l2_norm (pooling.py:205-213).
        is_training: BN mode of the key / value nets; records the backward closures.
    :return: UttAct handle [batch, vlad_num_centers * value dim].
    """
    eng = get_engine()
    training = bool(is_training)
    value = endpoints[params.vlad_value_input]
    key = endpoints[params.vlad_key_input]
    vl = _value_layers(params)
    if vl:
        value = _run_stack(eng, value, vl, params, training, endpoints, root="tdnn/vlad", buf_prefix="vlad")
    kl = _key_layers(params)
    if kl:
        key = _run_stack(eng, key, kl, params, training, endpoints, root="tdnn/vlad", buf_prefix="vlad")
    k, g = int(params.vlad_num_centers), int(params.vlad_num_ghosts)
    _, logits = eng.frame_affine(key, "tdnn/vlad/vlad_weight_affine/kernel", "tdnn/vlad/vlad_weight_affine/bias", 1, k + g,
                                 "vlad/weight_affine", training, bn=None, act=L.ACT_NONE)
    u, post = eng.vlad_pool(logits, value, "tdnn/vlad/vlad_centers", k, g, bool(params.vlad_final_l2_norm), training)
    endpoints["vlad_weights"] = post
    endpoints["vlad_value"] = value
    endpoints["vlad_key"] = logits
    endpoints["vlad_centers"] = eng.store.view("tdnn/vlad/vlad_centers")
    return u
