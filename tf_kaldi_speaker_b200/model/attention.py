"""Multi-head attentive statistics pooling with the reference's operator surface (model/pooling.py:37-192).

``self_attention(features, aux_features, endpoints, params, is_training)`` ignores ``features`` and reads its key /
value inputs from ``endpoints`` exactly like the original.  The key / value networks (dense_bn_relu / dense /
dense_relu / dense_tanh stacks, model/common.py:113-223) are frame layers on the tcgen05 GEMM; scores, softmax over
time, weighted mean / stddev, the multi-head penalty and all their gradients are the HBM-bound kernels of
csrc/xv_attention.cu (Engine.att_pool).
"""
from .. import _lib as L
from ..runtime import VarSpec, get_engine, _pad_to
from .common import activation_id

VAR2STD_EPSILON = 1e-12


def _endpoint_dim(name, params):
    p = int(params.dict.get("num_nodes_pooling_layer", 1500))
    if name.startswith("tdnn5"):
        return p
    if name.startswith("tdnn") and name[4] in "1234":
        return 512
    raise NotImplementedError("attention input %s is not a frame-level tdnn endpoint" % name)


def attention_value_dim(params):
    nodes = list(params.att_value_num_nodes)
    if len(nodes) > 0:
        return int(nodes[-1])
    return _endpoint_dim(params.att_value_input, params)


def attention_key_dim(params):
    nodes = list(params.att_key_num_nodes)
    assert len(nodes) > 0, "att_key_num_nodes must name at least one layer (pooling.py:83-99)"
    return int(nodes[-1])


def _stack_layers(kind, in_dim, nodes, last_type):
    """[(name, cin, cout, has_bn, act_kind)] of a key / value net; act_kind in {"none", "relu", "tanh"}."""
    out = []
    d = in_dim
    for i, n in enumerate(nodes):
        last = (i == len(nodes) - 1)
        if (not last) or last_type == 2:
            bn, act = True, "relu"
        elif last_type == 1:
            bn, act = False, "relu"
        elif last_type == 3:
            bn, act = False, "tanh"
        elif last_type == 0:
            bn, act = False, "none"
        else:
            raise NotImplementedError("att_%s_network_type %s" % (kind, last_type))
        out.append(("att_%s%d" % (kind, i), d, int(n), bn, act))
        d = int(n)
    return out


def _key_layers(params):
    return _stack_layers("key", _endpoint_dim(params.att_key_input, params), list(params.att_key_num_nodes),
                         int(params.att_key_network_type))


def _value_layers(params):
    nodes = list(params.att_value_num_nodes)
    if not nodes:
        return []
    return _stack_layers("value", _endpoint_dim(params.att_value_input, params), nodes, int(params.att_value_network_type))


def declare_attention_variables(engine, params):
    """Variables of the ``tdnn/attention`` scope: key / value dense stacks, the query
    [heads, key dim per head] ~ truncated_normal(0.1) (pooling.py:147-148) and the optional post BN."""
    import numpy as np
    st = engine.store
    l2 = float(params.weight_l2_regularizer)
    prelu = params.dict.get("network_relu_type", "relu") == "prelu"
    for name, cin, cout, bn, act in _key_layers(params) + _value_layers(params):
        scope = "tdnn/attention/%s/%s" % (name, name)
        cin_p, cout_p = _pad_to(cin, 64), _pad_to(cout, 64)
        st.declare(VarSpec(scope + "_dense/kernel", (cin, cout), (cin_p, cout_p), l2=l2, shadow="plain", init="glorot",
                           fans=(cin, cout)))
        st.declare(VarSpec(scope + "_dense/bias", (cout,), (cout_p,)))
        if bn:
            st.declare(VarSpec(scope + "_bn/gamma", (cout,), (cout_p,), init="ones"))
            st.declare(VarSpec(scope + "_bn/beta", (cout,), (cout_p,)))
            st.declare(VarSpec(scope + "_bn/moving_mean", (cout,), (cout_p,), trainable=False))
            st.declare(VarSpec(scope + "_bn/moving_variance", (cout,), (cout_p,), trainable=False, init="ones",
                               pad_value=1.0))
        if prelu and act == "relu":
            st.declare(VarSpec(scope + "_relu/alpha", (cout,), (cout_p,), init=0.01))
    h = int(params.att_num_heads)
    dk = attention_key_dim(params)
    if params.att_split_key:
        assert dk % h == 0
    dq = dk // h if params.att_split_key else dk
    st.declare(VarSpec("tdnn/attention/query", (h, dq), (h, dq), init="trunc_normal"))
    if params.dict.get("att_apply_nonlinear", False):
        dv = attention_value_dim(params)
        dvp = _pad_to(dv, 64)
        idx = np.concatenate([np.arange(dv), dvp + np.arange(dv)])
        pre = "tdnn/attention/att_post_bn"
        st.declare(VarSpec(pre + "/gamma", (2 * dv,), (2 * dvp,), row_map=idx, init="ones"))
        st.declare(VarSpec(pre + "/beta", (2 * dv,), (2 * dvp,), row_map=idx))
        st.declare(VarSpec(pre + "/moving_mean", (2 * dv,), (2 * dvp,), row_map=idx, trainable=False))
        st.declare(VarSpec(pre + "/moving_variance", (2 * dv,), (2 * dvp,), row_map=idx, trainable=False, init="ones",
                           pad_value=1.0))
        if prelu:
            st.declare(VarSpec("tdnn/attention/att_post_relu/alpha", (2 * dv,), (2 * dvp,), row_map=idx, init=0.01))


def _run_stack(eng, x, layers, params, training, endpoints, root="tdnn/attention", buf_prefix="att"):
    relu_act = activation_id(params)
    prelu = relu_act == L.ACT_PRELU
    mom = float(params.batchnorm_momentum)
    for name, cin, cout, bn, act in layers:
        scope = "%s/%s/%s" % (root, name, name)
        act_id = {"none": L.ACT_NONE, "tanh": L.ACT_TANH, "relu": relu_act}[act]
        bn_names = None
        if bn:
            bn_names = tuple(scope + "_bn/" + s for s in ("gamma", "beta", "moving_mean", "moving_variance"))
        y, a = eng.frame_affine(x, scope + "_dense/kernel", scope + "_dense/bias", 1, cout, buf_prefix + "/" + name, training,
                                bn=bn_names, act=act_id,
                                alpha=(scope + "_relu/alpha") if (prelu and act == "relu") else None,
                                unbiased_moving_var=False, momentum=mom)
        endpoints["%s_dense" % name] = y
        if bn:
            endpoints["%s_bn" % name] = y
        if act == "relu":
            endpoints["%s_relu" % name] = a
        elif act == "tanh":
            endpoints["%s_tanh" % name] = a
        x = a
    return x


def self_attention(features, aux_features, endpoints, params, is_training=None):
    """Self-attention pooling (model/pooling.py:37-192).

    Args:
        features: unused (the reference ignores it too, pooling.py:44-45).
        aux_features: unused.
        endpoints: outputs of the frame layers; ``endpoints[params.att_key_input]`` / ``[params.att_value_input]``
                   are the key / value sources.  Gains ``attention_weights`` ([batch, heads, length] view) and
                   ``att_output_before_nonlinear`` like the reference (pooling.py:160,172).
        params: att_* keys as documented at pooling.py:52-70.
        is_training: BN mode of the key / value nets and of the optional post BN; records the backward closures.
    :return: UttAct handle [batch, 2 * value dim] = [weighted mean, weighted stddev].
    """
    eng = get_engine()
    training = bool(is_training)
    value = endpoints[params.att_value_input]
    key = endpoints[params.att_key_input]
    key = _run_stack(eng, key, _key_layers(params), params, training, endpoints)
    vl = _value_layers(params)
    if vl:
        value = _run_stack(eng, value, vl, params, training, endpoints)
    n_heads = int(params.att_num_heads)
    assert value.C % n_heads == 0, "The dim of the value must be divided by the num of heads."
    if params.att_split_key:
        assert key.C % n_heads == 0
    coef = float(params.dict.get("att_penalty_term", 0.0))
    u, weights = eng.att_pool(key, value, "tdnn/attention/query", n_heads, bool(params.att_split_key),
                              bool(params.att_use_scale), coef, training)
    endpoints["attention_weights"] = weights
    endpoints["att_output_before_nonlinear"] = u
    if params.dict.get("att_apply_nonlinear", False):
        act = activation_id(params)
        pre = "tdnn/attention/att_post_bn"
        bn_u, u = eng.utt_bn_act(u, tuple(pre + "/" + s for s in ("gamma", "beta", "moving_mean", "moving_variance")),
                                 "att_post", training, act=act,
                                 alpha="tdnn/attention/att_post_relu/alpha" if act == L.ACT_PRELU else None,
                                 momentum=float(params.batchnorm_momentum))
        endpoints["att_post_bn"] = bn_u
        endpoints["att_post_relu"] = u
    return u
