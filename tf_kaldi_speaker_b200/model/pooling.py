"""Temporal pooling with the reference's operator surface (model/pooling.py:9-34, 37-192, 195-277).

statistics_pooling: per (segment, channel) mean and standard deviation over the valid frames in ONE pass over
the bf16 activations (shifted moments, fp32 accumulation), with the masked variable-length semantics of
model/multitask_v1/pooling.py:9-40 when the input carries per-row lengths.
"""
import numpy as np

from ..runtime import get_engine, _pad_to

VAR2STD_EPSILON = 1e-12


def pooling_output_dim(params):
    """(real dim, padded dim, row map real -> padded) of the pooling output = rows of tdnn6_dense/kernel."""
    P = int(params.dict.get("num_nodes_pooling_layer", 1500))
    if params.pooling_type == "statistics_pooling":
        Ppad = _pad_to(P, 64)
        rows = np.concatenate([np.arange(P), Ppad + np.arange(P)])
        return 2 * P, 2 * Ppad, rows
    if params.pooling_type == "self_attention":
        from .attention import attention_value_dim
        dv = attention_value_dim(params)
        dvp = _pad_to(dv, 64)
        rows = np.concatenate([np.arange(dv), dvp + np.arange(dv)])
        return 2 * dv, 2 * dvp, rows
    if params.pooling_type == "ghost_vlad":
        from .vlad import vlad_output_dim
        return vlad_output_dim(params)
    raise NotImplementedError("Not implement %s pooling" % params.pooling_type)


def statistics_pooling(features, aux_features, endpoints, params, is_training):
    """Statistics pooling (model/pooling.py:9-34).

    Args:
        features: FrameAct handle with logical shape [batch, length, dim] (the tdnn5_relu output).
        aux_features, endpoints, params: unused, kept for signature compatibility.
        is_training: record the backward closure.
    :return: UttAct handle [batch, 2*dim] = [mean, stddev].
    """
    return get_engine().stats_pool(features, bool(is_training))


def declare_attention_variables(engine, params):
    from .attention import declare_attention_variables as _decl
    return _decl(engine, params)


def self_attention(features, aux_features, endpoints, params, is_training=None):
    """Multi-head attentive statistics pooling (model/pooling.py:37-192); see model/attention.py."""
    from .attention import self_attention as _impl
    return _impl(features, aux_features, endpoints, params, is_training)


def declare_vlad_variables(engine, params):
    from .vlad import declare_vlad_variables as _decl
    return _decl(engine, params)


def ghost_vlad(features, aux_features, endpoints, params, is_training):
    """NetVLAD / GhostVLAD pooling (model/pooling.py:195-277); see model/vlad.py."""
    from .vlad import ghost_vlad as _impl
    return _impl(features, aux_features, endpoints, params, is_training)
