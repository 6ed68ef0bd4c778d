"""Multi-head attentive statistics pooling (model/pooling.py:37-192) -- kernels land in a later milestone."""


def attention_value_dim(params):
    nodes = list(params.att_value_num_nodes)
    if len(nodes) > 0:
        return int(nodes[-1])
    p = int(params.dict.get("num_nodes_pooling_layer", 1500))
    return p if params.att_value_input.startswith("tdnn5") else 512


def declare_attention_variables(engine, params):
    raise NotImplementedError("self_attention pooling is not built yet in this round")


def self_attention(features, aux_features, endpoints, params, is_training=None):
    raise NotImplementedError("self_attention pooling is not built yet in this round")
