"""Classification heads (model/loss.py:9-355) and metric-learning losses (model/loss.py:358-705) with the reference's
operator surface.

softmax / asoftmax / additive_margin_softmax / additive_angular_margin_softmax keep the reference signatures
``f(features, labels, num_outputs, params, is_training=None, reuse_variables=None, name="softmax")
-> (loss, endpoints)`` and parameter keys.  All four run ONE fused kernel sequence (Engine.margin_head): column
normalisation of the speaker matrix, row normalisation/scaling of the embeddings, the cosine-logit GEMM on
tcgen05 with the margin transform and an online log-sum-exp in its epilogue -- the [batch, speakers] logits
matrix is never written to HBM unless ``params.dict["debug_logits"]`` asks for ``endpoints["logits"]``.
"""
from collections import OrderedDict

import torch

from .. import _lib as L
from ..runtime import VarSpec, get_engine, _pad_to


def declare_head_variables(engine, embedding_dim, num_outputs, params, loss_type, name="softmax"):
    """softmax/output/kernel [E, C] (+ bias for the plain softmax), xavier-uniform, L2-regularised
    (loss.py:30-34, 100-102, 208-210, 294-296)."""
    l2 = float(params.weight_l2_regularizer)
    if "output_weight_l2_regularizer" in params.dict:
        l2 = float(params.output_weight_l2_regularizer)
    st = engine.store
    sh = engine.head_shard
    if sh is not None:
        # class-sharded head: this rank declares only its columns [lo, hi); initialisation draws the full [E, C] matrix
        # from the shared seed and slices it, so the sharded and the replicated model start from the same weights
        assert sh.num_outputs == num_outputs
        n_loc, cpad = sh.n_local, _pad_to(sh.n_local, 8)
        st.declare(VarSpec(name + "/output/kernel", (embedding_dim, n_loc), (embedding_dim, cpad), l2=l2, init="glorot",
                           fans=(embedding_dim, num_outputs), full_shape=(embedding_dim, num_outputs),
                           col_range=(sh.lo, sh.hi)))
        if loss_type == "softmax":
            st.declare(VarSpec(name + "/output/bias", (n_loc,), (cpad,), full_shape=(num_outputs,),
                               col_range=(sh.lo, sh.hi)))
        return
    cpad = _pad_to(num_outputs, 8)
    if loss_type == "generalized_angular_triplet_loss" and params.triplet_center == "average":
        # class centres updated on the fly like batch-norm statistics: not trainable, no regulariser (loss.py:755-760)
        st.declare(VarSpec(name + "/output/kernel", (embedding_dim, num_outputs), (embedding_dim, cpad), l2=0.0,
                           init="glorot", fans=(embedding_dim, num_outputs), trainable=False))
        return
    st.declare(VarSpec(name + "/output/kernel", (embedding_dim, num_outputs), (embedding_dim, cpad), l2=l2,
                       init="glorot", fans=(embedding_dim, num_outputs)))
    if loss_type == "softmax":
        st.declare(VarSpec(name + "/output/bias", (num_outputs,), (cpad,)))
    if "ring_loss" in params.dict.get("aux_loss_func", []):
        # the trainable ring radius, tf.get_variable("r", initializer=ring_loss_init) under <name>_ringloss (loss.py:1008-1011)
        st.declare(VarSpec(name + "_ringloss/r", (), (1,), init=float(params.ring_loss_init)))


def margin_lambda(lambda_min, lambda_base, lambda_gamma, lambda_power, global_step):
    """lambda = max(lambda_min, lambda_base * (1 + gamma*step)^(-power)); fa = 1/(1+lambda); fs = 1-fa
    (loss.py:144-147, 235-240, 333-337)."""
    lam = max(float(lambda_min), float(lambda_base) * (1.0 + float(lambda_gamma) * float(global_step)) ** (-float(lambda_power)))
    fa = 1.0 / (1.0 + lam)
    return lam, fa, 1.0 - fa


def margin_schedule(loss_type, params, global_step):
    """(fa, fs) that the head named ``loss_type`` will use at ``global_step`` (validation-neutralised margins excepted)."""
    pre = {"asoftmax": "asoftmax", "additive_margin_softmax": "amsoftmax",
           "additive_angular_margin_softmax": "arcsoftmax"}.get(loss_type)
    if pre is None or (loss_type == "asoftmax" and int(params.asoftmax_m) == 1):
        return (0.0, 1.0) if pre is not None else (1.0, 0.0)
    d = params.dict
    _, fa, fs = margin_lambda(d[pre + "_lambda_min"], d[pre + "_lambda_base"], d[pre + "_lambda_gamma"],
                              d[pre + "_lambda_power"], global_step)
    return fa, fs


def _aux_config(params, name, with_aux):
    """params.aux_loss_func (model/loss.py:985-1037) -> what Engine.margin_head fuses: ring loss on the head's input
    features with the trainable radius ``<name>_ringloss/r`` (loss.py:1008-1012) and MHE on the normalised weights."""
    if not with_aux or "aux_loss_func" not in params.dict or len(params.dict["aux_loss_func"]) == 0:
        return None
    aux = {}
    for loss_func in params.aux_loss_func:
        if loss_func == "ring_loss":
            aux["ring"] = (name + "_ringloss/r", float(params.ring_loss_lambda))
        elif loss_func == "mhe_loss":
            aux["mhe"] = float(params.mhe_lambda)
        else:
            raise NotImplementedError("Unsupported loss function %s" % loss_func)
    return aux


def _run_head(features, labels, num_outputs, params, is_training, name, head_type, margin=0.0, asoftmax_m=1,
              fa=1.0, fs=0.0, with_aux=True):
    eng = get_engine()
    assert features.data.dim() == labels.dim() + 1
    eng.set_sched(fa, fs)
    scaling = float(getattr(features, "scaling", 0.0) or 0.0)
    bias = (name + "/output/bias") if head_type == L.HEAD_SOFTMAX and (name + "/output/bias") in eng.store else None
    want_logits = bool(params.dict.get("debug_logits", False))
    aux = _aux_config(params, name, with_aux)
    if eng.head_shard is not None:
        if want_logits or aux:
            raise NotImplementedError("debug_logits / aux_loss_func are not available with the class-sharded head")
        loss, logits, x = eng.margin_head_sharded(features, labels, name + "/output/kernel", bias, head_type, num_outputs,
                                                  bool(is_training), margin=margin, asoftmax_m=asoftmax_m, scaling=scaling)
    else:
        loss, logits, x = eng.margin_head(features, labels, name + "/output/kernel", bias, head_type, num_outputs,
                                          bool(is_training), margin=margin, asoftmax_m=asoftmax_m, scaling=scaling,
                                          want_logits=want_logits, aux=aux)
    params.dict["softmax_w"] = eng.store.view(name + "/output/kernel")      # loss.py:103,211,297
    endpoints = OrderedDict()
    endpoints["logits"] = None if logits is None else logits[:, :num_outputs]
    endpoints["labels"] = labels
    return loss, endpoints


def softmax(features, labels, num_outputs, params, is_training=None, reuse_variables=None, name="softmax"):
    """Vanilla softmax loss: dense(num_outputs) with bias + mean cross entropy (loss.py:9-48)."""
    return _run_head(features, labels, num_outputs, params, is_training, name, L.HEAD_SOFTMAX)


def asoftmax(features, labels, num_outputs, params, is_training=None, reuse_variables=None, name="softmax"):
    """Angular softmax, m in {1, 2, 4} with the lambda annealing (loss.py:51-169)."""
    params.asoftmax_lambda_min = float(params.asoftmax_lambda_min)
    params.asoftmax_lambda_base = float(params.asoftmax_lambda_base)
    params.asoftmax_lambda_gamma = float(params.asoftmax_lambda_gamma)
    params.asoftmax_lambda_power = float(params.asoftmax_lambda_power)
    m = int(params.asoftmax_m)
    if m == 1:
        # plain xent on ||x|| cos(theta): no margin, no lambda (loss.py:110-115)
        # = the additive-margin epilogue with m = 0, fa = 0, fs = 1 (normalised weights, untouched target logit)
        return _run_head(features, labels, num_outputs, params, is_training, name, L.HEAD_AM, margin=0.0, fa=0.0, fs=1.0,
                         with_aux=False)       # loss.py:110-115 returns before the auxiliary losses
    if m not in (2, 4):
        raise NotImplementedError("[ERROR] m=%d is not unsupported." % m)
    _, fa, fs = margin_lambda(params.asoftmax_lambda_min, params.asoftmax_lambda_base, params.asoftmax_lambda_gamma,
                              params.asoftmax_lambda_power, params.dict["global_step"])
    return _run_head(features, labels, num_outputs, params, is_training, name, L.HEAD_ASOFTMAX, asoftmax_m=m, fa=fa, fs=fs)


def additive_margin_softmax(features, labels, num_outputs, params, is_training=None, reuse_variables=None, name="softmax"):
    """Additive margin softmax, phi = cos(theta) - m (loss.py:172-257)."""
    params.amsoftmax_lambda_min = float(params.amsoftmax_lambda_min)
    params.amsoftmax_lambda_base = float(params.amsoftmax_lambda_base)
    params.amsoftmax_lambda_gamma = float(params.amsoftmax_lambda_gamma)
    params.amsoftmax_lambda_power = float(params.amsoftmax_lambda_power)
    params.amsoftmax_m = float(params.amsoftmax_m)
    _, fa, fs = margin_lambda(params.amsoftmax_lambda_min, params.amsoftmax_lambda_base, params.amsoftmax_lambda_gamma,
                              params.amsoftmax_lambda_power, params.dict["global_step"])
    return _run_head(features, labels, num_outputs, params, is_training, name, L.HEAD_AM, margin=params.amsoftmax_m,
                     fa=fa, fs=fs)


def additive_angular_margin_softmax(features, labels, num_outputs, params, is_training=None, reuse_variables=None, name="softmax"):
    """Additive angular margin softmax (ArcFace), phi = cos(theta + m) with the monotone extension (loss.py:260-355)."""
    params.arcsoftmax_lambda_min = float(params.arcsoftmax_lambda_min)
    params.arcsoftmax_lambda_base = float(params.arcsoftmax_lambda_base)
    params.arcsoftmax_lambda_gamma = float(params.arcsoftmax_lambda_gamma)
    params.arcsoftmax_lambda_power = float(params.arcsoftmax_lambda_power)
    params.arcsoftmax_m = float(params.arcsoftmax_m)
    _, fa, fs = margin_lambda(params.arcsoftmax_lambda_min, params.arcsoftmax_lambda_base,
                              params.arcsoftmax_lambda_gamma, params.arcsoftmax_lambda_power, params.dict["global_step"])
    return _run_head(features, labels, num_outputs, params, is_training, name, L.HEAD_AAM, margin=params.arcsoftmax_m,
                     fa=fa, fs=fs)


# ---------------------------------------------------------------------------------------------------------------------
# Metric-learning losses on the embeddings (model/loss.py:358-705).  No speaker matrix: the batches come from the
# speakers x segments sampler (KaldiDataRandomQueue, params.num_speakers_per_batch / num_segments_per_speaker).
# ---------------------------------------------------------------------------------------------------------------------
METRIC_LOSSES = ("semihard_triplet_loss", "angular_triplet_loss", "e2e_valid_loss")
_ANGULAR_KINDS = {"asoftmax": 0, "additive_margin_softmax": 1, "additive_angular_margin_softmax": 2}


def _metric_endpoints(loss, labels):
    endpoints = OrderedDict()
    endpoints["loss"] = loss
    endpoints["labels"] = labels
    return endpoints


def semihard_triplet_loss(features, labels, num_outputs, params, is_training=None, reuse_variables=None, name="triplet_loss"):
    """Triplet loss with semi-hard negative mining (model/loss.py:358-498; TF metric_learning.triplet_semihard_loss).

    Args:
        features: [batch, dim] handle; the L2 normalisation is expected to have been applied (feature_norm).
        labels: int tensor [batch].
        num_outputs, reuse_variables, name: unused, kept for signature compatibility.
        params: params.margin, params.triplet_loss_squared.
        is_training: record the backward closure.
    :return: (loss, endpoints)
    """
    eng = get_engine()
    assert features.data.dim() == labels.dim() + 1
    loss, _ = eng.metric_loss(features, labels, "semihard", bool(is_training),
                              scaling=float(getattr(features, "scaling", 0.0) or 0.0), margin=float(params.margin),
                              squared=bool(params.triplet_loss_squared))
    return loss, _metric_endpoints(loss, labels)


def angular_triplet_loss(features, labels, num_outputs, params, is_training=None, reuse_variables=None,
                         name="angular_triplet_loss"):
    """Online-mined triplet loss on pairwise cosines (model/loss.py:501-634).

    Args:
        features: [batch, dim] handle.
        labels: int tensor [batch].
        params: params.margin; params.triplet_type "all" | "hard"; params.loss_type "asoftmax" |
                "additive_margin_softmax" | "additive_angular_margin_softmax" (the transform of the positive similarity).
        is_training: record the backward closure.
    :return: (loss, endpoints)
    """
    eng = get_engine()
    assert features.data.dim() == labels.dim() + 1
    assert params.triplet_type == "all" or params.triplet_type == "hard"
    assert params.loss_type in ["asoftmax", "additive_margin_softmax", "additive_angular_margin_softmax"]
    params.margin = float(params.margin)
    loss, _ = eng.metric_loss(features, labels, "angular", bool(is_training),
                              scaling=float(getattr(features, "scaling", 0.0) or 0.0), margin=params.margin,
                              angular_kind=_ANGULAR_KINDS[params.loss_type], hard=(params.triplet_type == "hard"))
    return loss, _metric_endpoints(loss, labels)


def e2e_valid_loss(features, labels, num_outputs, params, is_training=None, reuse_variables=None, name="valid_e2e_loss"):
    """Softmax generalized end-to-end loss for the validation set (model/loss.py:637-705); forward only.  The rows must be
    speaker-ordered: [s1, s1, ..., s2, s2, ...] with params.num_valid_speakers_per_batch x
    params.num_valid_segments_per_speaker rows."""
    assert "num_valid_speakers_per_batch" in params.dict and "num_valid_segments_per_speaker" in params.dict, \
        "Valid parameters should be set if E2E loss is selected"
    if is_training:
        raise NotImplementedError("e2e_valid_loss is the validation loss of angular_triplet_loss (trainer.py:272-275)")
    eng = get_engine()
    loss, _ = eng.metric_loss(features, labels, "e2e_valid", False, scaling=float(getattr(features, "scaling", 0.0) or 0.0),
                              speakers=int(params.num_valid_speakers_per_batch),
                              segments=int(params.num_valid_segments_per_speaker))
    return loss, _metric_endpoints(loss, labels)


def generalized_angular_triplet_loss(features, labels, num_outputs, params, is_training=None, reuse_variables=None,
                                     name="softmax"):
    """Angular triplet loss against the centres of the classes (model/loss.py:708-901, ``loss_compute = "raw"``).

    Args:
        features: [batch, dim] handle WITHOUT L2 normalisation.
        labels: int tensor [batch].
        num_outputs: #classes.
        params: triplet_center "learnable" | "average" (+ triplet_center_momentum), loss_compute "raw", margin,
                target_margin, triplet_topn (0: every class violating the margin, 1: hardest, k: top-k hardest),
                triplet_loss_weight, center_loss_weight, between_loss_weight, l2_loss_weight (must be 0).
        is_training: update the averaged centres / record the backward closure.
    :return: (loss, endpoints)
    """
    assert features.data.dim() == labels.dim() + 1
    assert params.triplet_center == "learnable" or params.triplet_center == "average"
    assert params.loss_compute == "raw" or params.loss_compute == "softplus"
    if params.loss_compute != "raw":
        raise NotImplementedError("Not implemented.")                # loss.py:826
    params.margin = float(params.margin)
    params.target_margin = float(params.target_margin)
    params.triplet_topn = int(params.triplet_topn)
    params.triplet_loss_weight = float(params.triplet_loss_weight)
    params.center_loss_weight = float(params.center_loss_weight)
    params.between_loss_weight = float(params.between_loss_weight)
    assert params.l2_loss_weight == 0.0, "The weight decay is applied by regularization term, not the loss!"
    eng = get_engine()
    if eng.head_shard is not None:
        raise NotImplementedError("generalized_angular_triplet_loss is not available with the class-sharded head")
    average = params.triplet_center == "average"
    loss, cosm, _ = eng.centre_triplet_head(
        features, labels, name + "/output/kernel", num_outputs, bool(is_training),
        scaling=float(getattr(features, "scaling", 0.0) or 0.0), average=average,
        momentum=float(params.triplet_center_momentum) if average else 0.0, margin=params.margin,
        target_margin=params.target_margin, topn=params.triplet_topn, w_triplet=params.triplet_loss_weight,
        w_center=params.center_loss_weight, w_between=params.between_loss_weight)
    endpoints = OrderedDict()
    endpoints["average_centers"] = eng.store.view(name + "/output/kernel")
    endpoints["cos"] = cosm[:, :num_outputs]
    endpoints["labels"] = labels
    return loss, endpoints
