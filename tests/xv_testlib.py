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


def _cm_tools():
    """The small format-1 compressor of tests/golden/make_golden_cm.py (test input generator, not product code)."""
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_golden_cm.py")
    spec = importlib.util.spec_from_file_location("make_golden_cm", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.compress, mod.entry


def make_kaldi_dir(root, num_speakers=6, utts_per_speaker=3, dim=30, min_frames=120, max_frames=260, seed=0):
    """A synthetic Kaldi data directory as the reference's loader expects it (dataset/data_loader.py:19-56,
    dataset/kaldi_io.py:27-62): feats.ark of compressed ('CM ') matrices, feats.scp with byte offsets, spk2utt, utt2spk,
    utt2num_frames, and a spklist file next to it.  Returns (data_dir, spklist_path, {utt: float32 matrix before compression})."""
    import os
    compress, entry = _cm_tools()
    rng = np.random.RandomState(seed)
    data = os.path.join(str(root), "data")
    os.makedirs(data, exist_ok=True)
    ark_path = os.path.join(data, "feats.ark")
    mats, scp, spk2utt, utt2spk, u2n = {}, [], [], [], []
    with open(ark_path, "wb") as ark:
        for s in range(num_speakers):
            spk = "spk%03d" % s
            utts = []
            offset_mean = rng.randn(1, dim) * 2.0
            for u in range(utts_per_speaker):
                utt = "%s-utt%02d" % (spk, u)
                n = int(rng.randint(min_frames, max_frames + 1))
                m = (offset_mean + rng.randn(n, dim) * (0.5 + rng.rand(1, dim))).astype(np.float32)
                mats[utt] = m
                blob = entry(utt, *compress(m))
                pos = ark.tell()
                ark.write(blob)
                scp.append("%s %s:%d" % (utt, ark_path, pos + len(utt) + 1))      # offset of the "\0B" marker
                utts.append(utt)
                utt2spk.append("%s %s" % (utt, spk))
                u2n.append("%s %d" % (utt, n))
            spk2utt.append("%s %s" % (spk, " ".join(utts)))
    for name, lines in (("feats.scp", scp), ("spk2utt", spk2utt), ("utt2spk", utt2spk), ("utt2num_frames", u2n)):
        with open(os.path.join(data, name), "w") as f:
            f.write("\n".join(lines) + "\n")
    spklist = os.path.join(str(root), "spklist")
    with open(spklist, "w") as f:
        f.write("".join("spk%03d %d\n" % (s, s) for s in range(num_speakers)))
    return data, spklist, mats
