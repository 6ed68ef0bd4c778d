"""Config objects of the reference (misc/utils.py:13-61): JSON -> attribute bag with a ``.dict`` view that the
operator functions both read and mutate (defaults, float() coercions, ``global_step``)."""
import json


class Params():
    """Loads hyper-parameters from a nnet_conf/*.json file (misc/utils.py:13-41); the files load unchanged."""

    def __init__(self, json_path):
        self.update(json_path)

    def save(self, json_path):
        with open(json_path, 'w') as f:
            json.dump(self.__dict__, f, indent=4)

    def update(self, json_path):
        with open(json_path) as f:
            params = json.load(f)
            self.__dict__.update(params)

    @property
    def dict(self):
        return self.__dict__


class ParamsPlain():
    """Manual parameter bag (misc/utils.py:44-61)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    @property
    def dict(self):
        return self.__dict__


def substring_in_list(s, varlist):
    """misc/utils.py:315-330."""
    if varlist is None:
        return False
    for v in varlist:
        if v in s:
            return True
    return False


# ---- host-side helpers the training driver imports (egs/voxceleb/v1/nnet/lib/train.py:8-11) ---------------------------
class ValidLoss():
    """Best validation loss so far and the epoch it was seen in (misc/utils.py:186-190)."""

    def __init__(self):
        self.min_loss = 1e16
        self.min_loss_epoch = -1


def load_lr(filename):
    """'<epoch> <learning rate>' per line -> [learning rate] (misc/utils.py:193-200)."""
    with open(filename, "r") as f:
        return [float(line.strip().split(" ")[1]) for line in f if line.strip()]


def load_valid_loss(filename):
    """'<epoch> <loss> ...' per line -> ValidLoss holding the minimum (misc/utils.py:203-214)."""
    best = ValidLoss()
    with open(filename, "r") as f:
        for line in f:
            if not line.strip():
                continue
            epoch, loss = line.strip().split(" ")[:2]
            if float(loss) < best.min_loss:
                best.min_loss, best.min_loss_epoch = float(loss), int(epoch)
    return best


def save_codes_and_config(cont, model, config):
    """misc/utils.py:64-123 without the source-tree snapshot (the reference copies its own Python packages into
    ``model/codes``; this package is installed, not copied): ``cont`` re-reads ``model/nnet/config.json``, otherwise an
    existing ``nnet`` is moved to ``.backup`` and the config file is copied to ``model/nnet/config.json``."""
    import os
    import shutil
    import sys
    nnet = os.path.join(model, "nnet")
    if cont:
        if not os.path.isdir(nnet):
            sys.exit("To continue training the model, nnet must be existed in %s." % model)
        return Params(os.path.join(nnet, "config.json"))
    if os.path.isdir(nnet):
        backup = os.path.join(model, ".backup")
        if os.path.isdir(backup):
            shutil.rmtree(backup)
        os.makedirs(backup)
        shutil.move(nnet, backup)
    os.makedirs(nnet)
    shutil.copyfile(config, os.path.join(nnet, "config.json"))
    return Params(config)


def compute_cos_pairwise_eer(embeddings, labels, max_num_embeddings=1000):
    """Pairwise cosine-scoring equal error rate of a set of embeddings (misc/utils.py:273-312): all i < j pairs, target
    iff the labels agree; the EER is where the false-accept and false-reject rates cross (linear interpolation between
    the two neighbouring thresholds instead of the reference's interp1d + brentq)."""
    import numpy as np
    e = np.asarray(embeddings, dtype=np.float64)
    lab = np.asarray(labels)
    e = e / np.sqrt(np.sum(e ** 2, axis=1, keepdims=True) + 1e-12)
    n = e.shape[0]
    if n > max_num_embeddings:
        step = n // max_num_embeddings
        e, lab = e[::step], lab[::step]
        n = e.shape[0]
    iu = np.triu_indices(n, 1)
    scores = (e @ e.T)[iu]
    keys = (lab[iu[0]] == lab[iu[1]])
    nt, nn = int(keys.sum()), int((~keys).sum())
    if nt == 0 or nn == 0:
        return 0.0
    order = np.argsort(-scores, kind="stable")
    k = keys[order]
    fa = np.concatenate([[0.0], np.cumsum(~k) / nn])          # accept the top-i scores
    fr = np.concatenate([[1.0], 1.0 - np.cumsum(k) / nt])
    d = fa - fr
    i = int(np.argmax(d >= 0))
    if i == 0:
        return float(fa[0])
    t = d[i - 1] / (d[i - 1] - d[i]) if d[i] != d[i - 1] else 0.0
    return float(fa[i - 1] + t * (fa[i] - fa[i - 1]))
