"""Training driver: the command line of egs/voxceleb/v1/nnet/lib/train.py:13-23 and the files it leaves in
``<model>/nnet`` (``learning_rate``, ``valid_loss``, ``feature_dim``, checkpoints), so that wrap/train_wrapper.sh and the
recipes around it keep working:

    python -m tf_kaldi_speaker_b200.nnet.train [-c] --config nnet_conf/xxx.json \\
        train_dir train_spklist valid_dir valid_spklist model_dir

One epoch = ``Trainer.train`` (CUDA-graphed steps fed by the host loader) followed by ``Trainer.valid`` and the pairwise-cosine
EER of the validation embeddings.  The learning-rate policy is the reference's (train.py:57-75, 106-139): a fixed schedule
when ``learning_rate`` names a file, otherwise halve the rate whenever the validation loss has not improved for
``reduce_lr_epochs`` epochs and stop when it falls below ``min_learning_rate`` or nothing improved for ``early_stop_epochs``.
Under ``torchrun`` every rank runs this loop on its own batches (parallel.DataParallel); rank 0 alone writes the files."""
import argparse
import os
import random
import re
import sys

import numpy as np


def _parse(argv):
    ap = argparse.ArgumentParser(description="x-vector training on the CUDA path")
    ap.add_argument("-c", "--cont", action="store_true", help="resume from the newest checkpoint in <model>/nnet")
    ap.add_argument("--config", type=str, help="network / optimisation settings (nnet_conf JSON)")
    ap.add_argument("train_dir", type=str, help="Kaldi data directory of the training set (feats.scp, utt2spk, ...)")
    ap.add_argument("train_spklist", type=str, help="speaker -> class index table of the training set")
    ap.add_argument("valid_dir", type=str, help="Kaldi data directory of the validation set")
    ap.add_argument("valid_spklist", type=str, help="speaker -> class index table of the validation set")
    ap.add_argument("model", type=str, help="output directory; everything is written to <model>/nnet")
    return ap.parse_args(argv)


def _checkpoint_step(nnet_dir):
    """Step number of the checkpoint the ``checkpoint`` index file points at (the digits closing its name)."""
    index = os.path.join(nnet_dir, "checkpoint")
    if not os.path.isfile(index):
        sys.exit("Cannot load checkpoint from %s" % nnet_dir)
    with open(index) as f:
        newest = re.search(r'"(.*)"', f.readline()).group(1)
    return int(re.findall(r"\d+", os.path.basename(newest))[-1])


class _RatePlan(object):
    """Learning rate per epoch.  ``rates[e]`` is the rate of epoch e; adaptive plans grow by one entry per finished epoch."""

    def __init__(self, params, nnet_dir, first_epoch):
        from ..misc.utils import ValidLoss, load_lr, load_valid_loss
        spec = str(params.learning_rate)
        history = os.path.join(nnet_dir, "learning_rate")
        self.fixed = os.path.isfile(spec)
        if self.fixed:                                   # one rate per line, longer than the run
            with open(spec) as f:
                self.rates = [float(tok) for tok in f.read().split()]
            assert len(self.rates) > params.num_epochs, "The learning rate file is shorter than the num of epochs."
        elif os.path.isfile(history):                    # a continued run picks up the rates it logged
            self.rates = load_lr(history)
            assert len(self.rates) == first_epoch + 1, "Not enough learning rates in the learning_rate file."
        else:
            self.rates = [float(params.learning_rate)] * (first_epoch + 1)
        losses = os.path.join(nnet_dir, "valid_loss")
        self.best = load_valid_loss(losses) if os.path.isfile(losses) else ValidLoss()
        self.patience = params.reduce_lr_epochs
        self.stop_after = params.dict.setdefault("early_stop_epochs", 10)
        self.floor = params.dict.setdefault("min_learning_rate", 1e-5)

    def after_epoch(self, epoch, valid_loss):
        """Record the epoch's validation loss, append the next epoch's rate; True when training should end."""
        if self.fixed:
            return False
        rate = self.rates[epoch]
        if valid_loss < self.best.min_loss:
            self.best.min_loss, self.best.min_loss_epoch = valid_loss, epoch
        elif epoch - self.best.min_loss_epoch >= self.patience:
            rate *= 0.5
            print("After epoch %d, no improvement. Reduce the learning rate to %.8f" % (self.best.min_loss_epoch, rate), flush=True)
            self.best.min_loss_epoch += 2               # two more epochs of grace before the next cut
        self.rates.append(rate)
        return rate < self.floor - 1e-12 or epoch - self.best.min_loss_epoch >= self.stop_after


class _Journal(object):
    """The text files downstream scripts read: ``feature_dim``, ``learning_rate`` ("epoch rate"), ``valid_loss``
    ("epoch loss eer").  Only rank 0 writes."""

    def __init__(self, nnet_dir, active):
        self.dir, self.active = nnet_dir, active

    def _append(self, name, line):
        if self.active:
            with open(os.path.join(self.dir, name), "a") as f:
                f.write(line)

    def feature_dim(self, dim):
        if self.active:
            with open(os.path.join(self.dir, "feature_dim"), "w") as f:
                f.write("%d\n" % dim)

    def epoch(self, epoch, plan, valid_loss, eer):
        if epoch == 0:
            self._append("learning_rate", "0 %.8f\n" % plan.rates[0])
        self._append("learning_rate", "%d %.8f\n" % (epoch + 1, plan.rates[epoch + 1]))
        self._append("valid_loss", "%d %f %f\n" % (epoch, valid_loss, eer))


def main(argv=None):
    args = _parse(argv)
    import torch
    from .. import parallel
    from ..dataset.data_loader import FeatureReader, KaldiDataRandomQueue
    from ..misc.utils import compute_cos_pairwise_eer, save_codes_and_config
    from ..model.trainer import Trainer

    rank, world = parallel.init_from_env()
    if torch.cuda.is_available():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    # rank 0 lays out <model>/nnet (config copy, fresh or continued); the others read what it wrote
    params = save_codes_and_config(args.cont, args.model, args.config) if rank == 0 else None
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        if rank != 0:
            params = save_codes_and_config(True, args.model, args.config)
    nnet_dir = os.path.join(args.model, "nnet")
    random.seed(params.seed)
    np.random.seed(params.seed)
    if world > 1 and "data_seed" in params.dict:
        params.dict["data_seed"] = int(params.dict["data_seed"]) + 7919 * rank        # distinct batches per replica

    first_epoch = _checkpoint_step(nnet_dir) // int(params.num_steps_per_epoch) if args.cont else 0
    plan = _RatePlan(params, nnet_dir, first_epoch)
    journal = _Journal(nnet_dir, rank == 0)

    dim = FeatureReader(args.train_dir).get_dim()
    journal.feature_dim(dim)
    speakers = KaldiDataRandomQueue(args.train_dir, args.train_spklist).num_total_speakers
    print("There are %d speakers in the training set and the dim is %d" % (speakers, dim), flush=True)

    trainer = Trainer(params, args.model)
    for mode in ("train", "valid"):
        trainer.build(mode, dim=dim, loss_type=params.loss_func, num_speakers=speakers)
    if world > 1:
        parallel.DataParallel(trainer, params.num_speakers_per_batch * params.num_segments_per_speaker)

    for epoch in range(first_epoch, params.num_epochs):
        trainer.train(args.train_dir, args.train_spklist, plan.rates[epoch])
        valid_loss, embeddings, labels = trainer.valid(args.valid_dir, args.valid_spklist, batch_type=params.batch_type,
                                                       output_embeddings=True)
        eer = compute_cos_pairwise_eer(embeddings, labels)
        print("[INFO] Valid EER: %f" % eer, flush=True)
        done = plan.after_epoch(epoch, valid_loss)
        journal.epoch(epoch, plan, valid_loss, eer)
        if done:
            break
    trainer.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
