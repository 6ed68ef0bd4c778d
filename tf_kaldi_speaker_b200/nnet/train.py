"""Training driver with the command line and control flow of egs/voxceleb/v1/nnet/lib/train.py:13-136:

    python -m tf_kaldi_speaker_b200.nnet.train [-c] --config nnet_conf/xxx.json \
        train_dir train_spklist valid_dir valid_spklist model_dir

Epoch loop, validation, learning-rate halving on a stalled validation loss, early stop and the ``learning_rate`` /
``valid_loss`` / ``feature_dim`` bookkeeping files are the reference's; every ``sess.run`` underneath is the CUDA path
(Trainer.train -> train_step).  Under ``torchrun`` (WORLD_SIZE > 1) each rank reads its own batches and the step is the
data-parallel one (parallel.DataParallel)."""
import argparse
import os
import random
import sys

import numpy as np


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("-c", "--cont", action="store_true", help="Continue training from an existing model.")
    parser.add_argument("--config", type=str, help="The configuration file.")
    parser.add_argument("train_dir", type=str, help="The data directory of the training set.")
    parser.add_argument("train_spklist", type=str, help="The spklist file maps the TRAINING speakers to the indices.")
    parser.add_argument("valid_dir", type=str, help="The data directory of the validation set.")
    parser.add_argument("valid_spklist", type=str, help="The spklist maps the VALID speakers to the indices.")
    parser.add_argument("model", type=str, help="The output model directory.")
    args = parser.parse_args(argv)

    import torch
    from .. import parallel
    from ..dataset.data_loader import FeatureReader, KaldiDataRandomQueue
    from ..misc.utils import (ValidLoss, compute_cos_pairwise_eer, load_lr, load_valid_loss, save_codes_and_config)
    from ..model.trainer import Trainer

    rank, world = parallel.init_from_env()
    if torch.cuda.is_available():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    params = save_codes_and_config(args.cont, args.model, args.config) if rank == 0 else None
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        if rank != 0:
            params = save_codes_and_config(True, args.model, args.config)
    model_dir = os.path.join(args.model, "nnet")
    random.seed(params.seed)
    np.random.seed(params.seed)
    if world > 1 and "data_seed" in params.dict:
        params.dict["data_seed"] = int(params.dict["data_seed"]) + 7919 * rank        # distinct batches per replica

    if args.cont:
        import re
        ck = os.path.join(model_dir, "checkpoint")
        if not os.path.isfile(ck):
            sys.exit("Cannot load checkpoint from %s" % model_dir)
        name = re.search(r'"(.*)"', open(ck).readline()).group(1)
        step = int(next(re.finditer(r"(\d+)(?!.*\d)", os.path.basename(name))).group(0))
        start_epoch = int(step / params.num_steps_per_epoch)
    else:
        start_epoch = 0

    learning_rate = params.learning_rate
    learning_rate_array = []
    if os.path.isfile(str(learning_rate)):
        with open(str(learning_rate), "r") as f:
            learning_rate_array = [float(line.strip()) for line in f if line.strip()]
        assert len(learning_rate_array) > params.num_epochs, "The learning rate file is shorter than the num of epochs."
    elif os.path.isfile(os.path.join(model_dir, "learning_rate")):
        learning_rate_array = load_lr(os.path.join(model_dir, "learning_rate"))
        assert len(learning_rate_array) == start_epoch + 1, "Not enough learning rates in the learning_rate file."
    else:
        learning_rate_array = [float(learning_rate)] * (start_epoch + 1)

    dim = FeatureReader(args.train_dir).get_dim()
    if rank == 0:
        with open(os.path.join(model_dir, "feature_dim"), "w") as f:
            f.write("%d\n" % dim)
    num_total_train_speakers = KaldiDataRandomQueue(args.train_dir, args.train_spklist).num_total_speakers
    print("There are %d speakers in the training set and the dim is %d" % (num_total_train_speakers, dim), flush=True)

    min_valid_loss = ValidLoss()
    if os.path.isfile(os.path.join(model_dir, "valid_loss")):
        min_valid_loss = load_valid_loss(os.path.join(model_dir, "valid_loss"))

    trainer = Trainer(params, args.model)
    trainer.build("train", dim=dim, loss_type=params.loss_func, num_speakers=num_total_train_speakers)
    trainer.build("valid", dim=dim, loss_type=params.loss_func, num_speakers=num_total_train_speakers)
    if world > 1:
        parallel.DataParallel(trainer, params.num_speakers_per_batch * params.num_segments_per_speaker)

    if "early_stop_epochs" not in params.dict:
        params.dict["early_stop_epochs"] = 10
    if "min_learning_rate" not in params.dict:
        params.dict["min_learning_rate"] = 1e-5

    for epoch in range(start_epoch, params.num_epochs):
        trainer.train(args.train_dir, args.train_spklist, learning_rate_array[epoch])
        valid_loss, valid_embeddings, valid_labels = trainer.valid(args.valid_dir, args.valid_spklist,
                                                                   batch_type=params.batch_type, output_embeddings=True)
        eer = compute_cos_pairwise_eer(valid_embeddings, valid_labels)
        print("[INFO] Valid EER: %f" % eer, flush=True)

        if not os.path.isfile(str(learning_rate)):
            new_learning_rate = learning_rate_array[epoch]
            if valid_loss < min_valid_loss.min_loss:
                min_valid_loss.min_loss = valid_loss
                min_valid_loss.min_loss_epoch = epoch
            elif epoch - min_valid_loss.min_loss_epoch >= params.reduce_lr_epochs:
                new_learning_rate /= 2
                print("After epoch %d, no improvement. Reduce the learning rate to %.8f"
                      % (min_valid_loss.min_loss_epoch, new_learning_rate), flush=True)
                min_valid_loss.min_loss_epoch += 2
            learning_rate_array.append(new_learning_rate)

        if rank == 0:
            if epoch == 0:
                with open(os.path.join(model_dir, "learning_rate"), "a") as f:
                    f.write("0 %.8f\n" % learning_rate_array[0])
            with open(os.path.join(model_dir, "learning_rate"), "a") as f:
                f.write("%d %.8f\n" % (epoch + 1, learning_rate_array[epoch + 1]))
            with open(os.path.join(model_dir, "valid_loss"), "a") as f:
                f.write("%d %f %f\n" % (epoch, valid_loss, eer))

        if not os.path.isfile(str(learning_rate)):
            if learning_rate_array[epoch + 1] < (params.min_learning_rate - 1e-12) or \
                    epoch - min_valid_loss.min_loss_epoch >= params.early_stop_epochs:
                break
    trainer.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
