"""Host logic of the training driver (tf_kaldi_speaker_b200/nnet/train.py) against the policy of
egs/voxceleb/v1/nnet/lib/train.py:57-75, 106-139, simulated epoch by epoch with a hand-written loss sequence:
learning-rate halving after ``reduce_lr_epochs`` epochs without improvement (with the two epochs of grace the reference
grants by moving the best epoch forward), early stop on ``min_learning_rate`` / ``early_stop_epochs``, fixed schedules
from a file, continuation from the ``learning_rate`` / ``valid_loss`` / ``checkpoint`` files, and the file formats."""
import os

import pytest

from tf_kaldi_speaker_b200.misc.utils import ParamsPlain, load_lr, load_valid_loss
from tf_kaldi_speaker_b200.nnet import train as D


def _reference_policy(losses, lr0, reduce_lr_epochs, early_stop_epochs, min_lr, num_epochs):
    """Plain restatement of train.py:106-139: returns (rates per epoch incl. the next one, epochs actually run)."""
    rates = [lr0]
    best, best_epoch = 1e16, -1
    ran = 0
    for epoch in range(num_epochs):
        loss = losses[epoch]
        ran += 1
        new = rates[epoch]
        if loss < best:
            best, best_epoch = loss, epoch
        elif epoch - best_epoch >= reduce_lr_epochs:
            new /= 2
            best_epoch += 2
        rates.append(new)
        if rates[epoch + 1] < (min_lr - 1e-12) or epoch - best_epoch >= early_stop_epochs:
            break
    return rates, ran


def _params(**kw):
    d = dict(learning_rate=0.01, num_epochs=30, reduce_lr_epochs=2, num_steps_per_epoch=100)
    d.update(kw)
    return ParamsPlain(**d)


@pytest.mark.parametrize("losses,kw", [
    ([5, 4, 3, 3.5, 3.2, 3.1, 2.9] + [3.0] * 23, {}),
    ([1.0] * 30, dict(early_stop_epochs=4)),                                  # never improves after epoch 0: early stop
    ([1.0 - 0.01 * i for i in range(30)], {}),                                # always improves: constant rate, full run
    ([2.0, 1.0] + [1.5] * 28, dict(min_learning_rate=2e-3, early_stop_epochs=100)),      # stops on the rate floor
])
def test_rate_plan_follows_the_reference_policy(tmp_path, losses, kw):
    p = _params(**kw)
    plan = D._RatePlan(p, str(tmp_path), 0)
    journal = D._Journal(str(tmp_path), True)
    ran = 0
    for epoch in range(p.num_epochs):
        ran += 1
        done = plan.after_epoch(epoch, losses[epoch])
        journal.epoch(epoch, plan, losses[epoch], 0.1)
        if done:
            break
    want, want_ran = _reference_policy(losses, 0.01, 2, p.dict["early_stop_epochs"], p.dict["min_learning_rate"], p.num_epochs)
    assert ran == want_ran
    assert plan.rates == pytest.approx(want)
    # files: "epoch rate" (the initial rate twice: line 0 and the rate of epoch 1), "epoch loss eer"
    assert load_lr(os.path.join(str(tmp_path), "learning_rate")) == pytest.approx(want, abs=6e-9)     # "%.8f"
    lines = open(os.path.join(str(tmp_path), "valid_loss")).read().strip().splitlines()
    assert len(lines) == ran and lines[0].split()[0] == "0" and len(lines[0].split()) == 3
    best = load_valid_loss(os.path.join(str(tmp_path), "valid_loss"))
    assert best.min_loss == pytest.approx(min(losses[:ran]))


def test_continuation_and_fixed_schedules(tmp_path):
    nnet = str(tmp_path)
    with open(os.path.join(nnet, "learning_rate"), "w") as f:
        f.write("0 0.01000000\n1 0.01000000\n2 0.00500000\n")
    with open(os.path.join(nnet, "valid_loss"), "w") as f:
        f.write("0 3.000000 0.200000\n1 3.100000 0.210000\n")
    with open(os.path.join(nnet, "checkpoint"), "w") as f:
        f.write('model_checkpoint_path: "model-200"\nall_model_checkpoint_paths: "model-100"\n')
    assert D._checkpoint_step(nnet) == 200
    p = _params()
    plan = D._RatePlan(p, nnet, first_epoch=2)            # epochs 0 and 1 are done: three rates are on file
    assert plan.rates == pytest.approx([0.01, 0.01, 0.005])
    assert plan.best.min_loss == pytest.approx(3.0) and plan.best.min_loss_epoch == 0
    with pytest.raises(AssertionError):
        D._RatePlan(p, nnet, first_epoch=5)               # "Not enough learning rates in the learning_rate file."
    # a fixed schedule: learning_rate names a file with one rate per line, longer than the run; never adapts, never stops
    sched = os.path.join(nnet, "schedule.txt")
    with open(sched, "w") as f:
        f.write("\n".join("%g" % (0.1 / (i + 1)) for i in range(6)) + "\n")
    pf = _params(learning_rate=sched, num_epochs=5)
    fixed = D._RatePlan(pf, str(tmp_path / "none"), 0)
    assert fixed.fixed and len(fixed.rates) == 6
    assert fixed.after_epoch(0, 100.0) is False and len(fixed.rates) == 6
    with pytest.raises(AssertionError):
        D._RatePlan(_params(learning_rate=sched, num_epochs=6), str(tmp_path / "none"), 0)
    with pytest.raises(SystemExit):
        D._checkpoint_step(str(tmp_path / "none"))
    # rank > 0 writes nothing
    silent = D._Journal(str(tmp_path / "none"), False)
    silent.feature_dim(30)
    silent.epoch(0, plan, 1.0, 0.1)
    assert not os.path.exists(str(tmp_path / "none"))


def test_command_line_matches_the_reference_driver():
    a = D._parse(["-c", "--config", "conf.json", "train", "train/spklist", "valid", "valid/spklist", "exp/model"])
    assert a.cont and a.config == "conf.json" and a.model == "exp/model"
    assert (a.train_dir, a.train_spklist, a.valid_dir, a.valid_spklist) == ("train", "train/spklist", "valid", "valid/spklist")
    assert D._parse(["--config", "c", "a", "b", "c", "d", "e"]).cont is False
