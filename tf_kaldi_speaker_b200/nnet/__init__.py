"""Command-line drivers with the reference's argument surface (egs/voxceleb/v1/nnet/lib/train.py, extract.py)."""
