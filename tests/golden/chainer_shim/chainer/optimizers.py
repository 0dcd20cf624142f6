"""chainer.optimizers: imported by call_prednet.py, only used for training (out of scope)."""
