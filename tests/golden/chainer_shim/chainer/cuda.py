"""chainer.cuda: the reference hard-codes gpu = 0 (generate_illusion.py:485); here "the GPU" is numpy on the host."""
import numpy

cupy = numpy
available = True


def check_cuda_available():
    pass


class _Device(object):
    def use(self):
        pass


def get_device(*args):
    return _Device()


get_device_from_id = get_device


def to_cpu(a):
    return a


def to_gpu(a, device=None):
    return a
