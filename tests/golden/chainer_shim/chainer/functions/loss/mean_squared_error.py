import numpy

from ...variable import Variable, _raw


def mean_squared_error(x0, x1):
    d = (_raw(x0) - _raw(x1)).ravel()
    return Variable(numpy.array(d.dot(d) / d.size, dtype=d.dtype))
