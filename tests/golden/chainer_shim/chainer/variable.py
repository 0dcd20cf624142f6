"""chainer.Variable / chainer.Parameter: a numpy array with arithmetic; no graph (inference only)."""
import numpy


def _raw(x):
    return x.data if isinstance(x, Variable) else x


class Variable(object):
    __array_priority__ = 200

    def __init__(self, data=None, **kwargs):
        self.data = data

    array = property(lambda self: self.data)
    shape = property(lambda self: self.data.shape)
    dtype = property(lambda self: self.data.dtype)

    # out-of-place like Chainer: `a += b` rebinds a to a new Variable, it never writes into a's array
    def __add__(self, o): return Variable(self.data + _raw(o))
    def __radd__(self, o): return Variable(_raw(o) + self.data)
    def __sub__(self, o): return Variable(self.data - _raw(o))
    def __rsub__(self, o): return Variable(_raw(o) - self.data)
    def __mul__(self, o): return Variable(self.data * _raw(o))
    def __rmul__(self, o): return Variable(_raw(o) * self.data)
    def __truediv__(self, o): return Variable(self.data / _raw(o))
    def __neg__(self): return Variable(-self.data)
    __iadd__, __isub__, __imul__ = __add__, __sub__, __mul__

    def __float__(self): return float(self.data)
    def __len__(self): return len(self.data)
    def __repr__(self): return "variable(%r)" % (self.data,)

    def unchain_backward(self): pass
    def to_cpu(self): return self
    def to_gpu(self, device=None): return self


class Parameter(Variable):
    def __init__(self, data=None, name=None):
        super(Parameter, self).__init__(data)
        self.name = name
