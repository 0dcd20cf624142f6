"""chainer.Link / chainer.Chain: parameter and child registration, `namedparams` paths as Chainer builds them."""
import numpy

from .variable import Parameter


class Link(object):
    def __init__(self, **params):
        self._params = []
        self.name = None
        for name, shape in params.items():      # legacy `Link(W=shape)` form (net.py:14)
            self.add_param(name, shape)

    xp = numpy

    def _registry(self, attr):
        if attr not in self.__dict__:            # subclasses may set attributes before calling __init__
            self.__dict__[attr] = []
        return self.__dict__[attr]

    def add_param(self, name, shape=None, dtype=numpy.float32, initializer=None):
        data = numpy.full(shape, numpy.nan, dtype=dtype)     # uninitialised, like Chainer
        self._registry("_params").append(name)
        setattr(self, name, Parameter(data, name))

    def params(self, include_uninit=True):
        for name in self._registry("_params"):
            yield getattr(self, name)

    def namedparams(self, include_uninit=True):
        for name in self._registry("_params"):
            yield "/" + name, getattr(self, name)

    def to_cpu(self): return self
    def to_gpu(self, device=None): return self


class Chain(Link):
    def __init__(self, **links):
        super(Chain, self).__init__()
        self._registry("_children")
        for name, link in links.items():         # legacy `Chain(name=link)` form (net.py:46)
            self.add_link(name, link)

    def add_link(self, name, link):
        link.name = name
        self._registry("_children").append(name)
        setattr(self, name, link)

    def params(self, include_uninit=True):
        for p in super(Chain, self).params(include_uninit):
            yield p
        for name in self._registry("_children"):
            for p in getattr(self, name).params(include_uninit):
                yield p

    def namedparams(self, include_uninit=True):
        for ret in super(Chain, self).namedparams(include_uninit):
            yield ret
        for name in self._registry("_children"):
            prefix = "/" + name
            for path, param in getattr(self, name).namedparams(include_uninit):
                yield prefix + path, param
