"""chainer.links used by the reference: Convolution2D and Classifier."""
import numpy

from . import functions as F
from .link import Chain, Link


class Convolution2D(Link):
    def __init__(self, in_channels, out_channels, ksize=None, stride=1, pad=0, nobias=False, initialW=None,
                 initial_bias=None):
        super(Convolution2D, self).__init__()
        self.stride, self.pad = stride, pad
        self.add_param("W", (out_channels, in_channels, ksize, ksize))
        fan_in = in_channels * ksize * ksize
        self.W.data[...] = numpy.random.normal(0, numpy.sqrt(1.0 / fan_in), self.W.data.shape)   # LeCunNormal default
        if nobias:
            self.b = None
        else:
            self.add_param("b", (out_channels,))
            self.b.data[...] = 0

    def __call__(self, x):
        return F.convolution_2d(x, self.W, self.b, self.stride, self.pad)


class Classifier(Chain):
    compute_accuracy = True

    def __init__(self, predictor, lossfun=None, accfun=None):
        super(Classifier, self).__init__()
        self.lossfun, self.accfun = lossfun, accfun
        self.y = self.loss = self.accuracy = None
        self.add_link("predictor", predictor)

    def __call__(self, *args):
        self.y = self.loss = self.accuracy = None
        self.y = self.predictor(*args[:-1])
        self.loss = self.lossfun(self.y, args[-1])
        if self.compute_accuracy:
            self.accuracy = self.accfun(self.y, args[-1])
        return self.loss
