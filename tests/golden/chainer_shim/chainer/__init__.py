"""Functional stand-in for the parts of Chainer (v5-v7 API) that the reference's PredNet stage touches.

TEST INFRASTRUCTURE ONLY.  Chainer is a third-party dependency of the reference (`chainer_prednet/README.md:17-34`) that
cannot be installed offline.  This package restates, with numpy, the published behaviour of the handful of primitives
that `/root/reference/chainer_prednet/PredNet/net.py` and `call_prednet.py` call, so that those two files - and with them
the whole `get_fitnesses_neat` - can be executed UNMODIFIED on the CPU (`tests/golden/ref_harness.py`).  What it pins is
the reference's own wiring: which convolution feeds what, gate order, state handling, the frame protocol, image
quantisation.  The primitives themselves (cross-correlation, 2x2 max pooling, nearest unpooling, clipped ReLU, the
tanh-form sigmoid, npz (de)serialisation by parameter path) are restated from Chainer's documentation, not executed from
Chainer.  Nothing in the product or on the GPU box imports this package.
"""
import contextlib

import numpy

from . import variable  # noqa: F401
from .variable import Parameter, Variable  # noqa: F401
from .link import Chain, Link  # noqa: F401
from . import computational_graph, cuda, functions, links, optimizers, serializers  # noqa: F401

__version__ = "0.0-shim"


@contextlib.contextmanager
def using_config(name, value):
    yield


def as_array(x):
    return x.data if isinstance(x, Variable) else numpy.asarray(x)
