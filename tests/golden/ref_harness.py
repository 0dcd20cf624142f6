"""Import the UNMODIFIED reference under stubs.

Only usable where /root/reference exists (the build container).  It is used by
tests/golden/make_golden.py to generate the committed fixtures and by the
`-m "not gpu"` tests marked `needs_reference` to pin the oracle against the
reference's own code.  Nothing on the GPU box imports this module.
"""
import os
import sys
import types

REF = os.environ.get("EIG_REFERENCE", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF, "generate_illusion.py"))


def required_for_output(inputs, outputs, connections):
    # neat-python 0.92 neat/graphs.py (third-party, not vendored): restated in SURVEY.md §8(c).
    required = set(outputs)
    s = set(outputs)
    while 1:
        t = set(a for (a, b) in connections if b in s and a not in s)
        if not t:
            break
        layer_nodes = set(x for x in t if x not in inputs)
        if not layer_nodes:
            break
        required = required.union(layer_nodes)
        s = s.union(t)
    return required


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_loaded = {}


def load():
    """Returns a namespace with the reference modules: gi (generate_illusion), fc, of, cppn."""
    if _loaded:
        return _loaded["ns"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF)
    # Chainer: a functional numpy stand-in for the primitives net.py / call_prednet.py use (tests/golden/chainer_shim),
    # so the reference's PredNet stage runs unmodified on the CPU
    shim = os.path.join(os.path.dirname(os.path.abspath(__file__)), "chainer_shim")
    if shim not in sys.path:
        sys.path.insert(0, shim)
    import chainer  # noqa: F401
    assert chainer.__version__.endswith("shim")
    _mod("google"); _mod("google.colab"); _mod("google.colab.patches", cv2_imshow=lambda *a, **k: None)
    neat = _mod("neat")
    neat.graphs = _mod("neat.graphs", required_for_output=required_for_output)
    neat.reporting = _mod("neat.reporting", BaseReporter=object)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import generate_illusion as gi
    import fitness_calculator as fc
    from optical_flow import optical_flow as of
    from pytorch_neat.pytorch_neat import cppn
    from chainer_prednet.PredNet import call_prednet
    ns = types.SimpleNamespace(gi=gi, fc=fc, of=of, cppn=cppn, call_prednet=call_prednet)
    _loaded["ns"] = ns
    return ns
