"""chainer.serializers.{save,load}_npz: one array per parameter, keyed by its path without the leading slash."""
import numpy


def save_npz(file, obj, compression=True):
    arrays = {path.lstrip("/"): p.data for path, p in obj.namedparams()}
    (numpy.savez_compressed if compression else numpy.savez)(file, **arrays)


def load_npz(file, obj, path="", strict=True):
    with numpy.load(file) as f:
        for name, p in obj.namedparams():
            key = path + name.lstrip("/")
            if key not in f.files:
                if strict:
                    raise KeyError("%s is not found in the npz file" % key)
                continue
            value = f[key]
            if value.shape != p.data.shape:
                raise ValueError("shape mismatch for %s: file %s, link %s" % (key, value.shape, p.data.shape))
            p.data[...] = value
