"""Drop-in for the single-image rating path of /root/reference/fitness_calculator.py (468-548).

  get_vectors(image_path, model_name, channels, w, h)          fitness_calculator.py:468-502
  calculate_fitness(structure, vectors, image_path, w, h)      fitness_calculator.py:505-548
Both run on the GPU through libeig.so.  `calculate_fitness` returns 0.0 where the reference raises
UnboundLocalError because no branch assigned `score_d` (SURVEY.md "defects").
"""
import numpy as np
import torch

from . import engine as engine_mod, runtime
from .grid import StructureType  # noqa: F401

REPEAT, EXTENSION = 20, 2
_score_engines = {}


def _load_image(image_path, c_dim, w, h):
    from PIL import Image
    im = Image.open(image_path)
    im = im.convert("RGB") if c_dim == 3 else im.convert("L")
    a = np.array(im)
    if a.shape[0] != h or a.shape[1] != w:
        raise ValueError("image is %dx%d, the model expects %dx%d" % (a.shape[1], a.shape[0], w, h))
    return np.ascontiguousarray(a.reshape(h, w, c_dim))


def get_vectors(image_path, model_name, channels, w, h):
    """Flow between the INPUT image and extension frame #2 after 20 static forwards (line 493-498).
    Returns an (n,4) float32 array of rows (x, y, dx, dy), or [None] when nothing was tracked."""
    c_dim = channels[0]
    eng = runtime.get_engine(w, h, channels, model_name, 1)
    img = torch.from_numpy(_load_image(image_path, c_dim, w, h))[None].to(eng.tdev)
    x = (img.double() / 255).float()  # read_image: float32(float64(u8)/255)
    frames = eng.prednet(x, REPEAT, EXTENSION)
    _, _, vectors, nvec = eng.flow(img, frames[2])
    n = int(nvec[0])
    if n == 0:
        return [None]
    return vectors[0, :n].cpu().numpy()


def calculate_fitness(structure, vectors, image_path, w, h, engine=None):
    if vectors is None or len(vectors) == 0 or (len(vectors) == 1 and vectors[0] is None):
        return 0.0
    eng = engine
    if eng is None:
        eng = next((e for e in runtime._engines.values() if (e.w, e.h) == (w, h)), None)
        if eng is None:   # scoring needs no PredNet weights: a cached score-only context of this image size
            eng = _score_engines.get((w, h))
            if eng is None:
                eng = _score_engines[(w, h)] = runtime.engine_factory(w, h, (1, 4, 4, 4), 1)
    if len(vectors) > engine_mod.MAX_CORNERS:
        raise ValueError("calculate_fitness: %d vectors, the flow stage never yields more than %d (maxCorners, "
                         "optical_flow.py:51)" % (len(vectors), engine_mod.MAX_CORNERS))
    v = np.zeros((1, engine_mod.MAX_CORNERS, 4), dtype=np.float32)
    arr = np.asarray(vectors, dtype=np.float32)[:engine_mod.MAX_CORNERS]
    v[0, :len(arr)] = arr
    nv = np.array([len(arr)], dtype=np.int32)
    fit = eng.score(torch.from_numpy(v).to(eng.tdev), torch.from_numpy(nv).to(eng.tdev), int(structure))
    return float(fit[0])
