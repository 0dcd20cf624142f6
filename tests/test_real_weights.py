"""Known-answer harness for REAL PredNet weights (VERDICT r1 "missing" 3).

The reference repository records the ratings its own `fitness_calculator.get_vectors` / `calculate_fitness` gave eight
published illusions with the published weight files (illusions_rating/gorilla_data/2025/eigen_own_ratings.csv:2-11;
images illusions_rating/EIGEN-images/*/small.png, committed under tests/golden/eigen_images with the table).  The weight
files themselves are on figshare (illusion_generation.ipynb:140-142: black-and-white model
https://doi.org/10.6084/m9.figshare.13280120, channels 1,16,32,64; colour model figshare 11931222, channels
3,48,96,192) and cannot be fetched here, so the test is keyed on the environment:

  EIG_REAL_MODEL_BW=/path/to/bw.model  EIG_REAL_MODEL_COLOR=/path/to/color.model  pytest -m gpu tests/test_real_weights.py -s

It rates every image through the drop-in `fitness_calculator` (tensor-core path under EIG_CONV=auto, with the range
guard report: a model whose activations leave the split-fp16 range falls back to the exact-fp32 path and says so) and
compares with the recorded rating.  The csv does not say which StructureType each rating used (rows 01/02 give two
ratings for one file), so the recorded value has to be reproduced by ONE of the structures the reference scores
(generate_illusion.py:557-616) within the table's 3-decimal print plus the 1e-3 relative parity tolerance.
"""
import json
import os
import warnings

import numpy as np
import pytest

from conftest import GOLDEN

IMAGES = os.path.join(GOLDEN, "eigen_images")


def _model_for(mode):
    return os.environ.get("EIG_REAL_MODEL_BW" if mode == "L" else "EIG_REAL_MODEL_COLOR")


def test_rating_table_and_images_are_committed():
    """Runs without a GPU: the fixture is complete and is what the reference's table describes."""
    from PIL import Image
    table = json.load(open(os.path.join(IMAGES, "ratings.json")))["ratings"]
    assert len(table) == 7
    for row in table:
        im = Image.open(os.path.join(IMAGES, row["file"] + ".png"))
        assert im.size == (160, 120) and im.mode in ("L", "RGB")
        assert 0.0 <= row["score"] < 1.0


@pytest.mark.gpu
@pytest.mark.skipif(not (os.environ.get("EIG_REAL_MODEL_BW") or os.environ.get("EIG_REAL_MODEL_COLOR")),
                    reason="published PredNet weight files not available (set EIG_REAL_MODEL_BW / EIG_REAL_MODEL_COLOR)")
def test_published_illusions_get_the_recorded_ratings():
    from PIL import Image
    from evolutionary_illusion_generator_b200 import fitness_calculator as FC, runtime
    from evolutionary_illusion_generator_b200.grid import StructureType
    table = json.load(open(os.path.join(IMAGES, "ratings.json")))["ratings"]
    w, h = 160, 120
    rated, report = 0, []
    for fname in sorted({r["file"] for r in table}):
        path = os.path.join(IMAGES, fname + ".png")
        mode = Image.open(path).mode
        model = _model_for(mode)
        if not model:
            continue
        channels = (1, 16, 32, 64) if mode == "L" else (3, 48, 96, 192)
        with warnings.catch_warnings(record=True) as caught:
            warnings.simplefilter("always")
            vec = FC.get_vectors(path, model, channels, w, h)
        eng = runtime.get_engine(w, h, channels, model, 1)
        scores = {st.name: FC.calculate_fitness(st, vec, path, w, h) for st in StructureType}
        fell_back = any("split-fp16 range" in str(c.message) for c in caught)
        report.append((fname, getattr(eng, "conv_mode", "?"), fell_back, 0 if vec[0] is None else len(vec), scores))
        for row in [r for r in table if r["file"] == fname]:
            tol = 5e-4 + 1e-3 * max(row["score"], 1e-3)
            assert any(abs(s - row["score"]) <= tol for s in scores.values()), (row, scores)
            rated += 1
    for line in report:
        print("real weights: %s conv=%s range-fallback=%s vectors=%d scores=%s" % line)
    assert rated > 0
    assert np.isfinite([v for r in report for v in r[4].values()]).all()
