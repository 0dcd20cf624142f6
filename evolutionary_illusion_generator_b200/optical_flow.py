"""Drop-in for `lucas_kanade` of /root/reference/optical_flow/optical_flow.py:40-89, on the GPU.

Corner detection and pyramidal LK follow OpenCV's arithmetic (csrc/flow.cuh); window size 50, quality 0.3,
maxCorners 100, minDistance 7, blockSize 7, maxLevel 2, criteria (EPS|COUNT, 10, 0.03) are the reference's
constants (optical_flow.py:51-60, config.yaml).
"""
import numpy as np
import torch

from . import runtime

_flow_engines = {}


def _engine_for(w, h, c_dim):
    key = (w, h, c_dim)
    if key not in _flow_engines:
        # flow-only context: the PredNet buffers are never used, so the channels are minimal and one image pair fits
        _flow_engines[key] = runtime.engine_factory(w, h, (c_dim, 4, 4, 4), 1)
    return _flow_engines[key]


_COLORS = {"blue": (0, 0, 255), "green": (0, 255, 0), "red": (255, 0, 0), "yellow": (255, 255, 0), "white": (255, 255, 255)}


def draw_tracks(img, data, vector_scale=60, circle_size=2, circle_color="yellow", line_width=2, line_color="red"):
    """`draw_tracks` of the reference (optical_flow.py:10-18) on a PIL RGB image: one line per vector from (x, y) to
    (x + scale*dx, y + scale*dy) and a filled circle at its origin (truncated integer coordinates, like the cv2 calls)."""
    from PIL import ImageDraw
    d = ImageDraw.Draw(img)
    for x, y, dx, dy in data:
        x0, y0 = int(x), int(y)
        d.line([(x0, y0), (int(x + vector_scale * dx), int(y + vector_scale * dy))], fill=_COLORS[line_color], width=line_width)
        d.ellipse([x0 - circle_size, y0 - circle_size, x0 + circle_size, y0 + circle_size], fill=_COLORS[circle_color])
    return img


def lucas_kanade_arrays(img1, img2, engine=None):
    """img1, img2: (h,w) or (h,w,3 RGB) uint8 arrays -> list of [x, y, dx, dy] rows (float32)."""
    a, b = np.asarray(img1), np.asarray(img2)
    h, w = a.shape[:2]
    c = 1 if a.ndim == 2 else a.shape[2]
    eng = engine or _engine_for(w, h, c)
    t1 = torch.from_numpy(np.ascontiguousarray(a.reshape(1, h, w, c))).to(eng.tdev)
    t2 = torch.from_numpy(np.ascontiguousarray(b.reshape(1, h, w, c))).to(eng.tdev)
    _, _, vectors, nvec = eng.flow(t1, t2)
    return [list(r) for r in vectors[0, :int(nvec[0])].cpu().numpy()]


def lucas_kanade(file1, file2, output_path="./", vector_scale=60, circle_size=2, circle_color="yellow",
                 line_width=2, line_color="red", save=True, verbose=1, save_name="", engine=None):
    """Same call surface as the reference; `save=True` writes the overlay image (drawn with PIL) and the csv of vectors."""
    import os
    from PIL import Image
    im1, im2 = Image.open(file1), Image.open(file2)
    mode = "RGB" if (im1.mode != "L" or im2.mode != "L") else "L"
    a, b = np.array(im1.convert(mode)), np.array(im2.convert(mode))
    data = lucas_kanade_arrays(a, b, engine=engine)
    image = None
    os.makedirs(os.path.join(output_path, "csv"), exist_ok=True)   # optical_flow.py:45-46 (always)
    if save:
        image = draw_tracks(im2.convert("RGB"), data, vector_scale, circle_size, circle_color, line_width, line_color)
        stem0 = os.path.splitext(os.path.basename(file1))[0]
        image.save(save_name if save_name else os.path.join(output_path, stem0 + ".png"))
        stem = os.path.splitext(os.path.basename(file1))[0]
        with open(os.path.join(output_path, "csv", stem + ".csv"), "w") as f:
            for row in data:
                f.write(",".join(str(v) for v in row) + "\n")
    return {"image": image, "vectors": data}
