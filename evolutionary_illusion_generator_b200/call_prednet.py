"""Stage-level drop-in for the inference entry points of the reference's `chainer_prednet/PredNet/call_prednet.py`,
running on the GPU through libeig.so (stateful stepping: `eig_prednet_reset` / `eig_prednet_forward`).

  read_image(full_path, size, offset, c)      follows call_prednet.py:29-49   (u8 / 255 in float64, channel-first)
  write_image(image, path)                    follows call_prednet.py:51-61   (float32 * 255, truncated to u8)
  test_image_list(...)                        follows call_prednet.py:129-205 (frame loop, extension block, file names)
  test_prednet(...)                           follows call_prednet.py:209-240 (argument list identical)
Files written: `<output_dir>/<step // skip : 010d>.png` for every `skip_save_frames`-th prediction,
`<... : 010d>_extended.png` for the self-fed steps after every `extension_start` frames, and `test_log.txt` with
"step, mse(prediction, next frame)".  The recurrent state runs on from frame to frame and is cleared after each extension
block, as in the reference.  `gpu` is accepted for compatibility (the engine lives on the current CUDA device);
`initmodel` is the Chainer npz that `serializers.load_npz` would read.  Training is out of scope.
`get_fitnesses_neat` does not come through here - it evaluates the whole population in one `eig_eval`.
"""
import numpy as np
import torch
from PIL import Image

from . import runtime


def read_image(full_path, size, offset, c=3):
    pixels = np.asarray(Image.open(full_path))
    if c < 3:
        planes = pixels.reshape(1, size[1], size[0])
    else:
        planes = np.moveaxis(pixels, 2, 0)
    return planes / 255


def _as_pil(hwc):
    return Image.fromarray(hwc) if hwc.shape[2] > 1 else Image.fromarray(hwc[:, :, 0], "L")


def write_image(image, path):
    image *= 255                                   # in place, float32, like the reference
    _as_pil(np.moveaxis(image, 0, 2).astype(np.uint8)).save(path)


def _device_frame(eng, chw):
    """(C,h,w) array -> float32 (1,h,w,C) tensor on the engine's device (float64 / 255 rounded once to float32)."""
    nhwc = np.ascontiguousarray(np.moveaxis(np.asarray(chw), 0, 2)[None].astype(np.float32))
    return torch.from_numpy(nhwc).to(eng.tdev)


def test_image_list(prednet, imagelist, model, output_dir, channels, size, offset, gpu, logf, skip_save_frames=0,
                    extension_start=0, extension_duration=100, reset_each=False, step=0, verbose=1, reset_at=-1,
                    input_len=-1, c=3):
    """`prednet`: the Engine (the reference hands its PredNet chain in here); `model` is not used."""
    eng = prednet

    def say(*words):
        if verbose == 1:
            print(*words)

    def save(frame_u8, index, suffix):
        name = "%s/%s%s.png" % (output_dir, str(index).zfill(10), suffix)
        say("writing ", name)
        _as_pil(frame_u8[0].cpu().numpy()).save(name)

    eng.prednet_reset(1)
    n_frames = len(imagelist)
    shown = n_frames if input_len <= 0 else min(n_frames, input_len + 1)     # the reference stops once i > input_len
    for i in range(shown):
        prediction, frame_u8 = eng.prednet_forward(_device_frame(eng, read_image(imagelist[i], size, offset, c)))
        if i + 1 < n_frames:                       # the loss is taken against the next frame of the list
            target = _device_frame(eng, read_image(imagelist[i + 1], size, offset, c))
            mse = float(torch.mean((prediction - target) ** 2))
            say("step ", step, " frame ", i, "loss:", mse)
            logf.write("%d, %s\n" % (step, mse))
            logf.flush()
        else:
            say("step ", step, " frame ", i, "loss: last frame.")
        if (step + 1) % skip_save_frames == 0:
            save(frame_u8, step // skip_save_frames, "")
        step += 1
        if extension_start == 0 or step == 0 or step % extension_start != 0:
            continue
        fed_back = prediction                      # unquantised float32, not the saved PNG
        for j in range(extension_duration):
            fed_back, frame_u8 = eng.prednet_forward(fed_back)
            save(frame_u8, step // skip_save_frames + j, "_extended")
        eng.prednet_reset(1)
    return step


def test_prednet(initmodel, sequence_list, size, channels, gpu, output_dir="result", skip_save_frames=0,
                 extension_start=0, extension_duration=0, offset=[0, 0], reset_each=False, verbose=1, reset_at=-1,
                 input_len=-1, c_dim=3):
    eng = runtime.get_engine(size[0], size[1], channels, initmodel, 1)
    if verbose == 1:
        print("sequence_list ", sequence_list)
    step = 0
    with open("test_log.txt", "w") as logf:
        for image_list in sequence_list:
            step = test_image_list(eng, image_list, None, output_dir, channels, size, offset, gpu, logf,
                                   skip_save_frames, extension_start, extension_duration, reset_each, step, verbose,
                                   reset_at, input_len, c_dim)
