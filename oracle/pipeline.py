"""Oracle: the whole fitness path for a population.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Order of work and frame pairing follow `get_fitnesses_neat` (/root/reference/generate_illusion.py:478-673):
render every genome, 20 PredNet forwards on the static image + 2 self-fed forwards, flow between prediction
#20 and extension #1 (lines 543-546), structure-specific score (557-616).  pair_mode 1 is the single-image
pairing of `fitness_calculator.get_vectors` (fitness_calculator.py:468-502): input image vs extension #2.
"""
import numpy as np

from . import cppn, flow, grid as ogrid, prednet, scoring

PAIR_POPULATION, PAIR_SINGLE_IMAGE = 0, 1


def evaluate_population(genomes, input_keys, output_keys, structure, weights, w, h, channels, c_dim=3,
                        gradient=1, pair_mode=PAIR_POPULATION, flow_impl="np", grid=None, keep=False,
                        conv_hook=None):
    grid = grid if grid is not None else ogrid.create_grid(structure, w, h, 10)
    net = prednet.PredNetOracle(weights, channels, w, h)
    if conv_hook is not None:
        net._conv = conv_hook(net)
    lk = flow.lucas_kanade_np if flow_impl == "np" else flow.lucas_kanade_cv2
    fitness, extra = [], []
    for g in genomes:
        img = cppn.render(grid, g, c_dim, w, h, input_keys, output_keys, bg=1, gradient=gradient)
        frames = prednet.run_genome_frames(net, img, repeat=20, extension=2)
        a, b = (frames[0], frames[1]) if pair_mode == PAIR_POPULATION else (img, frames[2])
        vec = lk(a, b)
        fitness.append(scoring.fitness_from_vectors(structure, vec, w, h))
        if keep:
            extra.append(dict(image=img, frames=frames, vectors=vec))
    return (np.array(fitness), extra) if keep else np.array(fitness)
