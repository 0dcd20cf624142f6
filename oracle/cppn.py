"""Oracle: genome -> CPPN graph -> image.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates, with torch-CPU tensors (the reference's own numeric back end for this stage):
  * neat-python 0.92 `neat.graphs.required_for_output` (third-party)
  * `create_cppn`                    /root/reference/pytorch_neat/pytorch_neat/cppn.py:168-235
  * `Node.activate` / `get_activs`   cppn.py:75-94
  * activations / aggregations       pytorch_neat/pytorch_neat/activations.py:19-55, aggregations.py:19-30
  * `get_image_from_cppn`            /root/reference/generate_illusion.py:372-460
Pinned byte-for-byte against the reference code itself by tests/test_oracle_vs_reference.py.
"""
import numpy as np
import torch


def required_nodes(input_keys, output_keys, connection_keys):
    """neat.graphs.required_for_output: walks ALL connection keys (disabled ones too, cppn.py:171-173)."""
    needed = set(output_keys)
    seen = set(output_keys)
    inputs = set(input_keys)
    while True:
        frontier = {a for (a, b) in connection_keys if b in seen and a not in seen}
        if not frontier:
            break
        hidden = frontier - inputs
        if not hidden:
            break
        needed |= hidden
        seen |= frontier
    return needed


def incoming_table(genome, input_keys, output_keys):
    """node key -> ordered [(src key, weight)] exactly as cppn.py:176-194 builds `node_inputs`."""
    needed = required_nodes(input_keys, output_keys, genome.connections)
    table = {k: [] for k in output_keys}
    outs = set(output_keys)
    for cg in genome.connections.values():
        if not cg.enabled:
            continue
        src, dst = cg.key
        if dst not in needed and src not in needed:
            continue
        if src in outs:
            continue
        table.setdefault(dst, []).append((src, cg.weight))
        table.setdefault(src, [])
    return table


_ACT = {
    "sigmoid": lambda t: torch.sigmoid(5 * t),
    "tanh": lambda t: torch.tanh(2.5 * t),
    "abs": torch.abs,
    "gauss": lambda t: torch.exp(-5.0 * t ** 2),
    "identity": lambda t: t,
    "sin": torch.sin,
    "relu": torch.nn.functional.relu,
}


def _aggregate(name, terms):
    if name == "sum":
        acc = 0
        for t in terms:
            acc = acc + t
        return acc
    if name == "prod":
        acc = 1
        for t in terms:
            acc = acc * t
        return acc
    raise KeyError(name)


def eval_output(genome, input_keys, output_keys, out_key, leaf_values):
    """Value of one output node over all pixels.  A fresh memo per call, like Node.__call__ (cppn.py:96-108)."""
    table = incoming_table(genome, input_keys, output_keys)
    shape = next(iter(leaf_values.values())).shape
    memo = dict(leaf_values)

    def value(key):
        if key in memo:
            return memo[key]
        gene = genome.nodes[key]
        srcs = table[key]
        if not srcs:
            # cppn.py:79-80: childless node is the constant bias (torch default dtype float32), no activation
            out = torch.full(shape, gene.bias)
        else:
            terms = [w * value(s) for s, w in srcs]
            pre = _aggregate(gene.aggregation, terms)
            out = _ACT[gene.activation](gene.response * pre + gene.bias)
        memo[key] = out
        return out

    return value(out_key)


def numpy_u8_cast(a):
    """What `np.array(float64_array, dtype=np.uint8)` does on x86-64 (generate_illusion.py:403,457)."""
    return np.array(a, dtype=np.uint8)


def render(grid, genome, c_dim, w, h, input_keys, output_keys, bg=1, gradient=1):
    """Restatement of get_image_from_cppn -> uint8 ndarray (h,w,3) or (h,w).

    Deviation from the reference, as SURVEY.md "defects" prescribes: grid planes are reshaped to (h,w)
    (Bands returns (1,w*h,1) and crashes the reference), and in colour only outputs 0..2 are used
    (6-output configs crash the reference at channel 3).
    """
    x_dat = np.asarray(grid["x_mat"], dtype=np.float64).reshape(h, w)
    y_dat = np.asarray(grid["y_mat"], dtype=np.float64).reshape(h, w)
    leaves = {input_keys[0]: torch.tensor(x_dat.flatten()), input_keys[1]: torch.tensor(y_dat.flatten())}
    is_bg = x_dat == -1

    def plane(k):
        v = eval_output(genome, input_keys, output_keys, output_keys[k], leaves)
        return np.reshape(v.numpy(), (h, w))

    if c_dim > 1:
        if gradient == 1:
            arr = np.zeros((h, w, c_dim))
            for c in range(min(c_dim, len(output_keys))):
                arr[:, :, c] = plane(c)
                arr[:, :, c][is_bg] = bg
            return numpy_u8_cast(arr * 255.0)
        idx = numpy_u8_cast(plane(0) * 4.0)
        img = np.zeros((h, w, 3))
        for c in range(3):
            img[:, :, c] = np.where(idx == 0, 255, img[:, :, c])
        img[:, :, 0] = np.where(idx == 1, 255, img[:, :, 0])
        img[:, :, 1] = np.where(idx == 2, 255, img[:, :, 1])
        img[:, :, 2] = np.where(idx == 3, 255, img[:, :, 2])
        img[is_bg] = bg * 255
        return numpy_u8_cast(img)
    arr = plane(0).copy()
    arr[is_bg] = bg
    if gradient == 0:
        arr = np.round(arr)
    return numpy_u8_cast(arr * 255.0)
