"""Host side of the CPPN stage: duck-typed genomes, the genome flattener and synthetic populations.

The render kernel (csrc/render.cuh) interprets a flat, topologically ordered program per genome.
`flatten_genome` builds that program with exactly the graph the reference builds in
`create_cppn` (/root/reference/pytorch_neat/pytorch_neat/cppn.py:168-235):

  * `required_for_output` over ALL connection keys, disabled ones included (cppn.py:171-173;
    neat-python 0.92 neat/graphs.py, third-party);
  * disabled connections and connections leaving an output node are dropped (cppn.py:178-186);
  * children keep `genome.connections` insertion order, because float addition is not associative;
  * a node without children is the constant `bias` and its activation is NOT applied (cppn.py:79-80).

dtype rule reproduced here: `torch.full(shape, bias)` is float32, and a python float
times a float32 tensor stays float32, so any sub-graph made only of constants is evaluated by the
reference in float32 (with torch's float32 kernels) and promoted to float64 where it first meets a
pixel-dependent value.  Constants are therefore folded HERE, with torch float32 ops on small tensors,
and shipped to the kernel as float64 literals; everything pixel-dependent is evaluated on the GPU in fp64.
"""
import configparser
import random
import struct
from types import SimpleNamespace

import numpy as np
import torch

ACT_IDS = {"sigmoid": 0, "tanh": 1, "abs": 2, "gauss": 3, "identity": 4, "sin": 5, "relu": 6}
AGG_IDS = {"sum": 0, "prod": 1}
SLOT_X, SLOT_Y, SLOT_ONE, SLOT_NODE0 = 0, 1, 2, 3
BLOB_MAGIC = 0x45494742  # 'EIGB'
OUT_F32_CONST = 1 << 30
_FOLD_N = 64  # fold constants on 64-element tensors so torch takes its vectorised (SLEEF) path like the reference


# ----------------------------------------------------------------------------- genome stand-ins
class NodeGene:
    __slots__ = ("key", "bias", "response", "activation", "aggregation")

    def __init__(self, key, bias=0.0, response=1.0, activation="sin", aggregation="sum"):
        self.key, self.bias, self.response = key, bias, response
        self.activation, self.aggregation = activation, aggregation


class ConnectionGene:
    __slots__ = ("key", "weight", "enabled")

    def __init__(self, key, weight, enabled=True):
        self.key, self.weight, self.enabled = key, weight, enabled


class Genome:
    """The attributes of neat.DefaultGenome the hot path reads plus `.fitness`."""

    def __init__(self, key=0):
        self.key = key
        self.connections = {}
        self.nodes = {}
        self.fitness = None


def make_config(num_inputs=2, num_outputs=3):
    gc = SimpleNamespace(input_keys=[-i - 1 for i in range(num_inputs)],
                         output_keys=list(range(num_outputs)))
    return SimpleNamespace(genome_config=gc)


# NEAT hyper-parameters of the five BASELINE configs that the synthetic generator needs
# (/root/reference/neat_configs/{default,circles_bw,circles,bands,free}.txt, [DefaultGenome] section).
_ACTS = "sin sigmoid gauss tanh relu abs".split()
NEAT_PRESETS = {
    "default":    dict(num_inputs=4, num_hidden=8,  num_outputs=6, weight_init_mean=0.0),
    "circles_bw": dict(num_inputs=2, num_hidden=20, num_outputs=1, weight_init_mean=0.1),
    "circles":    dict(num_inputs=2, num_hidden=20, num_outputs=3, weight_init_mean=0.1),
    "bands":      dict(num_inputs=2, num_hidden=8,  num_outputs=6, weight_init_mean=0.0),
    "free":       dict(num_inputs=2, num_hidden=20, num_outputs=6, weight_init_mean=0.1),
}
for _p in NEAT_PRESETS.values():
    _p.update(activation_default="sin", activation_options=_ACTS, activation_mutate_rate=0.5,
              bias_init_mean=0.0, bias_init_stdev=1.0, bias_min_value=-30.0, bias_max_value=30.0,
              weight_init_stdev=1.0, weight_min_value=-30.0, weight_max_value=30.0,
              connection_fraction=0.8)
PRESET_IDS = {"default": 0, "circles_bw": 1, "circles": 2, "bands": 3, "free": 4}


def load_neat_preset(path):
    """Read the [DefaultGenome] section of a neat-python config file into a preset dict."""
    cp = configparser.ConfigParser()
    cp.read(path)
    g = cp["DefaultGenome"]
    frac = 1.0
    ic = g.get("initial_connection", "partial_nodirect 0.8").split()
    if len(ic) > 1:
        frac = float(ic[1])
    return dict(num_inputs=g.getint("num_inputs"), num_hidden=g.getint("num_hidden"),
                num_outputs=g.getint("num_outputs"), weight_init_mean=g.getfloat("weight_init_mean"),
                weight_init_stdev=g.getfloat("weight_init_stdev"),
                weight_min_value=g.getfloat("weight_min_value"), weight_max_value=g.getfloat("weight_max_value"),
                bias_init_mean=g.getfloat("bias_init_mean"), bias_init_stdev=g.getfloat("bias_init_stdev"),
                bias_min_value=g.getfloat("bias_min_value"), bias_max_value=g.getfloat("bias_max_value"),
                activation_default=g.get("activation_default"), activation_options=g.get("activation_options").split(),
                activation_mutate_rate=g.getfloat("activation_mutate_rate"), connection_fraction=frac)


def synthetic_genome(preset, index, config_id=None, evolved=False, num_inputs=None):
    """Seeded genome shaped like `DefaultGenome.configure_new` under `preset`.

    initial_connection = partial_nodirect 0.8: all input->hidden and hidden->output pairs, shuffled,
    the first round(0.8*len) kept.  With `evolved=True` a few structural mutations are applied on top
    (disabled / deleted / hidden->hidden connections, split connections, non-unit response) so that the
    constant-folding and ordering rules of the flattener are exercised the way an evolved population does.
    """
    p = NEAT_PRESETS[preset] if isinstance(preset, str) else preset
    if config_id is None:
        config_id = PRESET_IDS.get(preset, 9) if isinstance(preset, str) else 9
    rng = random.Random(1000 * config_id + index)
    n_in = p["num_inputs"] if num_inputs is None else num_inputs
    n_out, n_hid = p["num_outputs"], p["num_hidden"]
    g = Genome(index)
    in_keys = [-i - 1 for i in range(n_in)]
    out_keys = list(range(n_out))
    hid_keys = list(range(n_out, n_out + n_hid))

    def clip(v, lo, hi):
        return max(lo, min(hi, v))

    def new_node(k):
        act = p["activation_default"]
        if rng.random() < p["activation_mutate_rate"]:
            act = rng.choice(p["activation_options"])
        g.nodes[k] = NodeGene(k, clip(rng.gauss(p["bias_init_mean"], p["bias_init_stdev"]),
                                      p["bias_min_value"], p["bias_max_value"]), 1.0, act, "sum")

    for k in out_keys + hid_keys:
        new_node(k)
    pairs = [(i, h) for i in in_keys for h in hid_keys] + [(h, o) for h in hid_keys for o in out_keys]
    if not hid_keys:
        pairs = [(i, o) for i in in_keys for o in out_keys]
    rng.shuffle(pairs)
    keep = int(round(len(pairs) * p["connection_fraction"]))

    def new_conn(key):
        g.connections[key] = ConnectionGene(
            key, clip(rng.gauss(p["weight_init_mean"], p["weight_init_stdev"]),
                      p["weight_min_value"], p["weight_max_value"]), True)

    for key in pairs[:keep]:
        new_conn(key)
    if evolved:
        order = {k: i for i, k in enumerate(hid_keys)}
        for _ in range(rng.randint(2, 8)):
            r = rng.random()
            keys = list(g.connections)
            if r < 0.25 and keys:
                g.connections[rng.choice(keys)].enabled = False
            elif r < 0.45 and keys:
                del g.connections[rng.choice(keys)]
            elif r < 0.7 and len(hid_keys) > 1:
                a, b = rng.sample(hid_keys, 2)
                if order[a] > order[b]:
                    a, b = b, a
                if (a, b) not in g.connections:
                    new_conn((a, b))
            elif r < 0.85 and keys:
                # split a connection with a new hidden node, as neat's mutate_add_node does
                ck = rng.choice(keys)
                src, dst = ck
                if src in order or src < 0:
                    nk = max(g.nodes) + 1
                    new_node(nk)
                    g.connections[ck].enabled = False
                    g.connections[(src, nk)] = ConnectionGene((src, nk), 1.0, True)
                    g.connections[(nk, dst)] = ConnectionGene((nk, dst), g.connections[ck].weight, True)
            else:
                k = rng.choice(hid_keys + out_keys)
                g.nodes[k].response = clip(rng.gauss(1.0, 0.5), -30.0, 30.0)
    return g


def synthetic_population(preset, n, evolved=False, num_inputs=None, start=0):
    return [(start + i, synthetic_genome(preset, start + i, evolved=evolved, num_inputs=num_inputs))
            for i in range(n)]


# ----------------------------------------------------------------------------- flattener
def required_for_output(inputs, outputs, connections):
    """neat.graphs.required_for_output restated: nodes whose value can reach an output, found layer by layer from the
    outputs backwards (the walk stops at a layer that holds only input pins)."""
    preds = {}
    for a, b in connections:
        preds.setdefault(b, []).append(a)
    needed = set(outputs)
    seen = set(outputs)
    inputs = set(inputs)
    frontier = list(outputs)
    while True:
        layer = set()
        for b in frontier:
            for a in preds.get(b, ()):
                if a not in seen:
                    layer.add(a)
        if not layer:
            return needed
        hidden = layer - inputs
        if not hidden:
            return needed
        needed |= hidden
        seen |= layer
        frontier = layer


_F32_ACT = {
    "sigmoid": lambda t: torch.sigmoid(5 * t),
    "tanh": lambda t: torch.tanh(2.5 * t),
    "abs": torch.abs,
    "gauss": lambda t: torch.exp(-5.0 * t ** 2),
    "identity": lambda t: t,
    "sin": torch.sin,
    "relu": torch.nn.functional.relu,
}


class FlatProgram:
    """Flat CPPN program: nodes in evaluation order; terms = (weight f64, source slot)."""

    def __init__(self):
        self.nodes = []  # (act, agg, term_begin, n_terms, bias, response)
        self.terms = []  # (weight, slot)
        self.out_slots = []
        self._n_slots = None

    @classmethod
    def from_bytes(cls, blob, n_slots):
        """A program that only exists in its packed form (what the C flattener returns)."""
        p = cls()
        p._bytes, p._n_slots = blob, n_slots
        return p

    @property
    def n_slots(self):
        return self._n_slots if self._n_slots is not None else SLOT_NODE0 + len(self.nodes)

    def to_bytes(self):
        if getattr(self, "_bytes", None) is not None:
            return self._bytes
        self._bytes = self._encode()
        return self._bytes

    def _encode(self):
        outs = list(self.out_slots)
        if len(outs) % 2:
            outs.append(0)
        b = [struct.pack("<4i", BLOB_MAGIC, len(self.nodes), len(self.terms), len(self.out_slots)),
             struct.pack("<%di" % len(outs), *outs)]
        for act, agg, t0, nt, bias, resp in self.nodes:
            b.append(struct.pack("<4i2d", act, agg, t0, nt, bias, resp))
        for wgt, slot in self.terms:
            b.append(struct.pack("<d2i", wgt, slot, 0))
        return b"".join(b)


def flatten_genome(genome, config, n_outputs=None):
    """genome -> FlatProgram.  `n_outputs` limits the outputs that are rendered (colour uses 0..2 of a
    6-output config, SURVEY.md "defects"); unreachable nodes cost nothing."""
    gc = config.genome_config
    in_keys, out_keys = list(gc.input_keys), list(gc.output_keys)
    if len(in_keys) != 2:
        raise ValueError("the CPPN render path takes exactly two leaves (x, y); got %d input keys "
                         "(cppn.py:198 asserts the same)" % len(in_keys))
    used_outs = out_keys if n_outputs is None else out_keys[:n_outputs]
    needed = required_for_output(in_keys, out_keys, genome.connections)
    out_set = set(out_keys)
    incoming = {k: [] for k in out_keys}
    for cg in genome.connections.values():
        if not cg.enabled:
            continue
        src, dst = cg.key
        if dst not in needed and src not in needed:
            continue
        if src in out_set:
            continue
        incoming.setdefault(dst, []).append((src, cg.weight))
        incoming.setdefault(src, [])

    prog = FlatProgram()
    nodes, terms_out, nodes_out = genome.nodes, prog.terms, prog.nodes
    # memo: node key -> int slot (value depends on x / y) or float32 tensor (input-independent, folded on the host)
    memo = {in_keys[0]: SLOT_X, in_keys[1]: SLOT_Y}

    def emit(act, agg, terms, bias, resp):
        t0 = len(terms_out)
        terms_out.extend(terms)
        nodes_out.append((act, agg, t0, len(terms), float(bias), float(resp)))
        return SLOT_NODE0 + len(nodes_out) - 1

    def visit(key):
        res = memo.get(key)
        if res is not None:
            return res
        gene = nodes[key]
        srcs = incoming[key]
        if not srcs:
            res = memo[key] = torch.full((_FOLD_N,), gene.bias)
            return res
        agg = gene.aggregation
        if agg not in AGG_IDS:
            raise KeyError("unsupported aggregation %r" % agg)
        vals = [(w, visit(s)) for s, w in srcs]
        n_var = 0
        for _, v in vals:
            if type(v) is int:
                n_var += 1
        if n_var == len(vals):          # the common case: every source depends on the inputs
            res = memo[key] = emit(ACT_IDS[gene.activation], AGG_IDS[agg], [(float(w), v) for w, v in vals],
                                   gene.bias, gene.response)
            return res
        is_sum = agg == "sum"
        if n_var == 0:
            acc = None
            for w, t in vals:
                term = w * t
                acc = term if acc is None else (acc + term if is_sum else acc * term)
            res = memo[key] = _F32_ACT[gene.activation](gene.response * acc + gene.bias)
            return res
        terms, prefix, seen_var = [], None, False
        for w, v in vals:
            if type(v) is not int:
                term = w * v
                if not seen_var:
                    prefix = term if prefix is None else (prefix + term if is_sum else prefix * term)
                else:
                    terms.append((float(term[0].item()), SLOT_ONE))
            else:
                if not seen_var and prefix is not None:
                    terms.append((float(prefix[0].item()), SLOT_ONE))
                seen_var = True
                terms.append((float(w), v))
        res = memo[key] = emit(ACT_IDS[gene.activation], AGG_IDS[agg], terms, gene.bias, gene.response)
        return res

    for k in used_outs:
        v = visit(k)
        if type(v) is not int:
            # constant output plane: identity node 1.0*(c*1.0)+0.0
            # (bit 30 tells the kernel the plane is float32 in the reference: gray images multiply it in fp32)
            v = emit(ACT_IDS["identity"], AGG_IDS["sum"], [(float(v[0].item()), SLOT_ONE)], 0.0, 1.0) | OUT_F32_CONST
        prog.out_slots.append(v)
    return prog


try:                                   # csrc/flatten.c, built by csrc/build.sh next to libeig.so
    from . import _flatten as _cflat
except ImportError:                    # not built: the Python flattener above is the same algorithm
    _cflat = None


def flatten_genome_fast(genome, config, n_outputs=None):
    """`flatten_genome` through the C extension (csrc/flatten.c, ~10x faster: the flattening of a generation is host time
    in front of every evaluation).  The result only carries the packed bytes.  Genomes whose constant sub-graphs need a
    transcendental float32 fold (torch's vectorised kernels, cppn.py:79-80 semantics) come back as None from C and take
    the Python flattener, which folds with torch itself."""
    if _cflat is not None:
        gc = config.genome_config
        r = _cflat.flatten(genome, list(gc.input_keys), list(gc.output_keys), n_outputs)
        if r is not None:
            return FlatProgram.from_bytes(r[0], r[1])
    return flatten_genome(genome, config, n_outputs=n_outputs)


def genome_fingerprint(genome):
    """Hash of everything `flatten_genome` reads from a genome (connection keys / weights / enabled flags, node
    attributes).  Two genomes with equal fingerprints flatten to the same program."""
    conns = tuple((k, c.weight, c.enabled) for k, c in genome.connections.items())
    nodes = tuple((k, n.bias, n.response, n.activation, n.aggregation) for k, n in genome.nodes.items())
    return hash((conns, nodes))


class ProgramCache:
    """genome id -> FlatProgram (SURVEY.md §8 f row 4) for the PYTHON flattener.  NEAT re-submits unchanged genomes (the
    elites of every species, `DefaultReproduction.reproduce`) generation after generation; a hit costs one fingerprint of
    the genome (~6x cheaper than the Python flattening) and re-uses the packed bytes.  With the C extension built
    (csrc/flatten.c) flattening itself is that cheap and `get` goes straight to it.  The fingerprint guards against a genome that was
    mutated in place under the same id.  Entries not touched for `keep` generations are dropped."""

    def __init__(self, keep=2):
        self.keep = keep
        self.generation = 0
        self.hits = 0
        self.misses = 0
        self.fast = 0          # genomes flattened by the C extension (no cache entry)
        self._entries = {}

    def get(self, genome_id, genome, config, n_outputs=None):
        gc = config.genome_config
        if _cflat is not None:
            # the C flattener costs about as much as the fingerprint that guards a cache entry: flatten every time; only
            # the genomes it hands back (constant sub-graphs that need torch's float32 kernels) go through the cache
            r = _cflat.flatten(genome, list(gc.input_keys), list(gc.output_keys), n_outputs)
            if r is not None:
                self.fast += 1
                return FlatProgram.from_bytes(r[0], r[1])
        key = (genome_id, n_outputs, tuple(gc.input_keys), tuple(gc.output_keys))
        fp = genome_fingerprint(genome)
        e = self._entries.get(key)
        if e is not None and e[0] == fp:
            e[2] = self.generation
            self.hits += 1
            return e[1]
        prog = flatten_genome(genome, config, n_outputs=n_outputs)
        prog.to_bytes()
        self._entries[key] = [fp, prog, self.generation]
        self.misses += 1
        return prog

    def flatten_population(self, population, config, n_outputs=None):
        """[(genome_id, genome)] -> [FlatProgram]; counts as one generation for eviction."""
        progs = [self.get(gid, g, config, n_outputs) for gid, g in population]
        self.end_generation()
        return progs

    def end_generation(self):
        self.generation += 1
        dead = [k for k, e in self._entries.items() if e[2] < self.generation - self.keep]
        for k in dead:
            del self._entries[k]

    def __len__(self):
        return len(self._entries)


def pack_population(programs):
    """[FlatProgram] -> (uint8 blob ndarray, int64 offsets[n+1], max slot count)."""
    chunks = [p.to_bytes() for p in programs]
    offsets = np.zeros(len(chunks) + 1, dtype=np.int64)
    for i, c in enumerate(chunks):
        offsets[i + 1] = offsets[i] + len(c)
    blob = np.frombuffer(b"".join(chunks), dtype=np.uint8).copy() if chunks else np.zeros(0, np.uint8)
    max_slots = max([p.n_slots for p in programs], default=SLOT_NODE0)
    return blob, offsets, max_slots
