"""Test helper: seeded random feed-forward genomes with the irregularities evolution produces (disabled connections,
hidden->hidden chains, product aggregation, nodes without inputs, dangling nodes, connections leaving output nodes,
non-unit responses), beyond what `genome.synthetic_genome` builds."""
import random

from evolutionary_illusion_generator_b200 import genome as G


def fuzz_genome(seed, n_out=3, max_hidden=12):
    rng = random.Random(seed)
    g = G.Genome(seed)
    outs = list(range(n_out))
    hidden = list(range(n_out, n_out + rng.randint(0, max_hidden)))
    acts = sorted(G.ACT_IDS)
    for k in outs + hidden:
        g.nodes[k] = G.NodeGene(k, bias=max(-30.0, min(30.0, rng.gauss(0, 1.5))),
                                response=1.0 if rng.random() < 0.6 else rng.uniform(-2, 2),
                                activation=rng.choice(acts), aggregation="sum" if rng.random() < 0.8 else "prod")
    order = [-1, -2] + hidden + outs          # a connection only runs forward in this order => feed-forward
    pos = {k: i for i, k in enumerate(order)}

    def connect(a, b, p_enabled=0.85):
        g.connections[(a, b)] = G.ConnectionGene((a, b), rng.gauss(0, 2.0), enabled=rng.random() < p_enabled)

    for b in hidden + outs:
        for a in order[:pos[b]]:
            if a in outs:
                continue
            if rng.random() < (0.35 if a < 0 else 0.2):
                connect(a, b)
    for o in outs:                             # connections that leave an output node are dropped by create_cppn
        for b in hidden:
            if rng.random() < 0.05:
                connect(o, b)
    return g
