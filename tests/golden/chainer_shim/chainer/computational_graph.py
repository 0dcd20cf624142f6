"""chainer.computational_graph: imported by call_prednet.py, only used for graph dumps while training."""
