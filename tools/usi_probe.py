"""Search speed of ONE game in latency mode (USI / the reference's mcts.py): simulations per second with the
reference-style Python loop (16 select_leaf + get_inputs + predict + 16 evaluate + 16 backpropagate per batch) and
with the native batch step (one C call per leaf batch), both on the same Network."""
import sys, time
sys.path.insert(0, ".")
from erweitern_55_b200 import mcts, weights as W
from erweitern_55_b200.minishogi import Position
from erweitern_55_b200.network import Network

net = Network(precision="bf16", weights=W.init_weights(seed=55))


class ReferenceStyle:
    def __init__(self, n):
        self.get_input, self.get_inputs, self.predict = n.get_input, n.get_inputs, n.predict


for name, nn in (("native batch step", net), ("reference-style Python loop", ReferenceStyle(net))):
    cfg = mcts.Config()
    cfg.simulation_num = 10 ** 9
    cfg.reuse_tree = False
    search = mcts.MCTS(cfg)
    p = Position()
    search.run(p, nn, timelimit=200)
    t0 = time.time()
    root = search.run(p, nn, timelimit=2000)
    dt = time.time() - t0
    sims = search.mcts.get_playouts(root, True)
    print(f"{name}: {sims} simulations in {dt:.2f} s = {sims / dt:.0f} simulations/s ({dt / (sims / 16) * 1e6:.0f} us per batch of 16)", flush=True)
