"""Stand-in for the reference's ``network`` module (``erweitern_55/network.py``): the same ``Network`` class name
backed by libe55.so, so ``import network; network.Network()`` in ``mcts.py`` / ``usi.py`` / ``client.py`` gets the
B200 evaluator.  Put this directory on ``sys.path`` ahead of the reference's own."""
from erweitern_55_b200.network import Network  # noqa: F401
