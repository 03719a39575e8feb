"""Stand-in for the external ``minishogilib`` module (Rust, v0.6.12) so that the reference's own ``mcts.py``,
``selfplay.py`` and ``main.py`` import and run unmodified: put this directory on ``sys.path`` ahead of everything
else.  Exposes what those files touch: ``Position``, ``MCTS``, ``Move``, ``__version__``."""
from erweitern_55_b200.minishogi import MCTS, Move, Position, __version__  # noqa: F401
