"""USI front-end with the reference's protocol behaviour (``erweitern_55/usi.py:13-164``) on the B200 evaluator.

Same commands, replies and search policy: ``isready`` builds the network (optionally loading ``--weight_file``) and a
tree-reusing search with an unbounded simulation count; ``go btime/wtime/byoyomi`` answers ``bestmove resign`` without
legal moves, plays a mate found by the 7-ply mate search at once, otherwise thinks for a tenth of the remaining time
(all of it plus the byoyomi when that is shorter than the byoyomi), samples among the top moves for the first
``softmax_sampling_moves`` plies, and afterwards keeps the tree warm (or ponders with ``setoption name ponder value
true``) on a background thread that the next ``position`` command stops.  Searching runs in latency mode: each leaf
batch of 16 is one direct forward (about 80 us on a B200).

Deliberate differences: ``position startpos moves ...`` applies the moves (the reference drops them), a ``go`` without
clock arguments thinks for one second instead of raising ``KeyError``, and ``quit`` leaves ``start()`` instead of
calling ``os._exit`` so the loop can be embedded and tested.
"""
from __future__ import annotations

import sys
import threading
from optparse import OptionParser

from . import mcts, minishogi


class USI:
    def __init__(self, weight_file=None, nn=None, stdin=None, stdout=None):
        self.weight_file = weight_file
        self.nn = nn
        self.search = None
        self.position = None
        self.ponder_thread = None
        self.option = {'ponder': False, 'softmax_sampling_moves': 30}
        self._in = stdin if stdin is not None else sys.stdin
        self._out = stdout if stdout is not None else sys.stdout

    def _say(self, text):
        print(text, file=self._out, flush=True)

    # -- engine state ------------------------------------------------------------------------------------------
    def isready(self):
        if self.nn is None:
            from .network import Network
            self.nn = Network()
            if self.weight_file is not None:
                self.nn.load(self.weight_file)
        self.config = mcts.Config()
        self.config.simulation_num = int(1e9)
        self.config.reuse_tree = True
        if self.search is None:
            self.search = mcts.MCTS(self.config)
        self.search.clear()
        self.position = minishogi.Position()
        self.ponder_thread = None

    def ponder_start(self):
        """Called with the opponent to move: set the new root (tree reuse), and search it when pondering is on."""
        self.ponder_thread = threading.Thread(
            target=self.search.run, args=(self.position, self.nn, 0, True, not self.option['ponder']))
        self.ponder_thread.start()

    def ponder_stop(self):
        if self.ponder_thread is not None:
            self.search.stop()
            self.ponder_thread.join()
            self.ponder_thread = None

    # -- commands ------------------------------------------------------------------------------------------------
    def _setoption(self, words):
        key, value = words[2], words[4]
        if value in ('true', 'True'):
            value = True
        elif value in ('false', 'False'):
            value = False
        elif value.isdigit():
            value = int(value)
        self.option[key.lower()] = value

    def _position(self, words):
        self.ponder_stop()
        if words[1] == 'sfen':
            self.position.set_sfen(' '.join(words[2:]))
        elif words[1] == 'startpos':
            self.position.set_start_position()
            if len(words) > 3 and words[2] == 'moves':
                for move in words[3:]:
                    self.position.do_move(minishogi.Move(move))
        else:
            self._say('ERROR: Unknown protocol.')

    def _go(self, words):
        clock = {'btime': 0, 'wtime': 0, 'byoyomi': 0}
        for i, word in enumerate(words[:-1]):
            if word in clock:
                clock[word] = int(words[i + 1])
        if len(self.position.generate_moves()) == 0:
            self._say('bestmove resign')
            return
        checkmate, best_move = self.position.solve_checkmate_dfs(7)
        if not checkmate:
            remain = clock['btime'] if self.position.get_side_to_move() == 0 else clock['wtime']
            think_time = remain // 10
            if think_time < clock['byoyomi']:
                think_time = remain + clock['byoyomi']
            if think_time <= 0:
                think_time = 1000
            self._say('info string think time {}'.format(think_time))
            root = self._search(think_time)
            if self.position.get_ply() < self.option['softmax_sampling_moves']:
                best_move = self.search.softmax_sample_among_top_moves(root)
            else:
                best_move = self.search.best_move(root)
        self._say('bestmove {}'.format(best_move))
        self.position.do_move(best_move)
        self.ponder_start()

    def _search(self, think_time):
        # the search prints its `info` lines with print(); send them where this front-end writes
        if self._out is sys.stdout:
            return self.search.run(self.position, self.nn, think_time, True)
        import contextlib
        with contextlib.redirect_stdout(self._out):
            return self.search.run(self.position, self.nn, think_time, True)

    def handle(self, line):
        """Process one protocol line; False after ``quit``."""
        words = line.split()
        if not words:
            return True
        cmd = words[0]
        if cmd == 'usi':
            self._say('id name erweitern_55')
            self._say('id author nyashiki')
            self._say('usiok')
        elif cmd == 'setoption':
            self._setoption(words)
        elif cmd == 'position':
            self._position(words)
        elif cmd == 'isready':
            self.isready()
            self._say('readyok')
        elif cmd == 'usinewgame':
            pass
        elif cmd == 'go':
            self._go(words)
        elif cmd == 'd':
            self._say(self.position.pretty().rstrip('\n'))
        elif cmd == 'quit':
            self.ponder_stop()
            return False
        else:
            self._say('ERROR: Unknown command. {}'.format(cmd))
        return True

    def start(self):
        for line in self._in:
            if not self.handle(line.strip()):
                break


if __name__ == '__main__':
    parser = OptionParser()
    parser.add_option('-w', '--weight_file', dest='weight_file', default=None,
                      help='Weights of neural network parameters')
    (options, args) = parser.parse_args()
    USI(options.weight_file).start()
