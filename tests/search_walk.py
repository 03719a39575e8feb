"""select_kernel / expand_kernel of the GPU-resident self-play against oracle/mcts_oracle.py, selection by selection.

One game goes through the debug stepper (one select round, the test in the network's place, one expand round, ...) of
either build of the kernels -- the CUDA one behind the C ABI (``e55_gpu_search_*``, tests/test_gpu_parity.py) or the host
one (tests/native/gs_stepper.cpp, tests/test_search_on_host.py) -- with uniform priors over a pseudo-random subset of the
legal moves and values k/8 (exact in float32).  Every selection of every round must take the oracle's path, every network
slot must hold the oracle position's legal moves, and the record of each search -- visit counts, Q, the move played --
must be the oracle's, also after the chosen subtree has become the tree (reuse_tree)."""
import ctypes as C
import zlib

import numpy as np

START = "rbsgk/4p/5/P4/KGSBR b - 1"


class CudaStepper:
    """The stepper of libe55 on a Network (the device build)."""

    def __init__(self, net):
        self.net, self.lib = net, net._lib

    def set_cap(self, enable, full, fast, frac):
        self.net._check(self.lib.e55_gpu_selfplay_set_cap(self.net._ctx, int(enable), full, fast, frac))

    def open(self, packed_prefix, simulations, reuse_tree):
        h = C.c_void_p()
        arr = (C.c_uint16 * max(1, len(packed_prefix)))(*packed_prefix)
        self.net._check(self.lib.e55_gpu_search_open(self.net._ctx, arr, len(packed_prefix), simulations, reuse_tree, C.byref(h)))
        return h

    def select(self, h, *bufs):
        self.net._check(self.lib.e55_gpu_search_select(h, *bufs))

    def expand(self, h, priors, value, info, record):
        self.net._check(self.lib.e55_gpu_search_expand(h, priors.ctypes.data, value.ctypes.data, info, record))

    def close(self, h):
        self.lib.e55_gpu_search_close(h)


class HostStepper:
    """The same calls on the host build of the kernels (tests/native/gs_stepper.cpp)."""

    def __init__(self, path):
        self.lib = C.CDLL(path)
        self.cap = (0, 800, 128, 0.25)
        self.lib.gs_search_open.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]
        self.lib.gs_search_select.argtypes = [C.c_void_p] * 9
        self.lib.gs_search_expand.argtypes = [C.c_void_p] * 5
        self.lib.gs_search_close.argtypes = [C.c_void_p]
        self.lib.gs_search_close.restype = None

    def set_cap(self, enable, full, fast, frac):
        self.cap = (int(enable), full, fast, frac)

    def open(self, packed_prefix, simulations, reuse_tree):
        h = C.c_void_p()
        arr = (C.c_uint16 * max(1, len(packed_prefix)))(*packed_prefix)
        assert self.lib.gs_search_open(arr, len(packed_prefix), simulations, reuse_tree, *self.cap, C.byref(h)) == 0
        return h

    def select(self, h, *bufs):
        assert self.lib.gs_search_select(h, *bufs) == 0

    def expand(self, h, priors, value, info, record):
        assert self.lib.gs_search_expand(h, priors.ctypes.data, value.ctypes.data, info, record) == 0

    def close(self, h):
        self.lib.gs_search_close(h)


def packed_move_sfen(mv):
    from erweitern_55_b200.network import packed_move_sfen as f
    return f(mv)


def _buffers():
    return dict(info=(C.c_int32 * 8)(), leaf=(C.c_int32 * 24)(), slot=(C.c_int32 * 24)(), plen=(C.c_int32 * 24)(),
                path=(C.c_uint16 * (24 * 96))(), cnt=(C.c_int32 * 16)(), mvs=(C.c_uint16 * (16 * 256))(), idx=(C.c_uint16 * (16 * 256))())


def pack_prefix(stepper, prefix):
    """The prefix moves as the kernels pack them: taken from the move lists of a stepper opened on the shorter prefix."""
    packed = []
    for usi in prefix:
        h = stepper.open(packed, 16, 0)
        b = _buffers()
        stepper.select(h, b["info"], b["leaf"], b["slot"], b["plen"], b["path"], b["cnt"], b["mvs"], b["idx"])
        table = {packed_move_sfen(b["mvs"][k]): b["mvs"][k] for k in range(b["cnt"][0])}
        stepper.close(h)
        assert usi in table, (usi, sorted(table))
        packed.append(table[usi])
    return packed


def fake_network(key, n_moves):
    hsh = zlib.crc32(key.encode())
    keep = [((hsh >> (i % 24)) ^ i) & 3 != 0 for i in range(n_moves)]
    if not any(keep):
        keep[hsh % n_moves] = True
    return keep, (hsh % 9) / 8.0


def walk_like_the_oracle(stepper, prefix, simulations, searches, full, expect_pruning=True):
    """Returns (selections compared, terminal leaves met, plies played, whether the game ended).  ``expect_pruning``: a full
    search must have had forced playouts for the target pruning to take back (true of searches of a few hundred simulations)."""
    from oracle import minishogi_oracle as MO
    from oracle.mcts_oracle import Tree

    packed_prefix = pack_prefix(stepper, prefix)
    played = list(prefix)

    def oracle_position():
        pos = MO.Position(START)
        for m in played:
            pos.do_move(m)
        return pos

    stepper.set_cap(full, simulations, 128, 1.0)                       # frac 1: every search a full one
    h = stepper.open(packed_prefix, simulations, 1)
    b = _buffers()
    info, leaf, slot, plen, path, cnt, mvs, idx = (b[k] for k in ("info", "leaf", "slot", "plen", "path", "cnt", "mvs", "idx"))
    priors = np.zeros((16, 256), dtype=np.float32)
    value = np.zeros(16, dtype=np.float32)
    record = (C.c_uint16 * 16384)()
    oracle = Tree()
    total_selections = terminals = 0
    search = 0
    for search in range(searches):
        ply0 = len(played)
        for rnd in range(400):
            stepper.select(h, info, leaf, slot, plen, path, cnt, mvs, idx)
            assert info[0] == ply0
            n_sel = info[2]
            assert n_sel > 0
            sel = []
            for s in range(n_sel):                                    # the round's selections, virtual losses piling up
                pos_o = oracle_position()
                leaf_o = oracle.select_leaf(oracle.root, pos_o, full)
                want = [packed_move_sfen(path[s * 96 + i]) for i in range(plen[s])]
                node, walk = leaf_o, []
                while node.parent is not None:
                    walk.append(node.parent.moves[node.parent_edge]); node = node.parent
                assert want == walk[::-1], (search, rnd, s, want, walk[::-1])
                assert leaf[s] >= 0
                sel.append((leaf_o, pos_o))
            priors[:] = 0
            for s, (leaf_o, pos_o) in enumerate(sel):                 # the network's answers (or the rules' verdict)
                k = slot[s]
                if leaf_o.expanded or leaf_o.terminal:
                    assert k < 0
                    continue
                legal = sorted(pos_o.generate_moves())
                if k < 0:                                             # the kernel found the game over at this leaf
                    oracle.evaluate(leaf_o, pos_o, legal, [], 0.0)
                    assert leaf_o.terminal, (search, rnd, s)
                    terminals += 1
                    continue
                moves = [packed_move_sfen(mvs[k * 256 + i]) for i in range(cnt[k])]
                assert sorted(moves) == legal
                keep, v01 = fake_network(pos_o.board_sfen(), len(moves))
                kept = np.float32(sum(keep))
                priors[k, :len(moves)] = np.where(keep, np.float32(1) / kept, np.float32(0))
                value[k] = 2 * v01 - 1
                oracle.evaluate(leaf_o, pos_o, moves, [0.0 if kp else -1e30 for kp in keep], v01)
                assert leaf_o.expanded
            for leaf_o, _ in sel:
                oracle.backpropagate(leaf_o)
            total_selections += n_sel
            stepper.expand(h, priors, value, info, record)
            if info[0] != ply0 or info[7]:
                break
        else:
            raise AssertionError("the search did not end")
        # the ply the kernels recorded: move, sum_N, Q, visit counts == the oracle's tree
        base, words = (4, info[6]) if info[7] else (0, info[6])
        w, last = base, None
        while w < base + words:
            k = record[w + 3]
            last = (record[w] & 0x7FFF, record[w + 1], record[w + 2] / 65535.0,
                    {packed_move_sfen(record[w + 4 + 2 * i]): record[w + 5 + 2 * i] for i in range(k)})
            assert not (record[w] >> 15)                              # (every ply here is a learning target)
            w += 4 + 2 * k
        mv, sum_n, q, visits = last
        assert visits == oracle.visits(target_pruning=full) and sum_n == sum(visits.values())
        if full and expect_pruning:
            assert visits != oracle.visits()                          # (the pruning did take forced playouts back)
        assert abs(q - sum(float(x) for x in oracle.root.w) / sum(oracle.root.n_edge)) < 2e-5
        best = max(range(len(oracle.root.moves)), key=lambda i: (oracle.root.n_edge[i], -i))
        assert packed_move_sfen(mv) == oracle.root.moves[best]
        if info[7]:
            break                                                     # the move ended the game
        kept = oracle.root.child[best] if not full else None          # a full search starts from a fresh tree
        played.append(oracle.root.moves[best])
        oracle.root = kept if (kept is not None and kept.expanded and not kept.terminal) else Tree().root
        oracle.root.parent, oracle.root.parent_edge = None, -1
    stepper.close(h)
    stepper.set_cap(0, 800, 128, 0.25)
    ended = bool(info[7])
    assert total_selections >= simulations * (search + 1) - 24 * (search + 1) or ended
    assert len(played) - len(prefix) == search + (0 if ended else 1)      # every search ended in a move
    return total_selections, terminals, played, ended
