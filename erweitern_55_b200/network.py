"""Drop-in ``Network`` for erweitern_55's search code, backed by libe55.so (sm_100a CUDA kernels).

Mirrors the public surface of the reference class ``erweitern_55/network.py:16-286`` that ``mcts.py``,
``selfplay.py``, ``usi.py``, ``client.py`` and ``train.py`` touch: ``predict`` (network.py:201-216),
``get_input/get_inputs/zero_inputs`` (network.py:255-274,283-286), ``get_weights/set_weights``
(network.py:228-243), ``load/save`` (network.py:218-226,245-253), ``iter`` (network.py:276-281),
``step`` (network.py:148-199; training is out of scope and raises), attributes ``input_shape``,
``network_type`` and ``model`` (``client.py:51`` calls ``nn.model.set_weights`` directly).
TensorFlow is replaced by the C ABI in ``include/e55.h``; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi, hostmem
from .weights import IN_PLANES, POLICY_SIZE, adapt_weights, infer_arch, init_weights, weight_spec

_PRECISIONS = {"fp32": _capi.PRECISION_FP32, "bf16": _capi.PRECISION_BF16_TC}


def _parse_device(device):
    if isinstance(device, int):
        return device
    d = str(device).lower()
    if d in ("gpu", "cuda"):
        return 0
    if d.startswith("cuda:") or d.startswith("gpu:"):
        return int(d.split(":", 1)[1])
    raise ValueError(
        f"device {device!r} is not supported: this build evaluates on a B200 only "
        "(the reference's 'cpu'/'tpu' session configs, network.py:21-42, have no counterpart)")


class _ModelShim:
    """Stands in for ``Network.model`` (a keras.Model in the reference) for the calls callers make."""

    def __init__(self, net):
        self._net = net

    def set_weights(self, weights):  # client.py:51
        self._net.set_weights(weights)

    def get_weights(self):
        return self._net.get_weights()

    def predict(self, images, batch_size=None, verbose=0, steps=None):  # network.py:213-214
        return list(self._net.predict(images))


class Network:
    def __init__(self, device="gpu", blocks=7, filters=128, in_planes=IN_PLANES, max_batch=1024,
                 precision="bf16", weights=None, seed=None):
        self._lib = _capi.load()
        self._device = _parse_device(device)
        self.network_type = "AlphaZero"                 # network.py:84
        self.input_shape = [in_planes, 5, 5]            # network.py:85
        self._blocks, self._filters, self._in_planes = blocks, filters, in_planes
        self._ctx = C.c_void_p()
        self._iterations = 0
        self._weights = None
        self._layout = None                             # (plane_perm, plane_scale, policy_perm) of set_layout_adapter
        self._service_cap = 0
        rc = self._lib.e55_create(C.byref(self._ctx), self._device, blocks, filters, in_planes, max_batch)
        if rc != _capi.E55_OK:
            _capi.check(None, rc)
        self.model = _ModelShim(self)
        self.set_precision(precision)
        if weights is None:
            # the reference starts from Keras' random initialisers (network.py:50-53)
            if seed is None:
                seed = int(np.random.randint(0, 2**31 - 1))
            weights = init_weights(blocks, filters, in_planes, seed=seed)
        self.set_weights(weights)
        # warm-up predict on a float64 random input, as the reference does (network.py:71-73)
        self.predict(np.random.rand(*([1] + self.input_shape)))

    # ---- lifecycle ---------------------------------------------------------------------------
    def close(self):
        ctx, self._ctx = self._ctx, C.c_void_p()
        if ctx:
            self._lib.e55_destroy(ctx)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        _capi.check(self._ctx, rc)

    # ---- configuration -------------------------------------------------------------------------
    def set_precision(self, precision):
        if precision not in _PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_PRECISIONS)}")
        self._check(self._lib.e55_set_precision(self._ctx, _PRECISIONS[precision]))
        self.precision = precision

    # ---- weights (network.py:228-243) ----------------------------------------------------------
    def get_weights(self):
        return [w.copy() for w in self._weights]

    def set_layout_adapter(self, plane_perm=None, plane_scale=None, policy_perm=None):
        """Declare that the weights handed to :meth:`set_weights` / :meth:`load` were trained on another input-plane /
        policy-channel layout than this package's native engine produces (i.e. by the reference with minishogilib, whose
        layout is not visible in the reference tree): see :func:`erweitern_55_b200.weights.adapt_weights` for the three
        tables.  The permutation is folded into the first and last tensors when weights are installed -- predict() costs
        the same -- and :meth:`get_weights` keeps returning the list in the checkpoint's own layout."""
        self._layout = None if (plane_perm is None and plane_scale is None and policy_perm is None) else (plane_perm, plane_scale, policy_perm)
        if self._weights is not None:
            self.set_weights(self._weights)

    def set_weights(self, weights):
        original = [np.ascontiguousarray(np.asarray(w, dtype=np.float32)) for w in weights]
        ws = original if self._layout is None else adapt_weights(original, *self._layout)
        blocks, filters, in_planes = infer_arch(ws)
        if (blocks, filters, in_planes) != (self._blocks, self._filters, self._in_planes):
            raise ValueError(f"weight list describes a {blocks}x{filters} tower on {in_planes} planes, "
                             f"this Network is {self._blocks}x{self._filters} on {self._in_planes}")
        n = len(ws)
        data = (C.POINTER(C.c_float) * n)(*[w.ctypes.data_as(C.POINTER(C.c_float)) for w in ws])
        shape_arrays = [(C.c_int64 * w.ndim)(*w.shape) for w in ws]
        shapes = (C.POINTER(C.c_int64) * n)(*[C.cast(s, C.POINTER(C.c_int64)) for s in shape_arrays])
        ndims = (C.c_int * n)(*[w.ndim for w in ws])
        self._check(self._lib.e55_set_weights(self._ctx, n, data, shapes, ndims))
        self._weights = original

    def load_pickled_weights(self, blob):
        """Install weights from the reference's wire format: ``_pickle.dumps(nn.get_weights(), protocol=4)``
        as the trainer serves it over HTTP (train.py:50,96-99) and the client applies it (client.py:47-51)."""
        import pickle
        self.set_weights(pickle.loads(blob))

    def dump_pickled_weights(self):
        import pickle
        return pickle.dumps(self.get_weights(), protocol=4)

    def load(self, filepath):
        """Load weights from a Keras HDF5 file (``.h5`` / ``.hdf5``: what the reference's ``save`` writes with
        ``keras.Model.save``, network.py:245-253, or a ``save_weights`` file; read by :mod:`erweitern_55_b200.hdf5`,
        optimizer state ignored), from :meth:`save`'s ``.npz``, or from a pickled ``get_weights()`` list (``.pkl``)."""
        path = str(filepath)
        if path.endswith((".pkl", ".pickle")):
            with open(filepath, "rb") as f:
                self.load_pickled_weights(f.read())
            return
        if path.endswith((".h5", ".hdf5", ".keras")):
            from . import hdf5
            if self._layout is None:
                import warnings
                warnings.warn("loading a Keras checkpoint without a layout adapter: it matches this package's native engine "
                              "(minishogi.Position / MCTS, GPU self-play) only if it was trained on this package's plane and "
                              "policy-channel layout; for checkpoints trained by the reference with minishogilib call "
                              "set_layout_adapter() first (predict() on the caller's own planes is unaffected)", stacklevel=2)
            self.set_weights(hdf5.read_keras_weights(filepath))
            return
        with np.load(filepath) as z:
            n = len([k for k in z.files if k.startswith("arr_")])
            self.set_weights([z[f"arr_{i}"] for i in range(n)])
            if "iterations" in z.files:
                self._iterations = int(z["iterations"])

    def save(self, filepath):
        """``.h5`` / ``.hdf5``: a Keras ``save_weights``-layout HDF5 file (``model.load_weights`` reads it in the
        reference; there is no optimizer state or model config to store); otherwise ``.npz``."""
        path = str(filepath)
        if path.endswith((".h5", ".hdf5")):
            from . import hdf5
            from .weights import keras_layers
            hdf5.write_keras_weights(filepath, self._weights, keras_layers(self._blocks, self._filters, self._in_planes))
            return
        np.savez(filepath, *self._weights, iterations=np.int64(self._iterations))

    # ---- inputs (network.py:255-274,283-286) ----------------------------------------------------
    def get_input(self, position, dim_4=False):
        shape = [1] + self.input_shape if dim_4 else self.input_shape
        return np.reshape(position.to_alphazero_input(), shape)

    def predict_positions(self, positions):
        """``predict(get_inputs(positions))`` without the float planes: positions of this package's engine are encoded
        bit-packed in one native call (identical results, 12.5x fewer bytes over PCIe)."""
        from .minishogi import encode_packed_batch
        return self.predict_packed(*encode_packed_batch(positions))

    def get_inputs(self, positions):
        if len(positions) and all(hasattr(p, "to_packed_input") and hasattr(p, "_h") for p in positions):
            # positions of this package's engine: one native encode for all of them instead of a Python list of 6 650
            # floats per position (the same planes: to_alphazero_input() is defined as this expansion)
            from .minishogi import encode_packed_batch
            from .packing import unpack_planes
            return np.ascontiguousarray(unpack_planes(*encode_packed_batch(positions)), dtype=np.float32)
        ins = np.zeros([len(positions)] + self.input_shape, dtype="float32")
        for i, position in enumerate(positions):
            ins[i] = self.get_input(position)
        return ins

    def zero_inputs(self, batch_size):
        ins = hostmem.empty([batch_size] + self.input_shape)   # page-locked when large: predict() can split it between cores and copy engine
        ins.fill(0)
        return ins

    # ---- evaluation (network.py:201-216) --------------------------------------------------------
    def _as_batch(self, images):
        x = np.ascontiguousarray(np.asarray(images, dtype=np.float32))
        if x.ndim != 4 or list(x.shape[1:]) != self.input_shape:
            raise ValueError(f"expected images of shape [N,{self.input_shape[0]},5,5], got {list(x.shape)}")
        if x.shape[0] == 0:
            raise ValueError("empty batch")
        return x

    def predict(self, images):
        """policy float32 ``[N,1725]`` raw logits, value float32 ``[N,1]`` (tanh); whole input = one batch.
        While the aggregation service runs (:meth:`start_service`) calls from many threads are coalesced."""
        x = self._as_batch(images)
        n = x.shape[0]
        policy = hostmem.empty((n, POLICY_SIZE))        # page-locked from 38 positions up: the copy engine writes into it
        value = np.empty((n, 1), dtype=np.float32)
        fn = self._lib.e55_service_predict if (self._service_cap and n <= self._service_cap) else self._lib.e55_predict
        self._check(fn(self._ctx, x.ctypes.data, n, policy.ctypes.data, value.ctypes.data))
        return policy, value

    # ---- cross-game leaf aggregation --------------------------------------------------------------
    def start_service(self, max_batch=1024, max_wait_us=200):
        """Coalesce predict() calls of many search threads (one per game) into large GPU batches."""
        self._check(self._lib.e55_service_start(self._ctx, max_batch, max_wait_us))
        self._service_cap = max_batch

    def stop_service(self):
        self._service_cap = 0
        self._check(self._lib.e55_service_stop(self._ctx))

    def service_stats(self):
        b, p = C.c_int64(), C.c_int64()
        self._check(self._lib.e55_service_stats(self._ctx, C.byref(b), C.byref(p)))
        return int(b.value), int(p.value)

    def selfplay(self, threads, total_moves, simulations=800, leaf_batch=16, mate_depth=7, seed=1, **kw):
        """Real self-play on the running service with the native minishogi engine + MCTS (``selfplay.run``'s loop,
        selfplay.py:10-113, with client.py's search settings): a dict of moves, evals, games, seconds, mean game length
        and mean GPU batch.  See :func:`erweitern_55_b200.selfplay.run_many` for the game records."""
        from . import selfplay as _selfplay
        stats, _ = _selfplay.run_many(self, threads, total_moves, simulations, leaf_batch, mate_depth, seed=seed,
                                      records=False, **kw)
        return stats

    def gpu_selfplay(self, games, total_moves, simulations=800, mate_depth=7, reuse_tree=True, use_dirichlet=True, num_sampling_moves=10,
                     max_moves=512, seed=1, records=0, playout_cap_oscillation=None):
        """GPU-resident self-play (``csrc/e55_gpu_selfplay.cuh``): ``games`` games at once with tree search, move
        generation and input encoding on the GPU (the host only answers the mate searches), until ``total_moves`` plies
        were played.  Returns ``(stats, game_records)``; ``game_records`` holds the first ``records`` finished games as
        :class:`erweitern_55_b200.gamerecord.GameRecord` -- what ``selfplay.run`` returns (selfplay.py:10-113)."""
        import time
        cap = playout_cap_oscillation or {"enable": False, "N": 800, "n": 128, "frac": 0.25}     # selfplay.py:10
        self._check(self._lib.e55_gpu_selfplay_set_cap(self._ctx, int(bool(cap["enable"])), int(cap["N"]), int(cap["n"]), float(cap["frac"])))
        st = _capi.SelfplayStats()
        words = 16384
        buf = np.zeros((max(records, 1), words), dtype=np.uint16)
        n_rec = C.c_int(0)
        self._check(self._lib.e55_gpu_selfplay_run(self._ctx, games, total_moves, simulations, mate_depth, int(reuse_tree), int(use_dirichlet),
                                                   num_sampling_moves, max_moves, seed, C.byref(st), buf.ctypes.data if records else None,
                                                   records, C.byref(n_rec) if records else None))
        stats = {"seconds": st.seconds, "moves": st.moves, "evals": st.evals, "games": st.games,
                 "mean_game_plies": st.mean_game_plies, "slots_per_launch": st.mean_gpu_batch,
                 "moves_per_s": st.moves / st.seconds, "evals_per_s": st.evals / st.seconds,
                 "tree_overflows": st.tree_overflows, "searches_suspended": st.searches_suspended,
                 "mate_searches": st.mate_searches, "mate_exhausted": st.mate_exhausted, "records_truncated": st.records_truncated}
        if st.tree_overflows:
            raise RuntimeError(f"GPU self-play abandoned {st.tree_overflows} descents on a full tree pool: its searches are not the reference's")
        return stats, decode_gpu_records(buf, n_rec.value, int(time.time()))

    def bench_selfplay(self, workers, moves, simulations=800, leaf_batch=16):
        """Synthetic self-play load on the running service -> (moves/s, evals/s, mean GPU batch)."""
        m, e, mb = C.c_double(), C.c_double(), C.c_double()
        self._check(self._lib.e55_bench_selfplay(self._ctx, workers, moves, simulations, leaf_batch,
                                                 C.byref(m), C.byref(e), C.byref(mb)))
        return float(m.value), float(e.value), float(mb.value)

    def predict_packed(self, bits, scale):
        """Same evaluation from bit-packed planes (:mod:`erweitern_55_b200.packing`): ``bits`` uint32 ``[N,C]``,
        ``scale`` float32 ``[N,C]``; 12.5x fewer bytes over PCIe than the float32 planes, identical results."""
        b = np.ascontiguousarray(np.asarray(bits, dtype=np.uint32))
        sc = np.ascontiguousarray(np.asarray(scale, dtype=np.float32))
        if b.ndim != 2 or b.shape[1] != self._in_planes or sc.shape != b.shape or b.shape[0] == 0:
            raise ValueError(f"expected bits and scale of shape [N,{self._in_planes}]")
        n = b.shape[0]
        policy = np.empty((n, POLICY_SIZE), dtype=np.float32)
        value = np.empty((n, 1), dtype=np.float32)
        fn = self._lib.e55_service_predict_packed if (self._service_cap and n <= self._service_cap) else self._lib.e55_predict_packed
        self._check(fn(self._ctx, b.ctypes.data, sc.ctypes.data, n, policy.ctypes.data, value.ctypes.data))
        return policy, value

    def predict_packed_moves(self, bits, scale, move_offset, move_index):
        """Compact legal-move tail: ``move_index[move_offset[i]:move_offset[i+1]]`` are the policy indices of position
        ``i``'s legal moves; returns ``(priors, value)`` with ``priors`` the flat float32 array of the per-position
        softmax over exactly those logits (what ``minishogilib.MCTS.evaluate`` derives from a policy row,
        mcts.py:60,105-106) -- a few dozen floats per position cross PCIe instead of 1725."""
        b = np.ascontiguousarray(np.asarray(bits, dtype=np.uint32))
        sc = np.ascontiguousarray(np.asarray(scale, dtype=np.float32))
        off = np.ascontiguousarray(np.asarray(move_offset, dtype=np.int32))
        idx = np.ascontiguousarray(np.asarray(move_index, dtype=np.uint16))
        if b.ndim != 2 or b.shape[1] != self._in_planes or sc.shape != b.shape or b.shape[0] == 0:
            raise ValueError(f"expected bits and scale of shape [N,{self._in_planes}]")
        n = b.shape[0]
        if off.shape != (n + 1,) or off[0] != 0 or int(off[-1]) != idx.shape[0]:
            raise ValueError("move_offset must have N+1 entries from 0 to len(move_index)")
        priors = np.empty(idx.shape[0], dtype=np.float32)
        value = np.empty((n, 1), dtype=np.float32)
        fn = (self._lib.e55_service_predict_packed_moves if (self._service_cap and n <= self._service_cap)
              else self._lib.e55_predict_packed_moves)
        self._check(fn(self._ctx, b.ctypes.data, sc.ctypes.data, n, off.ctypes.data, idx.ctypes.data,
                       priors.ctypes.data, value.ctypes.data))
        return priors, value

    def predict_masked(self, images, legal_mask):
        """Fused policy tail: softmax over legal moves (what minishogilib.MCTS.evaluate does with the
        logits, mcts.py:60,105-106).  ``legal_mask`` uint8/bool ``[N,1725]``."""
        x = self._as_batch(images)
        n = x.shape[0]
        m = np.ascontiguousarray(np.asarray(legal_mask).astype(np.uint8))
        if m.shape != (n, POLICY_SIZE):
            raise ValueError(f"legal_mask must have shape [{n},{POLICY_SIZE}]")
        probs = np.empty((n, POLICY_SIZE), dtype=np.float32)
        value = np.empty((n, 1), dtype=np.float32)
        self._check(self._lib.e55_predict_masked(self._ctx, x.ctypes.data, m.ctypes.data, n,
                                                 probs.ctypes.data, value.ctypes.data))
        return probs, value

    def predict_device(self, planes, policy, value):
        """Device-resident variant: torch CUDA tensors (float32, contiguous); asynchronous on the context
        stream -- call :meth:`sync` before reading the outputs from another stream."""
        n = int(planes.shape[0])
        self._check(self._lib.e55_predict_device(self._ctx, planes.data_ptr(), n, policy.data_ptr(), value.data_ptr()))

    def debug_activations(self, images, n_convs=None):
        """Activations after the first ``n_convs`` 3x3 convolutions (default: trunk output)."""
        x = self._as_batch(images)
        if n_convs is None:
            n_convs = 1 + 2 * self._blocks
        out = np.empty((x.shape[0], self._filters, 5, 5), dtype=np.float32)
        self._check(self._lib.e55_debug_activations(self._ctx, x.ctypes.data, x.shape[0], n_convs, out.ctypes.data))
        return out

    def sync(self):
        self._check(self._lib.e55_sync(self._ctx))

    def cuda_stream(self) -> int:
        s = C.c_void_p()
        self._check(self._lib.e55_stream(self._ctx, C.byref(s)))
        return int(s.value or 0)

    def set_kernel_timing(self, enable: bool):
        self._check(self._lib.e55_set_kernel_timing(self._ctx, 1 if enable else 0))

    def kernel_timing(self):
        """(total device ms of the dominant kernel, launches) since set_kernel_timing(True)."""
        ms, n = C.c_double(), C.c_int64()
        self._check(self._lib.e55_get_kernel_timing(self._ctx, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    def launch_count(self) -> int:
        n = C.c_int64()
        self._check(self._lib.e55_launch_count(self._ctx, C.byref(n)))
        return int(n.value)

    # ---- training (out of scope) ----------------------------------------------------------------
    def iter(self):
        return self._iterations

    def step(self, train_images, policy_labels, value_labels, assertion=False):
        raise NotImplementedError("training (network.py:148-199) is outside the accelerated path; "
                                  "train with the reference and hand the weights over with set_weights()")


def decode_gpu_records(buf, n_records, timestamp=0):
    """Finished games in the record layout of ``csrc/e55_gpu_selfplay.cuh`` (``finish_game`` / ``record_ply``: a four-word
    header ``[plies, winner, words, truncated]``, then per ply ``move | not-a-target << 15, sum_N, Q * 65535, k`` and ``k``
    ``(move, N)`` pairs) as :class:`erweitern_55_b200.gamerecord.GameRecord` objects -- what ``selfplay.run`` returns
    (selfplay.py:80-113)."""
    from .gamerecord import GameRecord
    sfen = packed_move_sfen
    out = []
    for r in range(n_records):
        row = buf[r]
        rec = GameRecord()
        rec.winner, used, truncated, w = int(row[1]), int(row[2]), int(row[3]), 4
        while w < 4 + used:
            mv, sum_n, q, k = int(row[w]) & 0x7FFF, int(row[w + 1]), int(row[w + 2]) / 65535.0, int(row[w + 3])
            target = not (int(row[w]) >> 15)       # fast searches under playout-cap oscillation are no targets (selfplay.py:90-91)
            visits = [(sfen(int(row[w + 4 + 2 * i])), int(row[w + 5 + 2 * i])) for i in range(k)]
            w += 4 + 2 * k
            rec.sfen_kif.append(sfen(mv))
            rec.mcts_result.append((sum_n, q, visits))
            if k > 0 and target:                   # (nor is a ply whose visit list did not fit the record)
                rec.learning_target_plys.append(rec.ply)
            rec.ply += 1
        rec.timestamp = timestamp
        # False if the record buffer was too small for this game (plies or visit lists are missing: flagged by the kernel)
        rec.complete = rec.ply == int(row[0]) and not truncated
        out.append(rec)
    return out


def packed_move_sfen(mv):
    """USI string of a move as the GPU-resident search packs it: from (31 = drop) | to << 5 | piece << 10 | promotes << 14."""
    frm, to, piece, promo = mv & 31, (mv >> 5) & 31, (mv >> 10) & 15, (mv >> 14) & 1
    sq = lambda q: "%d%s" % (5 - q % 5, "abcde"[q // 5])
    return " KGSBRP"[piece] + "*" + sq(to) if frm == 31 else sq(frm) + sq(to) + ("+" if promo else "")


def expected_weight_shapes(blocks=7, filters=128, in_planes=IN_PLANES):
    return [s for _, s in weight_spec(blocks, filters, in_planes)]
