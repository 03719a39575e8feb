#!/usr/bin/env python
"""Benchmark of the accelerated path: batched policy/value network evaluation (BASELINE.json metric
"policy+value evals/sec (batch 1024)").

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--batch 1024] [--precision bf16]

A step = one forward pass of the default 7x128 ResNet (network.py:75-146) over one batch of 1024
synthetic positions with random-init weights (configs[1] of BASELINE.json).  Multi-GPU runs are launched
with torchrun: one process per GPU, each evaluating its own batches (independent positions shard with no
data-path collective -> weak scaling); NCCL is used only for the barrier and the max/sum of the results.

The JSON line carries: `value` (device-resident inputs, CUDA-event time on the launching stream, max over
ranks), `e2e` (same metric through the reference-facing host-buffer call e55_predict / Network.predict with
pinned host inputs: H2D + kernels + D2H inside the timed region), `roofline` (dominant kernel vs the
measured bf16 peak), `cpu_baseline` (the CPU oracle port on the box's host cores, bounded sample).
`--impl reference` times the reference's CPU path stand-in (the oracle port; TensorFlow 1.15 cannot run
here) with all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_POSITION = 126_368_256          # dense-convention FLOPs of the default net


def flop_per_position(blocks, filters, in_planes=266):
    """Dense-convention FLOPs of one forward pass: 126 368 256 for 7x128, 1 240 685 056 for 20x256."""
    conv = lambda cin, cout, k: 2 * 25 * cout * cin * k * k
    return (conv(in_planes, filters, 3) + (2 * blocks + 1) * conv(filters, filters, 3) + conv(filters, 69, 1)
            + conv(filters, 1, 1) + 2 * 25 * 128 + 2 * 128)
METRIC = "policy+value evals/sec (batch 1024)"
UNIT = "evals/s"
L2_BYTES = 126 * 1024 * 1024


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return float(d["bf16_tflops_sustained"]), "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"
    except Exception:
        return 1590.0, "fallback 1.59 PFLOP/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons for one GPU while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for nme, val in zip(names, r[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        busy = [s for s in sm if s > 0.5 * max(sm)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def host_info():
    model = "unknown"
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except Exception:
        pass
    return os.cpu_count() or 1, model


def cpu_oracle_rate(batch, seconds, weights):
    """evals/s of the CPU oracle port (torch fp32, all host threads) on repeated batches for ~`seconds`."""
    import torch
    from oracle import network_oracle as O
    from erweitern_55_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    prep = O.prepare_torch(weights)
    x = torch.from_numpy(synth.synthetic_planes(batch, seed=1000))
    O.forward_prepared(prep, x)                       # warm-up
    reps, t0 = 0, time.perf_counter()
    while True:
        O.forward_prepared(prep, x)
        reps += 1
        dt = time.perf_counter() - t0
        if dt >= seconds or reps >= 200:
            break
    return batch * reps / dt, cores, reps, dt


def run_reference(args, rank, world):
    """Reference arm: the reference's own CPU forward is TF 1.15 (absent); its stand-in is the oracle port."""
    if rank != 0:
        return
    import torch
    from oracle import network_oracle as O
    from erweitern_55_b200 import synth, weights as W
    cores, model = host_info()
    torch.set_num_threads(cores)
    w = W.init_weights(blocks=args.blocks, filters=args.filters, seed=55)
    prep = O.prepare_torch(w)
    x = torch.from_numpy(synth.synthetic_planes(args.batch, seed=1000))
    for _ in range(max(1, min(args.warmup, 2))):
        O.forward_prepared(prep, x)
    steps = min(args.steps, 40)                       # bounded: ~0.2-0.4 s per 1024-position step
    t0 = time.perf_counter()
    for _ in range(steps):
        O.forward_prepared(prep, x)
    dt = time.perf_counter() - t0
    val = args.batch * steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"default 7x128 ResNet forward, batch {args.batch}, random-init weights (configs[1])",
                   "batch": args.batch, "host_cpu": model},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{steps} forward passes of batch {args.batch} (torch-CPU fp32 restatement of network.py; "
                                   "TensorFlow 1.15 is not installable here)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--selfplay-workers", type=int, default=1, help="0 = skip the self-play runs")
    ap.add_argument("--selfplay-moves", type=int, default=8, help="moves per worker thread in the synthetic self-play run")
    ap.add_argument("--gpu-games", type=int, default=16384, help="games in flight per GPU in the GPU-resident self-play run (0 = skip)")
    ap.add_argument("--gpu-plies", type=int, default=8, help="plies per game slot in the GPU-resident self-play run")
    ap.add_argument("--selfplay-total-moves", type=int, default=24000, help="plies played per GPU in the real self-play run")
    ap.add_argument("--blocks", type=int, default=7, help="residual blocks (20 with --filters 256 = BASELINE configs[4])")
    ap.add_argument("--filters", type=int, default=128, choices=[128, 256])
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 50)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from erweitern_55_b200 import sharding, synth, weights as W
    from erweitern_55_b200.network import Network

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback (use --impl reference for the CPU arm)")
    # libraries (NCCL's version banner) write to stdout; keep stdout for the ONE JSON line
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # several ranks on one box: give each its own contiguous block of host cores, so the game threads, the packing
    # pool and the service threads of one GPU neither migrate across sockets nor sit on another rank's cores
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
    my_cores = sorted(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else list(range(os.cpu_count() or 1))
    if local_world > 1 and len(my_cores) >= 2 * local_world:
        per = len(my_cores) // local_world
        my_cores = my_cores[local_rank * per:(local_rank + 1) * per]
        try:
            os.sched_setaffinity(0, my_cores)
        except OSError:
            pass
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    B = args.batch
    w = W.init_weights(blocks=args.blocks, filters=args.filters, seed=55)
    flops = flop_per_position(args.blocks, args.filters)
    net = Network(device=local_rank, blocks=args.blocks, filters=args.filters, precision=args.precision, weights=w, max_batch=B)
    stream = torch.cuda.ExternalStream(net.cuda_stream(), device=dev)

    # distinct input batches whose total size exceeds L2, used round-robin: every step reads HBM
    in_bytes = B * 266 * 25 * 4
    n_pool = max(2, -(-int(1.7 * L2_BYTES) // in_bytes))
    pool = [torch.from_numpy(synth.synthetic_planes(B, seed=1000 + 16 * rank + i)).to(dev) for i in range(min(n_pool, 4))]
    while len(pool) < n_pool:                          # cheap variety: roll the first batches along the batch axis
        pool.append(torch.roll(pool[len(pool) % 4], shifts=len(pool), dims=0).contiguous())
    policy = torch.empty((B, 1725), dtype=torch.float32, device=dev)
    value = torch.empty((B, 1), dtype=torch.float32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident throughput ("value") ---------------------------------------------------
    # Two forwards of batch 1024 are kept in flight on two contexts (= two streams), the way the evaluation service
    # feeds the GPU: a 1024-position launch occupies 128 of the 148 SMs, the next one starts on the other 20 and on
    # every SM the first one leaves.  Every step is still one whole batch-1024 forward; the one-in-flight number
    # (kernels strictly one after the other) is reported beside it and is what `roofline` is measured on.
    net_b = Network(device=local_rank, precision=args.precision, weights=w, max_batch=B, blocks=args.blocks, filters=args.filters)
    lanes = [(net, policy, value, stream),
             (net_b, torch.empty_like(policy), torch.empty_like(value), torch.cuda.ExternalStream(net_b.cuda_stream(), device=dev))]

    def run_steps(n_steps, in_flight):
        for i in range(n_steps):
            nn_, p_, v_, _ = lanes[i % in_flight]
            nn_.predict_device(pool[i % n_pool], p_, v_)

    def timed_steps(n_steps, in_flight):
        for nn_, _, _, _ in lanes:
            nn_.sync()
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(in_flight)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(in_flight)]
        for k in range(in_flight):
            starts[k].record(lanes[k][3])
        run_steps(n_steps, in_flight)
        for k in range(in_flight):
            ends[k].record(lanes[k][3])
        for nn_, _, _, _ in lanes:
            nn_.sync()
        return max(a.elapsed_time(b) for a in starts for b in ends) * 1e-3

    run_steps(args.warmup * 2, 2)
    barrier()
    launches0 = net.launch_count() + net_b.launch_count()
    with ClockSampler(local_rank) as clocks:
        t_wall0 = time.perf_counter()
        dev_s = timed_steps(args.steps, 2)
        t_wall = time.perf_counter() - t_wall0
        launches = net.launch_count() + net_b.launch_count() - launches0
        # a very short timed region ends before nvidia-smi's first sample: keep the same load running
        # (untimed) until the sampler has seen it
        if t_wall < 0.5:
            t_end = time.perf_counter() + 0.5
            while time.perf_counter() < t_end:
                run_steps(32, 2)
                net.sync(); net_b.sync()
    barrier()
    total_units, max_s = sharding.aggregate(B * args.steps, dev_s, device=dev)
    value_rate = total_units / max_s
    one_s = timed_steps(max(args.steps // 2, 1), 1)
    barrier()
    one_units, one_max = sharding.aggregate(B * max(args.steps // 2, 1), one_s, device=dev)
    value_one_in_flight = one_units / one_max
    net_b.close()
    lanes = lanes[:1]

    # ---- dominant kernel vs roofline (rank 0, CUDA events around the tower kernel) ----------------
    roofline = None
    if rank == 0 and args.precision == "bf16":
        net.set_kernel_timing(True)
        for i in range(min(args.steps, 100)):
            net.predict_device(pool[i % n_pool], policy, value)
        net.sync()
        k_ms, k_n = net.kernel_timing()
        net.set_kernel_timing(False)
        peak, peak_src = measured_peaks()
        achieved = B * flops / (k_ms / k_n * 1e-3) / 1e12
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "tower_kernel_dram_bytes.json")) as f:
                traffic = json.load(f).get(f"batch_{B}" if (args.blocks, args.filters) == (7, 128) else f"batch_{B}_{args.blocks}x{args.filters}")
        except Exception:
            pass
        roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "traffic": traffic, "kernel": "e55::tc::tower_kernel", "kernel_us": k_ms / k_n * 1e3,
                    "flop_per_launch": B * flops, "peak_source": peak_src}

    # ---- end-to-end through the host-buffer call (reference-facing Network.predict path) ----------
    e2e_steps = args.e2e_steps or min(args.steps, 50)
    host_pool = [torch.from_numpy(synth.synthetic_planes(B, seed=3000 + 16 * rank + i)).pin_memory() for i in range(3)]
    host_policy = torch.empty((B, 1725), dtype=torch.float32).pin_memory()
    host_value = torch.empty((B, 1), dtype=torch.float32).pin_memory()
    lib, ctx = net._lib, net._ctx

    def host_step(i):
        net._check(lib.e55_predict(ctx, host_pool[i % 3].data_ptr(), B, host_policy.data_ptr(), host_value.data_ptr()))

    def timed_host_steps():
        for i in range(12):      # warm-up; in the library's automatic mode these calls also try both pipelines
            host_step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            host_step(i)
        dt = time.perf_counter() - t0
        barrier()
        units, t_max = sharding.aggregate(B * e2e_steps, dt, device=dev)
        return units / t_max

    # this rank's share of the host cores, minus two, packs the float planes into bits before the copy (e55_host.cpp)
    import ctypes as C
    net._check(lib.e55_set_host_threads(ctx, 0))
    e2e_plain = timed_host_steps()
    net._check(lib.e55_set_host_threads(ctx, -1))      # the default: the library picks the thread count and the pipeline
    e2e_value = timed_host_steps()
    pk_threads, pk_us, fl_us = C.c_int(), C.c_double(), C.c_double()
    net._check(lib.e55_get_host_pipeline(ctx, C.byref(pk_threads), C.byref(pk_us), C.byref(fl_us)))
    split_us = (C.c_double * 6)()
    eighths = C.c_int()
    net._check(lib.e55_get_host_split(ctx, C.byref(eighths), split_us))
    shares = (0, 1, 2, 3, 4, 8)
    best_share = min((u, sh) for u, sh in zip(split_us, shares) if u > 0)[1] if any(u > 0 for u in split_us) else 8
    n_float = B if best_share == 8 else (B // 8 * best_share) // 64 * 64
    pack_threads, compacting = pk_threads.value, best_share < 8
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": (B - n_float) * 266 * 8 + n_float * 266 * 25 * 4,
           "host_pipeline": "float" if best_share == 8 else "compacting" if best_share == 0 else f"compacting, {best_share}/8 of the batch as floats",
           "us_per_position": {f"{sh}/8 as floats": u for sh, u in zip(shares, split_us)},
           "d2h_bytes_per_step": B * 1725 * 4 + B * 4, "steps": e2e_steps, "host_input_bytes_per_step": in_bytes,
           "host_pack_threads": pack_threads, "value_without_host_packing": e2e_plain,
           "timing": "host wall clock around blocking e55_predict calls on pinned float32 [B,266,5,5] planes, max over ranks; "
                     "in its default mode the call bit-packs the planes on a host thread pool chunk by chunk (exact, results identical; "
                     "2 128 B per position cross PCIe instead of 26 600) and sends a share of the batch as float planes meanwhile "
                     "(link and cores both busy); it measures the shares 0..4/8 and 8/8 in its first calls and keeps the fastest "
                     "(host_pipeline, us_per_position; h2d_bytes_per_step is counted for that share); "
                     "value_without_host_packing is the same call with e55_set_host_threads(0)"}

    # ---- same call with the compact input format (e55_predict_packed; not the reference's format) ---
    from erweitern_55_b200 import packing
    packed = [packing.pack_planes(hp.numpy()) for hp in host_pool]
    pk_bits = [torch.from_numpy(b.view(np.int32)).pin_memory() for b, _ in packed]
    pk_scale = [torch.from_numpy(sc).pin_memory() for _, sc in packed]

    def packed_step(i):
        net._check(lib.e55_predict_packed(ctx, pk_bits[i % 3].data_ptr(), pk_scale[i % 3].data_ptr(), B,
                                          host_policy.data_ptr(), host_value.data_ptr()))

    for i in range(3):
        packed_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        packed_step(i)
    pk_s = time.perf_counter() - t0
    barrier()
    pk_units, pk_max = sharding.aggregate(B * e2e_steps, pk_s, device=dev)
    e2e_packed = {"value": pk_units / pk_max, "unit": UNIT, "h2d_bytes_per_step": B * 266 * 8,
                  "d2h_bytes_per_step": B * 1725 * 4 + B * 4, "steps": e2e_steps,
                  "note": "e55_predict_packed: bit-packed planes (8 B/plane) instead of the reference's float32 planes; "
                          "identical outputs; not the drop-in input format, reported beside e2e"}

    # ---- synthetic self-play load through the cross-game aggregation service (tree ops excluded) -------
    selfplay = selfplay_real = None
    if args.selfplay_workers > 0 and args.precision == "bf16":
        # game threads = this rank's cores minus two (minus one on small hosts) for the service's threads; 12 games per thread and a
        # service batch of ~74 positions per thread keep two batches' worth of leaves in flight (tools/selfplay_real_sweep.py)
        sp_threads = max(1, len(my_cores) - (2 if len(my_cores) > 12 else 1))
        sp_cap = min(B, max(256, 64 * round(74 * sp_threads / 64)))
        net.start_service(max_batch=sp_cap, max_wait_us=200)
        net.bench_selfplay(workers=sp_threads, moves=1, simulations=160, leaf_batch=16)   # warm-up
        barrier()
        mv, ev, mb = net.bench_selfplay(workers=sp_threads, moves=args.selfplay_moves * 40, simulations=800, leaf_batch=16)
        barrier()
        # real self-play: native minishogi engine + MCTS (800 simulations, leaf batch 16, 7-ply mate search)
        # one worker thread per host core of this rank's share, minus two for the service's dispatcher / completer
        net.selfplay(threads=sp_threads, total_moves=512, simulations=160, leaf_batch=16, mate_depth=7, seed=99 + rank)   # warm-up
        barrier()
        real = net.selfplay(threads=sp_threads, total_moves=args.selfplay_total_moves, simulations=800, leaf_batch=16,
                            mate_depth=7, seed=1 + rank, games_per_thread=12)
        barrier()
        net.stop_service()
        real_mv, _ = sharding.aggregate(real["moves_per_s"], 1.0, device=dev)
        real_ev, _ = sharding.aggregate(real["evals_per_s"], 1.0, device=dev)
        selfplay_real = {"moves_per_s": real_mv, "evals_per_s": real_ev, "games_rank0": real["games"],
                         "mean_gpu_batch_rank0": real["mean_gpu_batch"], "threads_per_gpu": sp_threads, "games_in_flight_per_gpu": sp_threads * 12, "service_batch_cap": sp_cap,
                         "mean_game_plies_rank0": real["mean_game_plies"],
                         "simulations_per_move": 800, "leaf_batch": 16, "mate_search_plies": 7,
                         "note": "native C++ minishogi engine + batched PUCT search (csrc/minishogi.hpp, mcts.hpp) playing whole "
                                 "games from the start position with random-init weights (client.py's settings: tree reuse, "
                                 "Dirichlet noise) on the asynchronous evaluation service with GPU-side legal-move softmax "
                                 "(e55_service_submit_packed_moves); rules pinned by the reference's 17 KATs, input encoding / "
                                 "policy index are this repo's own"}
        tot_ev, _ = sharding.aggregate(ev, 1.0, device=dev)
        tot_mv, _ = sharding.aggregate(mv, 1.0, device=dev)
        selfplay = {"moves_per_s": tot_mv, "evals_per_s": tot_ev, "mean_gpu_batch_rank0": mb,
                    "workers_per_gpu": sp_threads, "simulations_per_move": 800, "leaf_batch": 16,
                    "note": "synthetic: what the evaluation service itself sustains -- host threads keep 16-position requests "
                            "(bit-packed planes + 30 legal moves each, the shape mcts.MCTS.run produces, mcts.py:91-106) in flight on "
                            "asynchronous tickets; no tree work, no move generation"}

    # ---- GPU-resident self-play: the search itself on the GPU, no host game threads ---------------------------
    selfplay_gpu = None
    if args.selfplay_workers > 0 and args.precision == "bf16" and (args.blocks, args.filters) == (7, 128) and args.gpu_games > 0:
        # (this section is an extra beside the contract's numbers: a failure here is reported in the line, not fatal)
        gs, gpu_error = {"moves_per_s": 0.0, "evals_per_s": 0.0, "games": 0, "mean_game_plies": 0.0}, None
        try:
            net.gpu_selfplay(games=args.gpu_games, total_moves=args.gpu_games, simulations=800, seed=77 + rank)   # warm-up
        except Exception as exc:
            gpu_error = str(exc)
        barrier()
        if gpu_error is None:
            try:
                gs, _ = net.gpu_selfplay(games=args.gpu_games, total_moves=args.gpu_games * args.gpu_plies, simulations=800,
                                         mate_depth=7, seed=5 + rank)
            except Exception as exc:
                gpu_error = str(exc)
        barrier()
        g_mv, _ = sharding.aggregate(gs["moves_per_s"], 1.0, device=dev)
        g_ev, _ = sharding.aggregate(gs["evals_per_s"], 1.0, device=dev)
        selfplay_gpu = {"moves_per_s": g_mv, "evals_per_s": g_ev, "games_in_flight_per_gpu": args.gpu_games,
                        "plies_per_game_slot": args.gpu_plies, "games_finished_rank0": gs["games"],
                        "mean_game_plies_rank0": gs["mean_game_plies"], "simulations_per_move": 800, "mate_search_plies": 7,
                        "host_game_threads": 0, "error_rank0": gpu_error,
                        "note": "the same self-play with the search on the GPU (csrc/e55_gpu_selfplay.cuh): one thread per game does "
                                "the PUCT descent with virtual loss, move generation, repetition test, input encoding, expansion, "
                                "back-up and move choice between the network launches (16 slots per game per launch); tree reuse "
                                "and Dirichlet noise as in client.py; the host only answers the 7-ply mate searches.  All games start "
                                "together at ply 0, so a short run sees mostly openings (cheaper positions, few mates)"}

    # ---- CPU baseline (rank 0, N=1 only) ----------------------------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1:
        rate, cores, reps, dt = cpu_oracle_rate(B, args.cpu_seconds, w)
        cpu_baseline = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{reps} forward passes of batch {B} in {dt:.1f} s (torch-CPU fp32 oracle port of network.py)"}
        # BASELINE.json configs[0]: the batches mcts.py / usi.py really use (root 1, leaves 16), CPU vs this library
        small = {}
        for nb in (1, 16, 64):
            r_cpu, _, _, _ = cpu_oracle_rate(nb, 1.0, w)
            xs = synth.synthetic_planes(nb, seed=77)
            net.predict(xs)
            t0 = time.perf_counter()
            for _ in range(200):
                net.predict(xs)
            small[str(nb)] = {"cpu_port_evals_per_s": r_cpu, "b200_predict_evals_per_s": nb * 200 / (time.perf_counter() - t0)}
        cpu_baseline["small_batches"] = small

    if rank == 0:
        ncpu, model = host_info()
        line = {
            "metric": METRIC, "value": value_rate, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": max_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": (f"default 7x128 ResNet forward (stem + 7 residual blocks + policy/value heads), batch {B} "
                                    "per GPU, random-init weights (BASELINE.json configs[1])") if (args.blocks, args.filters) == (7, 128)
                       else f"scaled tower {args.blocks}x{args.filters} forward, batch {B} per GPU, random-init weights (BASELINE.json configs[4])",
                       "batch_per_gpu": B, "parallelism": f"replicas x{world} (independent positions, no collective)",
                       "in_flight": "two batch-%d forwards in flight per GPU on two streams (as the evaluation service runs them); "
                                    "value_one_in_flight = kernels strictly one after the other" % B,
                       "l2": f"inputs rotate over {n_pool} distinct device batches = {n_pool * in_bytes >> 20} MiB > 126 MiB L2",
                       "flop_per_position": flops, "host_cpu": model, "host_cores_per_rank": len(my_cores), "host_cores": ncpu},
            "value_one_in_flight": value_one_in_flight, "e2e": e2e, "e2e_packed": e2e_packed, "selfplay": selfplay_real, "selfplay_gpu": selfplay_gpu, "selfplay_synthetic": selfplay, "gpu_launches": launches, "clocks": clocks.summary(),
            "roofline": roofline, "cpu_baseline": cpu_baseline, "wall_s": t_wall,
        }
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    net.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
