#!/usr/bin/env python
"""Benchmark of the accelerated path: batched policy/value network evaluation (BASELINE.json metric
"policy+value evals/sec (batch 1024) and self-play moves/sec at 1/2/4/8 B200").

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--batch 1024] [--precision bf16]

A step = one forward pass of the default 7x128 ResNet (network.py:75-146) over one batch of 1024 synthetic positions
with random-init weights (configs[1] of BASELINE.json).  Multi-GPU runs are launched with torchrun: one process per
GPU, each evaluating its own batches (independent positions shard with no data-path collective -> weak scaling); NCCL
is used only for the barrier and the max/sum of the results.

What the ONE JSON line holds is described in DESIGN.md section 5 (the line itself carries numbers, not prose):
`value` (device-resident inputs, CUDA events on the launching streams, K steps per timed region, regions repeated
for >= 0.6 s, median region, max over ranks), `e2e` (the drop-in call Network.predict(numpy) -> numpy on ordinary
pageable numpy arrays, copies inside the timed region), `roofline`, `cpu_baseline`, and the other configs of BASELINE.json
(`selfplay*`: configs[3], `scaled_20x256`: configs[4], `e2e.small_batches`: configs[0]/[2]).
`--impl reference` times the reference's CPU path stand-in (the oracle port; TensorFlow 1.15 cannot run here).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import erweitern_55_b200  # noqa: E402,F401  (sets CUDA_DEVICE_MAX_CONNECTIONS before torch creates the CUDA context)

METRIC = "policy+value evals/sec (batch 1024)"
UNIT = "evals/s"
L2_BYTES = 126 * 1024 * 1024
MIN_TIMED_S = 0.6            # every reported rate is the median of K-step regions repeated for at least this long


def flop_per_position(blocks, filters, in_planes=266):
    """Dense-convention FLOPs of one forward pass: 126 368 256 for 7x128, 1 240 685 056 for 20x256."""
    conv = lambda cin, cout, k: 2 * 25 * cout * cin * k * k
    return (conv(in_planes, filters, 3) + (2 * blocks + 1) * conv(filters, filters, 3) + conv(filters, 69, 1)
            + conv(filters, 1, 1) + 2 * 25 * 128 + 2 * 128)


def workload(blocks, filters, batch):
    """config.workload, the same string in both arms."""
    if (blocks, filters) == (7, 128):
        return f"default 7x128 ResNet forward, batch {batch} per GPU, random-init weights (BASELINE.json configs[1])"
    return f"scaled tower {blocks}x{filters} forward, batch {batch} per GPU, random-init weights (BASELINE.json configs[4])"


def measured_peaks():
    """(sustained, burst, source) dense bf16 TFLOP/s."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        return float(d["bf16_tflops_sustained"]), float(d["bf16_tflops"]), "MEASURED_PEAKS.json"
    except Exception:
        return 1590.0, 1590.0, "fallback 1.59 PFLOP/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons for one GPU while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
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
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


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
    steps = min(args.steps, 40)                       # bounded: ~0.1 s per 1024-position step on 16 cores
    t0 = time.perf_counter()
    for _ in range(steps):
        O.forward_prepared(prep, x)
    dt = time.perf_counter() - t0
    val = args.batch * steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload(args.blocks, args.filters, args.batch), "batch_per_gpu": args.batch, "host_cpu": model},
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
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--selfplay-workers", type=int, default=1, help="0 = skip the self-play runs")
    ap.add_argument("--selfplay-moves", type=int, default=8, help="moves per worker thread in the synthetic self-play run")
    ap.add_argument("--gpu-games", type=int, default=16384, help="games in flight per GPU in the GPU-resident self-play run (0 = skip)")
    ap.add_argument("--gpu-plies", type=int, default=24, help="plies per game slot in the GPU-resident self-play run")
    ap.add_argument("--selfplay-total-moves", type=int, default=24000, help="plies played per GPU in the real self-play run")
    ap.add_argument("--blocks", type=int, default=7, help="residual blocks (20 with --filters 256 = BASELINE configs[4])")
    ap.add_argument("--filters", type=int, default=128, choices=[128, 256])
    ap.add_argument("--scaled", type=int, default=1, help="also time the scaled 20x256 tower (configs[4]); 0 = skip")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import ctypes as C
    import torch
    import torch.distributed as dist
    from erweitern_55_b200 import hostmem, packing, sharding, synth, weights as W
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

    B, K = args.batch, args.steps
    default_net = (args.blocks, args.filters) == (7, 128)
    w = W.init_weights(blocks=args.blocks, filters=args.filters, seed=55)
    flops = flop_per_position(args.blocks, args.filters)
    net = Network(device=local_rank, blocks=args.blocks, filters=args.filters, precision=args.precision, weights=w, max_batch=B)
    lib, ctx = net._lib, net._ctx
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

    host_exchange = sharding.exchange                  # all ranks' payloads through the rendezvous store (host side only)

    def repeat_regions(region, est_s):
        """Runs `region()` (-> seconds of one K-step region) until MIN_TIMED_S have been timed; median, count, total."""
        reps = max(3, min(2000, int(math.ceil(MIN_TIMED_S / max(est_s, 1e-6)))))
        ts = [region() for _ in range(reps)]
        return statistics.median(ts), reps, sum(ts)

    # ---- device-resident throughput ("value") ---------------------------------------------------
    # Two forwards of batch 1024 are kept in flight on two contexts (= two streams), the way the evaluation service
    # feeds the GPU: a 1024-position launch occupies 128 of the 148 SMs, the next one starts on the other 20 and on
    # every SM the first one leaves.  value_one_in_flight: kernels strictly one after the other.
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
    est = timed_steps(K, 2)
    barrier()
    launches0 = net.launch_count() + net_b.launch_count()
    with ClockSampler(local_rank) as clocks:
        t_wall0 = time.perf_counter()
        dev_s, regions, timed_total = repeat_regions(lambda: timed_steps(K, 2), est)
        t_wall = time.perf_counter() - t_wall0
    launches = net.launch_count() + net_b.launch_count() - launches0
    barrier()
    total_units, max_s = sharding.aggregate(B * K, dev_s, device=dev)
    value_rate = total_units / max_s
    one_s, _, _ = repeat_regions(lambda: timed_steps(K, 1), est)
    barrier()
    one_units, one_max = sharding.aggregate(B * K, one_s, device=dev)
    value_one_in_flight = one_units / one_max
    net_b.close()
    lanes = lanes[:1]

    # ---- dominant kernel vs roofline (rank 0, CUDA events around the tower kernel, one launch at a time) -------
    roofline = None
    if rank == 0 and args.precision == "bf16":
        net.set_kernel_timing(True)
        for i in range(max(100, min(K, 1000))):
            net.predict_device(pool[i % n_pool], policy, value)
        net.sync()
        k_ms, k_n = net.kernel_timing()
        net.set_kernel_timing(False)
        peak, burst, peak_src = measured_peaks()
        achieved = B * flops / (k_ms / k_n * 1e-3) / 1e12
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "tower_kernel_dram_bytes.json")) as f:
                traffic = json.load(f).get(f"batch_{B}" if default_net else f"batch_{B}_{args.blocks}x{args.filters}")
        except Exception:
            pass
        roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "traffic": traffic, "kernel": "e55::tc::tower_kernel", "kernel_us": k_ms / k_n * 1e3,
                    "flop_per_launch": B * flops, "peak_source": peak_src + " bf16_tflops_sustained",
                    "peak_burst": burst, "frac_of_burst": achieved / burst, "launches_timed": k_n}

    # ---- end to end: the drop-in call, Network.predict(numpy) -> numpy (network.py:201-216) ----------------------
    # Inputs: three numpy arrays in rotation (81 MB: the planes come from DRAM, not from the host's cache) -- ordinary
    # pageable ones for the headline, page-locked ones (as Network.zero_inputs hands them out) beside it; results: the
    # fresh numpy arrays predict returns.
    def pinned_copy(a):
        b = hostmem.empty(a.shape)
        b[...] = a
        return b

    host_src = [synth.synthetic_planes(B, seed=3000 + 16 * rank + i) for i in range(3)]
    host_pool = [pinned_copy(a) for a in host_src]

    def e2e_region(arrs, nb, steps):
        def region():
            t0 = time.perf_counter()
            for i in range(steps):
                net.predict(arrs[i % len(arrs)])
            return time.perf_counter() - t0
        return region

    def e2e_rate(arrs, nb, steps, warm=20):
        """evals/s of net.predict on `arrs` (whole job: sum over ranks / slowest rank), and seconds per call."""
        for i in range(warm):                      # in the library's automatic mode these calls also try every host split
            net.predict(arrs[i % len(arrs)])
        region = e2e_region(arrs, nb, steps)
        est_s = region()
        barrier()
        med, _, _ = repeat_regions(region, est_s)
        barrier()
        units, t_max = sharding.aggregate(nb * steps, med, device=dev)
        return units / t_max, med / steps

    net._check(lib.e55_set_host_threads(ctx, -1))      # the default: the library picks the thread count and the pipeline
    # the headline: ordinary (pageable) numpy arrays, what mcts.py hands to predict (mcts.py:101-102); page-locked inputs
    # (Network.zero_inputs) are measured beside it -- there the library may send a share of the batch through the copy engine
    e2e_value, e2e_call_s = e2e_rate(host_src, B, K)
    e2e_pinned, _ = e2e_rate(host_pool, B, K)
    pk_threads, pk_us, fl_us = C.c_int(), C.c_double(), C.c_double()
    net._check(lib.e55_get_host_pipeline(ctx, C.byref(pk_threads), C.byref(pk_us), C.byref(fl_us)))
    split_us = (C.c_double * 8)()
    eighths = C.c_int()
    net._check(lib.e55_get_host_split(ctx, C.byref(eighths), split_us))
    shares = (0, 1, 2, 3, 4, 5, 6, 8)
    best_share = min((u, sh) for u, sh in zip(split_us, shares) if u > 0)[1] if any(u > 0 for u in split_us) else 0
    n_float = 0                                        # pageable inputs are always packed whole (the cores read them either way)
    net._check(lib.e55_set_host_threads(ctx, 0))
    e2e_plain, _ = e2e_rate(host_pool, B, max(3, K // 4), warm=3)
    net._check(lib.e55_set_host_threads(ctx, -1))
    small = {}
    for nb in (1, 16):                                 # the reference's own batch sizes (mcts.py:57-58,101-102)
        xs = [synth.synthetic_planes(nb, seed=70 + i) for i in range(3)]
        rate, call_s = e2e_rate(xs, nb, 200)
        small[str(nb)] = {"us_per_call": call_s * 1e6, "evals_per_s": rate}
    # ---- batch sweep (BASELINE.json configs[2]) on this rank: device-resident us per launch, Network.predict us per call ----
    batch_sweep = None
    if default_net and args.precision == "bf16":
        try:
            big = torch.cat(pool[:4])[:4096] if B * min(len(pool), 4) >= 4096 else None
            sweep_p = torch.empty((4096, 1725), dtype=torch.float32, device=dev)
            sweep_v = torch.empty((4096, 1), dtype=torch.float32, device=dev)
            host_big = np.concatenate(host_src + [host_src[0]])[:4096] if 4 * B >= 4096 else None
            batch_sweep = {}
            for nb in (1, 16, 64, 256, 512, 1024, 1184, 2048, 4096):
                if big is None or host_big is None or nb > big.shape[0]:
                    continue
                xs_d = big[:nb]
                for _ in range(5):
                    net.predict_device(xs_d, sweep_p[:nb], sweep_v[:nb])
                net.sync()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 40 if nb <= 1184 else 12
                e0.record(stream)
                for _ in range(reps):
                    net.predict_device(xs_d, sweep_p[:nb], sweep_v[:nb])
                e1.record(stream)
                net.sync()
                dev_us = e0.elapsed_time(e1) * 1e3 / reps
                xs_h = host_big[:nb]
                for _ in range(3):
                    net.predict(xs_h)
                t0 = time.perf_counter()
                for _ in range(reps):
                    net.predict(xs_h)
                call_us = (time.perf_counter() - t0) * 1e6 / reps
                batch_sweep[str(nb)] = {"device_us": round(dev_us, 1), "predict_us": round(call_us, 1)}
            del big, sweep_p, sweep_v, host_big
        except Exception as exc:                       # an extra: never at the cost of the line
            batch_sweep = {"error": str(exc)[:200]}
    # the host-memory roofline of this number: all ranks at once read 256 MiB each with their pool threads (nothing else:
    # no packing, no copies); every position costs 26 600 bytes of such reads, by cores or copy engine
    gbs = C.c_double()
    probe = np.ones(256 << 20, dtype=np.uint8)         # 256 MiB per rank, written first: far beyond the host's caches
    barrier()
    net._check(lib.e55_host_read_bandwidth(ctx, probe.ctypes.data, probe.nbytes, 3, C.byref(gbs)))
    barrier()
    del probe
    host_gbs, _ = sharding.aggregate(gbs.value, 1.0, device=dev)
    e2e = {"value": e2e_value, "unit": UNIT, "host_read_gbs": host_gbs, "host_read_bound_evals_per_s": host_gbs * 1e9 / (266 * 25 * 4), "h2d_bytes_per_step": (B - n_float) * 266 * 8 + n_float * 266 * 25 * 4,
           "d2h_bytes_per_step": B * 1725 * 4 + B * 4, "call": "Network.predict(numpy)->numpy", "us_per_call": e2e_call_s * 1e6,
           "input_memory": "numpy arrays (pageable), 3 in rotation", "host_input_bytes_per_step": in_bytes,
           "host_pack_threads": pk_threads.value, "value_pinned_input": e2e_pinned, "pinned_input_float_eighths": best_share,
           "pinned_input_us_per_position_by_float_eighths": {str(sh): round(u, 4) for sh, u in zip(shares, split_us)},
           "value_without_host_packing": e2e_plain, "small_batches": small, "batch_sweep_rank0": batch_sweep}

    # ---- same call with the compact input format (e55_predict_packed; not the reference's format) ---
    packed = [packing.pack_planes(a) for a in host_src]
    pk_bits = [torch.from_numpy(b.view(np.int32)).pin_memory() for b, _ in packed]
    pk_scale = [torch.from_numpy(sc).pin_memory() for _, sc in packed]
    host_policy = torch.empty((B, 1725), dtype=torch.float32).pin_memory()
    host_value = torch.empty((B, 1), dtype=torch.float32).pin_memory()

    def packed_region():
        t0 = time.perf_counter()
        for i in range(K):
            net._check(lib.e55_predict_packed(ctx, pk_bits[i % 3].data_ptr(), pk_scale[i % 3].data_ptr(), B,
                                              host_policy.data_ptr(), host_value.data_ptr()))
        return time.perf_counter() - t0

    est = packed_region()
    barrier()
    pk_s, _, _ = repeat_regions(packed_region, est)
    barrier()
    pk_units, pk_max = sharding.aggregate(B * K, pk_s, device=dev)
    e2e_packed = {"value": pk_units / pk_max, "unit": UNIT, "h2d_bytes_per_step": B * 266 * 8,
                  "d2h_bytes_per_step": B * 1725 * 4 + B * 4, "call": "e55_predict_packed"}

    # ---- scaled tower 20x256 (BASELINE.json configs[4]) at the full-machine batch, device-resident -----------------
    scaled = None
    if args.scaled and default_net and args.precision == "bf16":
        sb, sblocks, sfilters = 1184, 20, 256
        snet = Network(device=local_rank, blocks=sblocks, filters=sfilters, precision="bf16",
                       weights=W.init_weights(blocks=sblocks, filters=sfilters, seed=55), max_batch=sb)
        sx = [torch.from_numpy(synth.synthetic_planes(sb, seed=500 + rank + i)).to(dev) for i in range(5)]   # 157 MB > L2
        sp, sv = torch.empty((sb, 1725), device=dev), torch.empty((sb, 1), device=dev)
        sstream = torch.cuda.ExternalStream(snet.cuda_stream(), device=dev)
        ssteps = 20

        def scaled_region():
            snet.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(sstream)
            for i in range(ssteps):
                snet.predict_device(sx[i % 5], sp, sv)
            e1.record(sstream)
            snet.sync()
            return e0.elapsed_time(e1) * 1e-3

        est = scaled_region()
        barrier()
        s_s, _, _ = repeat_regions(scaled_region, est)
        barrier()
        s_units, s_max = sharding.aggregate(sb * ssteps, s_s, device=dev)
        sflops = flop_per_position(sblocks, sfilters)
        peak, burst, _ = measured_peaks()
        per_gpu_tflops = sb * ssteps * sflops / s_s / 1e12
        scaled = {"workload": workload(sblocks, sfilters, sb), "value": s_units / s_max, "unit": UNIT, "ms_per_step": s_max / ssteps * 1e3,
                  "flop_per_position": sflops, "tflops_rank0": per_gpu_tflops, "frac_rank0": per_gpu_tflops / peak,
                  "frac_of_burst_rank0": per_gpu_tflops / burst}
        snet.close()
        del sx, sp, sv

    # ---- self-play (BASELINE.json configs[3]) ------------------------------------------------------------------------
    selfplay = selfplay_real = None
    if args.selfplay_workers > 0 and args.precision == "bf16" and default_net:
        # host driver: game threads = this rank's cores minus two (minus one on small hosts) for the service's threads; 12 games per
        # thread and a service batch of ~74 positions per thread keep two batches' worth of leaves in flight
        sp_threads = max(1, len(my_cores) - (2 if len(my_cores) > 12 else 1))
        sp_cap = min(1184, max(256, 64 * round(74 * sp_threads / 64)))
        net.start_service(max_batch=sp_cap, max_wait_us=200)
        net.bench_selfplay(workers=sp_threads, moves=1, simulations=160, leaf_batch=16)   # warm-up
        barrier()
        mv, ev, mb = net.bench_selfplay(workers=sp_threads, moves=args.selfplay_moves * 40, simulations=800, leaf_batch=16)
        barrier()
        net.selfplay(threads=sp_threads, total_moves=512, simulations=160, leaf_batch=16, mate_depth=7, seed=99 + rank)   # warm-up
        barrier()
        real = net.selfplay(threads=sp_threads, total_moves=args.selfplay_total_moves, simulations=800, leaf_batch=16,
                            mate_depth=7, seed=1 + rank, games_per_thread=12)
        barrier()
        net.stop_service()
        real_mv, _ = sharding.aggregate(real["moves_per_s"], 1.0, device=dev)
        real_ev, _ = sharding.aggregate(real["evals_per_s"], 1.0, device=dev)
        selfplay_real = {"moves_per_s": real_mv, "evals_per_s": real_ev, "driver": "host threads + evaluation service",
                         "threads_per_gpu": sp_threads, "games_in_flight_per_gpu": sp_threads * 12, "service_batch_cap": sp_cap,
                         "mean_gpu_batch_rank0": real["mean_gpu_batch"], "mean_game_plies_rank0": real["mean_game_plies"],
                         "games_rank0": real["games"], "simulations_per_move": 800, "leaf_batch": 16, "mate_search_plies": 7}
        tot_ev, _ = sharding.aggregate(ev, 1.0, device=dev)
        tot_mv, _ = sharding.aggregate(mv, 1.0, device=dev)
        selfplay = {"moves_per_s": tot_mv, "evals_per_s": tot_ev, "mean_gpu_batch_rank0": mb, "workers_per_gpu": sp_threads}

    # ---- GPU-resident self-play: the last device work of the run, and none of it behind a device collective ------------
    # (the ranks meet through the rendezvous store, on the host: a search kernel that ever hung would hang a NCCL
    # barrier on every rank with it, and with the barrier the whole line; this way the line still carries everything
    # measured above, and says what went wrong here)
    selfplay_gpu = None
    device_hung = False
    if args.selfplay_workers > 0 and args.precision == "bf16" and default_net and args.gpu_games > 0:
        gs, gpu_error = {"moves_per_s": 0.0, "evals_per_s": 0.0, "games": 0, "mean_game_plies": 0.0, "moves": 0, "evals": 0, "seconds": 0.0}, None
        try:
            net.gpu_selfplay(games=args.gpu_games, total_moves=args.gpu_games, simulations=800, seed=77 + rank)   # warm-up
        except Exception as exc:
            gpu_error = str(exc)
        host_exchange("gpu_selfplay_warm", {"error": gpu_error})
        if gpu_error is None:
            try:
                gs, _ = net.gpu_selfplay(games=args.gpu_games, total_moves=args.gpu_games * args.gpu_plies, simulations=800,
                                         mate_depth=7, seed=5 + rank)
            except Exception as exc:
                gpu_error = str(exc)
        got = host_exchange("gpu_selfplay", {"moves": gs.get("moves", 0), "evals": gs.get("evals", 0), "seconds": gs.get("seconds", 0.0),
                                             "error": gpu_error})
        device_hung = any(x["error"] for x in got)       # any failure here: leave without touching the devices again
        if rank == 0:
            worst_s = max(float(x["seconds"]) for x in got)
            errors = {str(i): x["error"] for i, x in enumerate(got) if x["error"]}
            ok = not errors and worst_s > 0
            selfplay_gpu = {"moves_per_s": sum(x["moves"] for x in got) / worst_s if ok else 0.0,
                            "evals_per_s": sum(x["evals"] for x in got) / worst_s if ok else 0.0,
                            "aggregate": "moves of all ranks / the slowest rank's seconds",
                            "driver": "GPU-resident search", "games_in_flight_per_gpu": args.gpu_games,
                            "plies_per_game_slot": args.gpu_plies, "games_finished_rank0": gs["games"],
                            "mean_game_plies_rank0": gs["mean_game_plies"], "simulations_per_move": 800, "mate_search_plies": 7,
                            "tree_pool_overflows_rank0": gs.get("tree_overflows"), "mate_budget_exhausted_rank0": gs.get("mate_exhausted"),
                            "errors_by_rank": errors or None}

    # ---- CPU baseline (rank 0, N=1 only) ----------------------------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1:
        rate, cores, reps, dt = cpu_oracle_rate(B, args.cpu_seconds, w)
        cpu_baseline = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{reps} forward passes of batch {B} in {dt:.1f} s (torch-CPU fp32 oracle port of network.py)"}
        # BASELINE.json configs[0]: the batches mcts.py / usi.py really use (root 1, leaves 16), CPU port
        cpu_baseline["small_batches"] = {str(nb): {"cpu_port_evals_per_s": cpu_oracle_rate(nb, 1.0, w)[0],
                                                   "b200_predict_evals_per_s": small.get(str(nb), {}).get("evals_per_s")} for nb in (1, 16)}

    if rank == 0:
        ncpu, model = host_info()
        line = {
            "metric": METRIC, "value": value_rate, "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": args.warmup, "ms_per_step": max_s / K * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": workload(args.blocks, args.filters, B), "batch_per_gpu": B,
                       "parallelism": f"replicas x{world}, no collective", "in_flight": 2,
                       "l2": f"inputs rotate over {n_pool} device batches = {n_pool * in_bytes >> 20} MiB > 126 MiB L2",
                       "flop_per_position": flops, "host_cpu": model, "host_cores_per_rank": len(my_cores), "host_cores": ncpu},
            "timed_regions": regions, "timed_s": timed_total, "wall_s": t_wall, "gpu_launches": launches, "clocks": clocks.summary(),
            "value_one_in_flight": value_one_in_flight, "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e,
            "e2e_packed": e2e_packed, "scaled_20x256": scaled, "selfplay_synthetic": selfplay, "selfplay": selfplay_real,
            "selfplay_gpu": selfplay_gpu,
        }
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if device_hung:
        sys.stderr.write("[bench] GPU self-play failed; leaving without device teardown\n")
        sys.stderr.flush()
        os._exit(0)
    net.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
