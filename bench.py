#!/usr/bin/env python
"""Benchmark of the per-timestep vehicle update (the loop of test/profile_cars_update.py:28-34) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4_city_1m] [--impl reference]

One "step" = one simulation timestep over every car of the workload (Cars.update + TrafficLights.update, in the
profile script's order).  Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definitions:
  value     vehicle-updates/s with the state resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e       the same metric through the host-buffer C-ABI call tb2_step_host_async: per frame interval, H2D of the
            light-face states from pinned memory, k steps, D2H of the position frame + face states (north_star: frames
            every k steps); at least 8 frames whatever --steps is
  roofline  algorithmic bytes per car-step x cars / mean step-kernel duration vs the measured HBM peak
  cpu_baseline  the C oracle port of the reference, one thread, on a bounded sample of the same workload
  small_world   BASELINE configs[1] (2k cars): microseconds per step of the cluster-resident kernel and what bounds it
  ensemble      BASELINE configs[4]: --replicas rollouts of the 2k-car world with randomised light timings, sharded over
                the ranks, run to arrival or --ensemble-max-steps, trip times + rewards gathered over NCCL (strong scaling)
  facade        BASELINE configs[0] (33 cars) through the drop-in Cars / TrafficLights classes
With --gpus N (torchrun, one rank per GPU) every rank advances an independent replica of the headline workload (the path
shards only by replica).  --impl reference times the reference's CPU algorithm (the oracle port, all host
threads) on the same workload: N replicas one after the other when N > 1, so that the two arms do the same work.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

# stdout carries exactly ONE JSON line: libraries that write to fd 1 (NCCL prints its version there) go to stderr
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line):
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()

# algorithmic bytes per car-step of the fused fp64 step kernel (DESIGN.md "Kernels", table step_f64), for the launches
# of a tb2_step batch that do not write the observation-only columns (all but the last of the batch):
#   reads : pos 16, route_time 8, cursor 4, path_end 4, path_eff_end 4, cell 4, cached head point 16,
#           cached curvature constants 16, cell-table entry 8, candidate position 16            = 96
#   writes: pos 16, route_time 8, two 4-byte atomics into the next cell table 8                 = 32
# (lights / faces / cell->lights CSR are O(L) << N and L2-resident: counted as 0)
# fp32 throughput mode: the same fields at half the width for reals: reads 8+4+4+4+4+4+8+8+8+8 = 60, writes 8+4+8 = 20
ALGO_BYTES = {"f64": 128.0, "f32": 80.0}
L2_BYTES = 126e6


def peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons, power = [], [], set(), []
        for ts, ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            inside = t0 - 0.05 <= ts <= t1 + 0.15
            try:
                if inside:
                    sm.append(float(f[1])); power.append(float(f[3]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            if inside:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        if not sm:  # region shorter than the sampling period: use every sample we have
            for ts, ln in self.lines:
                f = [x.strip() for x in ln.split(",")]
                try:
                    sm.append(float(f[1]))
                except Exception:
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w": float(np.median(power)) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def config_of(args, w, n_ranks, frame_k=None):
    """The `config` object of the JSON line: identical keys (and values) in the native and the reference arm."""
    ws = w.n_cars * 136 + min(len(w.path_xy) * 16, w.n_cars * 64)
    return {"workload": args.workload, "cars_per_gpu": int(w.n_cars), "lights": int(w.n_lights), "faces": int(w.n_faces),
            "path_points": int(len(w.path_xy)), "cells": int(w.grid.n_cells), "dt": args.dt, "order": "cars_then_lights",
            "replicas": int(n_ranks), "parallelism": "replica-per-gpu", "neighbor": args.neighbor,
            "l2": ("working set %.0f MB per step > %.0f MB L2, no explicit flush" if ws > L2_BYTES else
                   "working set %.0f MB per step FITS the %.0f MB L2 and is not flushed (state is resident by design)")
                  % (ws / 1e6, L2_BYTES / 1e6),
            "frame_every_k_steps": int(args.frame_every)}


def make_workload(name, seed):
    from traffic_b200 import synth
    cfg = dict(synth.CITY_CONFIGS[name])
    cfg["seed"] = seed
    if os.environ.get("TB2_SPATIAL_ORDER") == "1":   # experiment: cars numbered in bin order at t=0
        cfg["spatial_order"] = True
    w, _ = synth.make_city(**cfg)
    return w


def cpu_port_updates_per_s(w, dt, order, budget_s, threads):
    """Time the C oracle (the CPU port of the reference algorithm) on a bounded sample of the workload."""
    from oracle import oracle
    wc = w.copy()
    oracle.run(wc, dt, 1, order, threads=threads)  # warm caches / page in
    steps, t_used = 0, 0.0
    chunk = 2
    while t_used < budget_s and steps < 4000:
        t0 = time.perf_counter()
        oracle.run(wc, dt, chunk, order, threads=threads)
        t_used += time.perf_counter() - t0
        steps += chunk
        chunk = min(chunk * 2, 256)
    return w.n_cars * steps / t_used, steps, t_used


def facade_profile_loop():
    """BASELINE configs[0]: the loop of test/profile_cars_update.py:28-34 (33 cars, 1,000 steps, dt=0.1) driven through the
    drop-in Cars / TrafficLights classes and initialisers, i.e. the reference-facing API end to end, state read back at
    the end.  (Its parity is tests/test_facade.py's business; here it is timed.)"""
    import random
    import torch
    from traffic_b200 import init, synth
    from traffic_b200.facade import Cars, TrafficLights
    graph = synth.SynthGraph(19, 19, seed=0)                       # Piedmont-sized grid, SURVEY 8(d) c1
    random.seed(0)
    cars_df = init.init_random_node_start_location(34, graph)     # -> 33 cars (simulation.py:205-206)
    lights_df = init.init_traffic_lights(graph, prescale=40)      # profile_cars_update.py's value
    cars_object = Cars(init_state=cars_df, graph=graph)
    lights_object = TrafficLights(light_state=lights_df, graph=graph)
    T = 1000
    cars_object.update(dt=0.1, lights=lights_object.state); lights_object.update(dt=0.1)   # binds + warms up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(T - 1):
        cars_object.update(dt=0.1, lights=lights_object.state)
        lights_object.update(dt=0.1)
    x = cars_object.state["x"].to_numpy()     # materialises one column of the DataFrame from the device
    el = time.perf_counter() - t0
    return {"workload": "c1 (33 cars, profile_cars_update.py loop) through the drop-in Cars/TrafficLights classes",
            "cars": int(len(x)), "steps": T - 1, "value": len(x) * (T - 1) / el, "unit": "vehicle-updates/s",
            "us_per_step": el / (T - 1) * 1e6, "finite": bool(np.isfinite(x).all()),
            "reference_python_updates_per_s": "300-450 (tests/golden/*.npz ref_updates_per_s)"}


def small_world_block(args, prec):
    """BASELINE configs[1]: the 2k-car Manhattan-sized world, one simulation on one GPU.  The world lives on chip (one
    thread-block cluster, DSMEM tables, a cluster barrier per step): it is bound by the dependent-issue chain of one
    car-step plus the barrier, not by memory - report the step latency and the instruction-issue view of it."""
    import torch
    from traffic_b200 import native
    w2 = make_workload("c2_manhattan_2k", args.seed)
    s2 = native.DeviceSim(w2, precision=prec)
    dt, order = args.dt, native.ORDER_CARS_LIGHTS
    s2.step(dt, 200, order)
    torch.cuda.synchronize()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = s2.launch_count
    a0.record()
    s2.step(dt, 4000, order)
    a1.record()
    torch.cuda.synchronize()
    ms2 = a0.elapsed_time(a1)
    us = ms2 / 4000 * 1e3
    return {"workload": "c2_manhattan_2k", "cars": int(w2.n_cars), "lights": int(w2.n_lights), "steps": 4000,
            "value": w2.n_cars * 4000 / (ms2 * 1e-3), "unit": "vehicle-updates/s", "us_per_step": us,
            "gpu_launches": int(s2.launch_count - l0),
            "kernel": "resident_kernel (16 CTAs x 128 threads, one cluster, world in registers + distributed shared memory)",
            "bound": "latency: ~930 dependent-issue-limited instructions per car-step on one warp per SM sub-partition "
                     "+ one cluster barrier per step; HBM is touched only at load / write-back",
            "hbm_bytes_per_step": 0}


def ensemble_block(args, prec, rank, world):
    """BASELINE configs[4]: R rollouts of the 2k-car world with per-replica light timings, round-robin over the ranks,
    each run until its agent arrives (FrontView.end_of_route) or max_steps; [trip_time, steps, done, id] + the reward of
    environment.py:126-150 gathered with one all_gather.  Strong scaling: R is fixed, ranks share it."""
    import torch
    import torch.distributed as dist
    from traffic_b200 import ensemble, native
    w = make_workload("c2_manhattan_2k", args.seed)
    agent = ensemble.pick_agent(w) if args.agent < 0 else args.agent
    ids = ensemble.shard(args.replicas, world, rank)
    ens = ensemble.Ensemble(w, ids, agent_car=agent, seed=args.seed, write_aux=False, precision=prec)
    ens.sim.step(args.dt, 3, native.ORDER_LIGHTS_CARS)       # warm up (compiles nothing, touches everything)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ens.sim.launch_count
    e0.record()
    res = ens.rollout(args.dt, args.ensemble_max_steps)       # [R_local, 4]
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = ens.sim.launch_count - l0
    allres = ensemble.gather_results(res, args.replicas)      # NCCL all_gather (the only collective on this path)
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    steps_run = np.nan_to_num(allres[:, 1], nan=0.0)
    car_steps = float(w.n_cars * steps_run.sum())
    done = np.nan_to_num(allres[:, 2], nan=0.0) > 0
    rewards = ensemble.rewards_of(allres[:, 0], args.dt)      # Env.step's reward if the replicas were successive episodes
    return {"workload": "c5_ensemble", "replicas": int(args.replicas), "cars_per_replica": int(w.n_cars),
            "replicas_per_gpu": int(len(ids)), "agent_car": int(agent), "max_steps": int(args.ensemble_max_steps),
            "arrived": int(done.sum()), "mean_trip_time_arrived": float(allres[done, 0].mean()) if done.any() else None,
            "mean_steps": float(steps_run.mean()), "reward_sum": float(np.sum(rewards)),
            "value": car_steps / (ms_max * 1e-3), "unit": "vehicle-updates/s", "seconds": ms_max * 1e-3,
            "scaling": "strong", "n_gpus": int(world), "gpu_launches": int(launches),
            "gathered": "[trip_time, steps, done, id] per replica, one all_gather; reward per environment.py:126-150"}


def run_reference_arm(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the Python reference itself cannot travel to the
    GPU box) with every host thread, same workload / metric / unit / config.  At N > 1 rank 0 advances N independent
    replicas one after the other (the native arm advances N, one per GPU), so the two arms do the same work."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle
    cores = os.cpu_count() or 1
    n_rep = max(1, int(args.gpus))
    dt, order = args.dt, 0
    total_s, w0 = 0.0, None
    for r in range(n_rep):
        # c5_ensemble: the CPU arm advances plain replicas of the 2k-car world (replicas are independent, equal in cost)
        w = make_workload("c2_manhattan_2k" if args.workload == "c5_ensemble" else args.workload, args.seed + r)
        w0 = w0 or w
        for _ in range(args.warmup if r == 0 else 1):
            oracle.run(w, dt, 1, order, threads=cores)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            oracle.run(w, dt, 1, order, threads=cores)
        total_s += time.perf_counter() - t0
    val = n_rep * w0.n_cars * args.steps / total_s
    line = {"impl": "reference", "metric": "vehicle-updates/sec", "value": val, "unit": "vehicle-updates/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_s / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(args, w0, n_rep),
            "cpu_baseline": {"value": val, "unit": "vehicle-updates/s", "cores": cores, "kind": "port",
                             "sample": "the full workload, %d timed steps of %d replica(s) one after the other, pthread "
                                       "fan-out over cars" % (args.steps, n_rep)},
            "e2e": {"value": val, "unit": "vehicle-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "C port of the reference algorithm (oracle/traffic_oracle.c, pinned bit-exact to the reference's "
                    "golden vectors); the unmodified Python reference measured 300-450 vehicle-updates/s on one core "
                    "(tests/golden/*.npz ref_updates_per_s)"}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--workload", default="c4_city_1m",
                    help="c4_city_1m | c3_sf_50k | c2_manhattan_2k: one simulation per GPU")
    ap.add_argument("--replicas", type=int, default=4096, help="ensemble block: total number of rollouts (BASELINE configs[4])")
    ap.add_argument("--ensemble-max-steps", type=int, default=3000,
                    help="ensemble block: rollout cap.  SURVEY c5 names 20,000 (1.75 s on one GPU, 4,076 of 4,096 agents arrive); "
                         "after ~3,000 steps a few hundred replicas are left and a step shrinks towards the launch "
                         "floor, so the default measures the part of the sweep that exercises the kernel")
    ap.add_argument("--agent", type=int, default=-1, help="ensemble block: agent car (-1: ensemble.pick_agent; learn.py uses 17)")
    ap.add_argument("--neighbor", default="atomic", choices=["atomic", "sort"],
                    help="how the per-bin leader table is built: atomicMin at the kernel tail (default) or the radix-sort "
                         "pipeline (same results)")
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"],
                    help="f64: parity mode (reference arithmetic, bit-exact discrete outputs); f32: preview mode, no parity claim")
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--dt", type=float, default=1e-3)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--frame-every", type=int, default=100, help="k: steps between position frames in the e2e leg")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-small", action="store_true", help="skip the small_world (2k cars) and facade (33 cars) blocks")
    ap.add_argument("--no-ensemble", action="store_true", help="skip the ensemble (c5) block")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from traffic_b200 import native

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    dt, order = args.dt, native.ORDER_CARS_LIGHTS  # the profile script's order
    prec = native.F64 if args.precision == "f64" else native.F32
    real_bytes = 8 if prec == native.F64 else 4
    nmode = native.NEIGHBOR_SORT if args.neighbor == "sort" else native.NEIGHBOR_ATOMIC
    w = make_workload(args.workload, args.seed + rank)  # every rank: an independent replica (different seed)
    sim = native.DeviceSim(w, precision=prec, write_aux=True, neighbor_mode=nmode)
    N = w.n_cars

    # ---- device-resident throughput ----
    # (the clock sampler is started BEFORE the warm-up steps, so that the GPU does not sit idle between the warm-up and
    # the timed region while nvidia-smi spins up)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier()
    sim.step(dt, args.warmup, order)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = sim.launch_count
    barrier()
    t_wall0 = time.time()
    ev0.record()
    sim.step(dt, args.steps, order)
    ev1.record()
    barrier()
    t_wall1 = time.time()
    ms = ev0.elapsed_time(ev1)
    launches = sim.launch_count - l0
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    tmax = torch.tensor([ms], device="cuda", dtype=torch.float64)
    cars_total = torch.tensor([float(N)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(cars_total, op=dist.ReduceOp.SUM)
    ms_max = float(tmax.item())
    value = float(cars_total.item()) * args.steps / (ms_max * 1e-3)

    # replica statistics of the timed simulation, gathered over NCCL further down
    rt = sim.tensor(native.ROUTE_TIME)
    cur = sim.tensor(native.CURSOR)
    end = sim.tensor(native.PATH_END)
    stats = torch.stack([rt.mean(), (cur >= end).double().mean()]).to(torch.float32)

    # ---- end to end through the host-buffer C-ABI call: at least 8 frame intervals of k steps, whatever --steps is ----
    k = max(1, args.frame_every)
    frames = max(8, -(-args.steps // k))
    # a FRESH simulation of the same world, warmed up by the same number of steps, so that the end-to-end leg advances
    # through the same phase of the simulation as the device-resident leg (step cost drifts as queues build up)
    del sim
    sim = native.DeviceSim(w, precision=prec, write_aux=True, neighbor_mode=nmode)
    sim.step(dt, max(args.warmup - 2 * (k // 2), 0), order)
    # pinned host buffers, allocated once: light-face states in / out, two position frames (frame f lands while f+1 runs)
    go_in = torch.from_numpy(sim.download(native.FACE_GO).copy()).pin_memory()
    go_out = torch.empty_like(go_in).pin_memory()
    frames_buf = [torch.empty((N, 2), dtype=torch.float64 if prec == native.F64 else torch.float32).pin_memory()
                  for _ in range(2)]
    for f in range(2):   # warm: both frame buffers go through the copy engine once
        sim.step_host_async(dt, k // 2, order, go_in, frames_buf[f], go_out)
        sim.frame_wait()
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    g0.record()
    for f in range(frames):
        go_in.copy_(go_out)                      # this interval's input: the light-face states, from pinned host memory
        sim.step_host_async(dt, k, order, go_in, frames_buf[f & 1], go_out)
        sim.frame_wait(faces_only=True)          # the next interval's input is on the host; the positions still travel
    sim.frame_wait()                             # last position frame landed
    g1.record()
    torch.cuda.synchronize()
    e2e_s = max(time.perf_counter() - t0, g0.elapsed_time(g1) * 1e-3)   # host clock around device work AND copies
    te = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = float(cars_total.item()) * frames * k / float(te.item())
    h2d = go_in.numel() / k
    d2h = (frames_buf[0].numel() * real_bytes + go_out.numel()) / k
    del sim

    # ---- replica statistics gathered over NCCL ----
    if world > 1:
        allstats = torch.empty(world * 2, device="cuda", dtype=torch.float32)
        dist.all_gather_into_tensor(allstats, stats)
    else:
        allstats = stats
    allstats = allstats.cpu().numpy().reshape(-1, 2)

    # ---- BASELINE configs[4]: the rollout ensemble, sharded over the ranks (every N) ----
    ens_block = None
    if not args.no_ensemble:
        try:
            ens_block = ensemble_block(args, prec, rank, world)
        except Exception as e:  # the headline must not depend on the secondary measurement
            ens_block = {"error": repr(e)}

    if rank == 0:
        peak, peak_src = peaks()
        # mean duration of the step kernel; with --neighbor sort the sort launches share the step, so this is an upper bound
        kernel_us = ms / args.steps * 1e3
        algo = ALGO_BYTES[args.precision] * N
        achieved = algo / (kernel_us * 1e-6) / 1e9
        traffic = None
        tp = os.path.join(REPO, "profiles", "traffic_bytes.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(args.workload + ":" + args.precision)
            except Exception:
                traffic = None
        big = N >= 2 * 128 * 5 * 148
        line = {
            "metric": "vehicle-updates/sec", "value": value, "unit": "vehicle-updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": config_of(args, w, world),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic,
                         "kernel": ("pipe_kernel<%s>" if big else "step_kernel<%s>") % ("double" if prec == native.F64 else "float"),
                         "kernel_us": kernel_us, "algorithmic_bytes_per_car_step": ALGO_BYTES[args.precision],
                         "peak_source": peak_src},
            "e2e": {"value": e2e_value, "unit": "vehicle-updates/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "frames": frames, "steps_per_frame": k, "steps": frames * k},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "replica_stats": {"mean_route_time": [float(x) for x in allstats[:, 0]],
                              "parked_fraction": [float(x) for x in allstats[:, 1]]},
        }
        if ens_block is not None:
            line["ensemble"] = ens_block
        if not args.no_cpu_baseline and world == 1:   # reported at N=1 only
            v, s_, t_ = cpu_port_updates_per_s(w, dt, order, args.cpu_budget, threads=1)
            line["cpu_baseline"] = {"value": v, "unit": "vehicle-updates/s", "cores": 1, "kind": "port",
                                    "sample": "%d steps of the same %d-car workload (%.1f s), C oracle port, 1 thread; "
                                              "the unmodified Python reference runs 300-450 updates/s on one core"
                                              % (s_, w.n_cars, t_)}
        if not args.no_small and world == 1:
            for name, fn in (("small_world", lambda: small_world_block(args, prec)), ("facade", facade_profile_loop)):
                try:
                    line[name] = fn()
                except Exception as e:  # the headline must not depend on the secondary measurements
                    line[name] = {"error": repr(e)}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
