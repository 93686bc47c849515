#!/usr/bin/env python
"""Randomised light-timing sweep / Q-learning style rollout ensemble (BASELINE configs[4]): R independent replicas of the
2k-car Manhattan-sized world, every replica with its own light switch times, sharded round-robin over the GPUs of one
node; each rank advances its replicas in ONE batched device simulation until every agent has arrived (or max_steps), then a
single all_gather collects [trip_time, steps, done] per replica and rank 0 prints the sweep summary + the rewards of
Env.step (environment.py:126-150) for the sequence of trip times.

    python examples/ensemble_sweep.py --replicas 64 --max-steps 3000
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/ensemble_sweep.py --replicas 4096
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch                                     # noqa: E402
import torch.distributed as dist                 # noqa: E402

from traffic_b200 import ensemble, synth         # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--replicas", type=int, default=64)
    ap.add_argument("--max-steps", type=int, default=3000)
    ap.add_argument("--agent", type=int, default=-1,
                    help="agent car (learn.py:14 uses 17); -1 picks a car with a short route whose last edge runs along x - "
                         "FrontView.end_of_route wants 5 m on both axes but a car parks when it enters the 5.7 m x 42 m "
                         "crossing window of its last route point, so many agents never 'arrive' (SURVEY Appendix B-9)")
    ap.add_argument("--dt", type=float, default=1e-3)
    ap.add_argument("--seed", type=int, default=0)
    opt = ap.parse_args()

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    if world > 1:
        dist.init_process_group("nccl")
    city, _ = synth.make_city(**synth.CITY_CONFIGS["c2_manhattan_2k"])   # same world on every rank (same seed)
    agent = opt.agent
    if agent < 0:
        lens = city.path_end - city.path_start
        last = city.path_xy[city.path_end - 1] - city.path_xy[np.maximum(city.path_end - 2, city.path_start)]
        ok = (lens >= 3) & (lens <= 8) & (np.abs(last[:, 0]) > 2.0 * np.abs(last[:, 1]))
        agent = int(np.flatnonzero(ok)[0])
    mine = ensemble.shard(opt.replicas, world, rank)
    ens = ensemble.Ensemble(city, mine, agent_car=agent, seed=opt.seed)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    local = ens.rollout(opt.dt, opt.max_steps, chunk=500)
    torch.cuda.synchronize()
    el = time.perf_counter() - t0
    res = ensemble.gather_results(local, opt.replicas)                   # [R, 4]: trip_time, steps, done, replica id
    if rank == 0:
        done = res[:, 2] > 0
        print("%d replicas x %d cars on %d GPU(s), agent = car %d: %d agents arrived within %d steps; rollout %.2f s"
              % (opt.replicas, city.n_cars, world, agent, int(done.sum()), opt.max_steps, el))
        if done.any():
            print("trip time of the arrived agents: min %.3f  median %.3f  max %.3f (simulated seconds)"
                  % (res[done, 0].min(), np.median(res[done, 0]), res[done, 0].max()))
        history, rewards = [], []
        for ep in range(min(opt.replicas, 12)):                         # treat the first replicas as successive episodes
            history.append(res[ep, 0])
            rewards.append(ensemble.reward(history, (ep, opt.replicas), opt.dt)[0])
        print("Env.step rewards of the first episodes:", ["%.3f" % r for r in rewards])
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
