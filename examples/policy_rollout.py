"""Agent-side usage: a PyTorch policy picks one composite dispatching rule per environment
copy and step (the loop of agents/DDQN/DDQN.py:108-121 / agents/HMPSAC/A3C_v5.*.py:259-283,
for a whole batch at once).  States, rewards and done flags never leave the GPU.

    python examples/policy_rollout.py --envs 1024 --steps 200
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_reinforcement_learning_for_fjsp_b200 import FJSPInstance, FJSPVecEnv  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--variant", default="MO_DFJSP")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    insts = [FJSPInstance.generate(i, 1.0, 10, 3, "DA3C") for i in range(64)]
    env = FJSPVecEnv(insts, np.arange(args.envs) % len(insts), args.variant)
    nt, nm = env.actions_size
    q_net = torch.nn.Sequential(torch.nn.Linear(env.state_size, 128), torch.nn.ReLU(),
                                torch.nn.Linear(128, nt * nm)).to(dev)
    state = env.reset(dtype=torch.float32)
    gen = torch.Generator(device=dev).manual_seed(0)
    returns = torch.zeros(args.envs, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        with torch.no_grad():
            a = q_net(state).argmax(1)                                   # composite rule index
        actions = torch.stack([a // nm, a % nm], 1).to(torch.int32).contiguous()
        rnd = torch.randint(-2**31, 2**31 - 1, (args.envs, 2), device=dev, dtype=torch.int32, generator=gen)
        state, reward, done = env.step(actions, rnd, reward_policy=1, state_dtype=torch.float32)
        returns += reward
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    info = env.info()
    print(f"{args.envs * args.steps / dt / 1e6:.2f} M env-steps/s with the policy in the loop; "
          f"episodes finished {int(info['episodes'].sum())}, mean return {returns.mean().item():.1f}, "
          f"errors {int((info['error'] != 0).sum())}")


if __name__ == "__main__":
    main()
