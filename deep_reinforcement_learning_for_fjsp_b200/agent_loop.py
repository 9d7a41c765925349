"""The agent side of the drop-in: a policy network in the loop, one step() per launch.

The reference's agents run   action = policy(state); state, reward, done = env.step(action)
once per environment step (agents/DDQN/DDQN.py:104-121, agents/MPPPO/MPPPO.py:236-250,
agents/HMPSAC/A3C_v5.1.py:267-285, utilities/Parallel_Experience_Generator.py:48-66).  With the
vector environment the same loop body serves every copy at once; `PolicyRollout` keeps the
whole iteration -- policy forward, action sampling, the packing kernels and the step kernel
(T = 1) -- on the device, and replays it as ONE CUDA graph so that the per-step cost is the
kernels, not the launches.  The policy/value networks stay plain PyTorch modules (they are
tiny MLPs, not the hot path); only device memory, streams and graphs are used from torch.
"""
from __future__ import annotations

from typing import Callable, Optional


class PolicyRollout:
    """state -> policy -> (task rule, machine rule) -> FJSPVecEnv.step, per launch.

    policy(state[B, state_size] float32) -> logits [B, n_task_rules * n_machine_rules]
    (the flat composite-rule action space of DDQN / MPPPO) or a pair of logits
    ([B, n_task_rules], [B, n_machine_rules]) (the two actors of DA3C / HMPSAC).
    Actions are sampled with the Gumbel-max trick (argmax(logits + g)), which needs no
    host round trip and can be captured in a CUDA graph.
    """

    def __init__(self, vec, policy: Callable, reward_policy: int = 1, completion: float = 1.0,
                 tardiness: float = 1.0, energy: float = 1.0, use_graph: bool = True, greedy: bool = False):
        import torch
        self.torch, self.vec, self.policy = torch, vec, policy
        self.kw = dict(reward_policy=reward_policy, completion=completion, tardiness=tardiness, energy=energy)
        self.greedy = greedy
        B, S = vec.n_envs, vec.state_size
        dev = vec.dev
        self.state = vec.reset(dtype=torch.float32)                       # [B, S]
        self.actions = torch.zeros((1, B, 2), dtype=torch.int32, device=dev)
        self.rnd = torch.zeros((1, B, 2), dtype=torch.int32, device=dev)
        self.out = {"state": torch.empty((1, B, S), dtype=torch.float32, device=dev),
                    "reward": torch.empty((1, B), dtype=torch.float64, device=dev),
                    "done": torch.empty((1, B), dtype=torch.int32, device=dev)}
        self.reward_sum = torch.zeros(B, dtype=torch.float64, device=dev)
        self.steps = 0
        self.graph: Optional["torch.cuda.CUDAGraph"] = None
        if use_graph:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):                # warm-up outside capture (allocator, lazy init)
                for _ in range(3):
                    self._body()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._body()

    def _sample(self, logits):
        t = self.torch
        if self.greedy:
            return logits.argmax(-1)
        u = t.rand_like(logits).clamp_(1e-12, 1.0)
        return (logits - t.log(-t.log(u))).argmax(-1)

    def _body(self):
        t = self.torch
        nt, nm = self.vec.actions_size
        with t.no_grad():
            o = self.policy(self.state)
            if isinstance(o, (tuple, list)):
                a_task, a_mach = self._sample(o[0]), self._sample(o[1])
            else:
                a = self._sample(o)
                a_task, a_mach = a // nm, a % nm
            self.actions[0, :, 0] = a_task.to(t.int32)
            self.actions[0, :, 1] = a_mach.to(t.int32)
            self.rnd.random_(0, 2**31 - 1)               # draws consumed by the random rules
            self.vec.rollout(self.actions, self.rnd, out=self.out, state_dtype=t.float32, **self.kw)
            self.state.copy_(self.out["state"][0])
            self.reward_sum += self.out["reward"][0]

    def step(self):
        """One env step of every copy (no host synchronisation)."""
        if self.graph is not None:
            self.graph.replay()
        else:
            self._body()
        self.steps += 1

    def run(self, n_steps: int):
        for _ in range(n_steps):
            self.step()
        return self.state, self.out["reward"][0], self.out["done"][0]


def make_mlp(state_size: int, n_out: int, hidden: int = 200, layers: int = 3, device=None, seed: int = 0):
    """The reference's ActorNet shape (agents/DDQN/DDQN.py:84: hidden_size=200, hidden_layer=3)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    mods, n_in = [], state_size
    for _ in range(layers):
        lin = torch.nn.Linear(n_in, hidden)
        mods += [lin, torch.nn.ReLU()]
        n_in = hidden
    mods.append(torch.nn.Linear(n_in, n_out))
    net = torch.nn.Sequential(*mods)
    with torch.no_grad():
        for p in net.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * 0.05)
    return net.to(device) if device is not None else net
