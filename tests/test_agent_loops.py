"""The reference's agent loops consume the environment as a drop-in (GPU tests):
  * the loop body of agents/DDQN/DDQN.py:108-121 (pick_action -> env.step -> save_experience -> learn) on
    FJSPEnv, the CPU oracle stepped side by side with the same actions and draws;
  * the batched loop the vector environment gives the MPPPO / DA3C / HMPSAC agents (one policy forward + one
    step() of every copy per iteration: deep_reinforcement_learning_for_fjsp_b200.agent_loop.PolicyRollout, a
    CUDA graph), its actions replayed on the oracle."""
import numpy as np
import pytest

import parity_common as pc

pytestmark = pytest.mark.gpu


def test_ddqn_loop_body_on_the_drop_in_env():
    import torch
    import oracle_py
    from deep_reinforcement_learning_for_fjsp_b200.vec_env import FJSPEnv
    from deep_reinforcement_learning_for_fjsp_b200.agent_loop import make_mlp
    torch.manual_seed(0)
    insts, _ = pc.random_batch(61, "SO_DFJSP", 1, 1)
    env = FJSPEnv(variant="SO_DFJSP", instance=insts[0], seed=3)
    ora = oracle_py.OracleEnv(insts[0].to_blob(), "SO_DFJSP")
    draws = np.random.default_rng(3)                      # FJSPEnv(seed=3) draws the same words for its random rules
    n_actions = env.actions_size[0] * env.actions_size[1]
    q_local, q_target = make_mlp(env.state_size, n_actions, seed=1), make_mlp(env.state_size, n_actions, seed=1)
    opt = torch.optim.Adam(q_local.parameters(), lr=1e-3, eps=1e-4)
    memory, rng, epsilon, gamma = [], np.random.default_rng(9), 1.0, 0.99
    for episode in range(2):                              # DDQN.step(): one episode per call, the env object is re-used
        state, so = env.reset(), ora.reset()
        pc.assert_states_close(state, so, "reset")
        done = False
        while not done:
            # pick_action: epsilon-greedy on the local Q network (agents/DDQN/DDQN.py:151-163)
            with torch.no_grad():
                q = q_local(torch.from_numpy(state).float().unsqueeze(0))
            epsilon = max(0.01, epsilon - 0.01)
            action = int(rng.integers(n_actions)) if rng.random() < epsilon else int(torch.argmax(q))
            next_state, reward, done = env.step([action])                      # the flat composite-rule index
            rule = env.action_tuple[action]
            so, ro, do, _ = ora.step(rule, draws.integers(0, 2**32, (1, 1, 2), dtype=np.uint64).astype(np.uint32)[0, 0])
            assert reward == ro and done == do
            pc.assert_states_close(next_state, so, f"episode {episode} step {env.step_count}")
            memory.append((state, action, reward, next_state, done))           # save_experience
            state = next_state
        assert env.delay_time_sum == ora.info()["delay_sum"]
        # learn(): double-DQN target on a sampled batch (agents/DDQN/DDQN.py:165-190)
        idx = rng.integers(0, len(memory), 32)
        s = torch.tensor(np.stack([memory[i][0] for i in idx]), dtype=torch.float32)
        a = torch.tensor([memory[i][1] for i in idx]).unsqueeze(1)
        r = torch.tensor([memory[i][2] for i in idx], dtype=torch.float32).unsqueeze(1)
        s2 = torch.tensor(np.stack([memory[i][3] for i in idx]), dtype=torch.float32)
        d = torch.tensor([float(memory[i][4]) for i in idx]).unsqueeze(1)
        with torch.no_grad():
            best = q_local(s2).argmax(1, keepdim=True)
            target = r + gamma * q_target(s2).gather(1, best) * (1 - d)
        loss = torch.nn.functional.mse_loss(q_local(s).gather(1, a), target)
        opt.zero_grad(); loss.backward(); opt.step()
        assert torch.isfinite(loss)
    assert len(memory) > 40


@pytest.mark.parametrize("variant,use_graph", [("MO_DFJSP", True), ("SO_DFJSP", False)])
def test_policy_in_the_loop_rollout_matches_oracle(variant, use_graph):
    """One policy forward + one step() of every copy per iteration, on the device (CUDA graph or eager):
    the sampled actions and draws are replayed on the oracle; every transition must agree."""
    import torch
    import oracle_py
    from deep_reinforcement_learning_for_fjsp_b200.vec_env import FJSPVecEnv
    from deep_reinforcement_learning_for_fjsp_b200.agent_loop import PolicyRollout, make_mlp
    insts, env_instance = pc.random_batch(71, variant, 6, 40)
    blobs = [i.to_blob() for i in insts]
    vec = FJSPVecEnv(None, env_instance, variant, blobs=blobs)
    nt, nm = vec.actions_size
    net = make_mlp(vec.state_size, nt * nm, device=vec.dev, seed=2)
    envs = [oracle_py.OracleEnv(blobs[k], variant) for k in env_instance]
    o0 = np.stack([e.reset() for e in envs])
    # (PolicyRollout resets the batch and, with a graph, runs three warm-up iterations plus the capture)
    pr = PolicyRollout(vec, net, reward_policy=1, use_graph=False)
    np.testing.assert_allclose(pr.state.cpu().numpy(), o0, rtol=1e-6, atol=1e-6)
    if use_graph:
        pr.graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=vec.dev)
        side.wait_stream(torch.cuda.current_stream(vec.dev))
    n_steps, B = 80, len(env_instance)
    captured = False
    for t in range(n_steps):
        if use_graph and not captured and t == 3:
            # capture one iteration on the live state: the capture itself does not run it
            torch.cuda.synchronize()
            with torch.cuda.graph(pr.graph):
                pr._body()
            captured = True
        if captured:
            pr.graph.replay()
        else:
            pr._body()
        torch.cuda.synchronize()
        a = pr.actions.cpu().numpy()                                           # [1, B, 2] what the policy chose
        r = pr.rnd.cpu().numpy().view(np.uint32)
        ref = oracle_py.batch_rollout(envs, a, r, 1, want_rec=False)
        assert np.array_equal(pr.out["done"].cpu().numpy(), ref["done"]), t
        assert np.array_equal(pr.out["reward"].cpu().numpy(), ref["reward"]), t
        np.testing.assert_allclose(pr.out["state"].cpu().numpy(), ref["state"], rtol=2e-6, atol=1e-6)   # float32 observation
    assert (vec.info()["error"] == 0).all()
