"""Categorical DQN (C51) whole learn step on the GPU against the oracle restatement (oracle/c51.py, whose numpy part is
pinned to the reference agents by tests/test_oracle_golden.py).  Reference: rl_coach/agents/categorical_dqn_agent.py:
105-165, architectures/tensorflow_components/heads/categorical_q_head.py:41-57."""
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def close(got, want, rtol=1e-5, name="", atol=0.0):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    err = np.abs(got - want).max() if got.size else 0.0
    tol = rtol * np.abs(want).max() + atol
    assert err <= tol, "%s: max abs err %.3e > %.3e" % (name, err, tol)


def _agent(obs_shape, A, B, per, atoms=51):
    from coach_b200.agents.categorical_dqn_agent import CategoricalDQNAgent, CategoricalDQNAgentParameters
    from coach_b200.memories.memory import MemoryGranularity
    from coach_b200.memories.prioritized_experience_replay import PrioritizedExperienceReplayParameters
    from coach_b200.schedules import LinearSchedule
    ap = CategoricalDQNAgentParameters()
    ap.algorithm.atoms = atoms
    if per:
        ap.memory = PrioritizedExperienceReplayParameters()
        ap.memory.beta = LinearSchedule(0.4, 1, 1000)
    ap.memory.max_size = (MemoryGranularity.Transitions, 1024)
    ap.network_wrappers["main"].batch_size = B
    return CategoricalDQNAgent(ap, observation_shape=obs_shape, num_actions=A, seed=0)


@pytest.mark.parametrize("obs,A,B,per", [((4,), 2, 32, False), ((84, 84, 4), 6, 16, True), ((84, 84, 4), 6, 128, True)],
                         ids=["cartpole_B32", "atari_per_B16", "atari_per_B128"])
def test_c51_learn_step_matches_oracle(obs, A, B, per):
    from oracle import c51, nets as on
    from test_learn_gpu import _device_relu_masks
    torch.manual_seed(0)
    agent = _agent(obs, A, B, per)
    N = 51
    assert agent.head_outputs == A * N and agent.head_desc is None
    rng = np.random.RandomState(5)
    n = max(256, 2 * B)
    if len(obs) == 3:
        s = rng.randint(0, 256, (n,) + obs).astype(np.uint8)
        s2 = rng.randint(0, 256, (n,) + obs).astype(np.uint8)
    else:
        s = rng.uniform(-1, 1, (n,) + obs).astype(np.float32)
        s2 = rng.uniform(-1, 1, (n,) + obs).astype(np.float32)
    a = rng.randint(0, A, n).astype(np.int64)
    r = rng.choice([-1.0, 0.0, 1.0, 0.37, 11.0], n).astype(np.float64)
    done = (rng.rand(n) < 0.2).astype(np.uint8)
    agent.memory.store_columns({"state:observation": s, "next_state:observation": s2, "action": a, "reward": r,
                                "game_over": done})
    if per:
        agent.memory.update_priorities(np.arange(n), np.abs(rng.randn(n)))
    store = agent.net_def.store
    net = agent.networks["main"]
    net.theta_target.copy_(store.theta * 0.9 + 0.01)
    net.target_changed()
    oracle32 = on.QNetOracle(obs, A * N, False, torch.float32)
    oracle64 = on.QNetOracle(obs, A * N, False, torch.float64)
    z = c51.z_values(-10.0, 10.0, N)
    np.testing.assert_array_equal(agent.z_values, z)
    for step in range(2):
        online_named = store.export_named()
        target_named = store.export_named(net.theta_target)
        random.seed(20 + step)
        np.random.seed(20 + step)
        batch = agent.sample_batch()
        if step == 0:
            opt32 = on.AdamTF([torch.from_numpy(v) for v in online_named.values()], 2.5e-4, 0.9, 0.99, 1e-4)
        loss, losses, gnorm = agent.learn_from_batch(batch)
        torch.cuda.synchronize()
        for k in ("state:observation", "next_state:observation"):
            batch.column(k)
        cols = {k: v.cpu().numpy() for k, v in batch.columns.items()}
        ob = dict(states=cols["state:observation"], next_states=cols["next_state:observation"],
                  actions=cols["action"], rewards=cols["reward"], game_overs=cols["game_over"].astype(bool))
        masks = _device_relu_masks(agent)
        k32, k64 = dict(masks=masks, tol=1e-5), dict(masks=masks, tol=1e-5)
        ref = c51.c51_learn_step(oracle32, oracle32.cast(online_named), oracle32.cast(target_named), opt32, ob, 0.99,
                                 z, A, kink=k32)
        opt64 = on.AdamTF([torch.from_numpy(v).double() for v in online_named.values()], 2.5e-4, 0.9, 0.99, 1e-4,
                          dtype=torch.float64)
        ref64 = c51.c51_learn_step(oracle64, oracle64.cast(online_named), oracle64.cast(target_named), opt64, ob, 0.99,
                                   z, A, kink=k64)
        assert k32.get("hard", 0) == 0 and k64.get("hard", 0) == 0, "ReLU masks differ away from the kink"
        np.testing.assert_array_equal(agent.target_actions.cpu().numpy(), ref["target_actions"])
        close(agent.targets.cpu().numpy().reshape(B, A, N), ref["targets"], name="TD_targets", atol=1e-7)
        close(agent.loss_rows.cpu().numpy(), ref["loss_rows"], name="loss rows")
        close(agent.td_err.cpu().numpy(), ref["td_errors"], name="PER errors")
        close(agent.q_online.cpu().numpy(), ref["q_online"], name="q_online", atol=1e-6)
        close(loss, ref["loss"], name="loss")
        close(gnorm, ref["grad_norm"], name="grad_norm", rtol=2e-5)
        got_grads = store.export_named(store.grad)
        for name in ref["grads"]:
            want = ref["grads"][name].numpy()
            e_ours = np.abs(got_grads[name] - ref64["grads"][name].numpy()).max()
            e_orc = np.abs(want - ref64["grads"][name].numpy()).max()
            try:
                close(got_grads[name], want, name="grad " + name)
            except AssertionError as exc:
                # ill-conditioned weight-gradient sums: not farther from the fp64 evaluation than the fp32 oracle is
                # (same clause as tests/test_learn_gpu.py)
                assert e_ours <= 1.5 * e_orc, "%s; vs fp64: ours %.3e, fp32 oracle %.3e" % (exc, e_ours, e_orc)
            assert e_ours <= 4 * e_orc + 2e-6 * (np.abs(want).max() + 1e-30), (name, e_ours, e_orc)
        got_params = store.export_named()
        for name in ref["new_params"]:
            want = ref["new_params"][name].numpy()
            try:
                close(got_params[name], want, name="param " + name)
            except AssertionError as exc:
                w64 = ref64["new_params"][name].numpy()
                e_ours, e_orc = np.abs(got_params[name] - w64).max(), np.abs(want - w64).max()
                assert e_ours <= 2 * e_orc, "%s; vs fp64: ours %.3e, fp32 oracle %.3e" % (exc, e_ours, e_orc)
    if per:
        idx = cols["idx"]
        leaves = agent.memory.sum_tree.cpu().numpy()[agent.memory.power_of_2_size - 1:]
        td = agent.td_err.cpu().numpy()
        last = {int(i): k for k, i in enumerate(idx)}
        for i, k in last.items():
            assert leaves[i] == (td[k] + 1e-6) ** 0.6


def test_c51_acting_q_values():
    """get_all_q_values_for_states = the head's q_values output: tensordot(softmax fp64, fp32-rounded support)"""
    agent = _agent((4,), 3, 32, False)
    x = np.random.RandomState(1).uniform(-1, 1, (16, 4)).astype(np.float32)
    q = agent.get_all_q_values_for_states(x).cpu().numpy()
    inst = agent._acting[16][1]
    logits = inst.q.cpu().numpy().reshape(16, 3, 51).astype(np.float64)
    e = np.exp(logits - logits.max(-1, keepdims=True))
    p = (e / e.sum(-1, keepdims=True)).astype(np.float32).astype(np.float64)
    want = p @ agent.z_values.astype(np.float32).astype(np.float64)
    np.testing.assert_allclose(q, want, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("kind", ["c51", "dueling_nomw"])
def test_graph_replay_matches_eager_on_the_general_head_path(monkeypatch, kind):
    """the CUDA-graph / multi-stream schedule of the learn step is bit-identical to the eager launch sequence also where
    the head is not the fused DQN head: Categorical DQN (loss in the forward part) and the dueling network of BASELINE
    config 5 (global-norm clipping: the norm stays in the backward graph)"""
    results = []
    for graph in (0, 1):
        monkeypatch.setenv("CB200_DQN_GRAPH", str(graph))
        torch.manual_seed(0)
        if kind == "c51":
            agent = _agent((84, 84, 4), 6, 128, True)
        else:
            from test_learn_gpu import _make_agent
            agent = _make_agent((84, 84, 4), 6, 128, True, True, True, 10.0, True, seed=5, middleware=False)
        assert agent.use_graph == bool(graph)
        rng = np.random.RandomState(3)
        n = 512
        agent.memory.store_columns({
            "state:observation": rng.randint(0, 256, (n, 84, 84, 4)).astype(np.uint8),
            "next_state:observation": rng.randint(0, 256, (n, 84, 84, 4)).astype(np.uint8),
            "action": rng.randint(0, 6, n).astype(np.int64), "reward": rng.randint(-1, 2, n).astype(np.float64),
            "game_over": (rng.rand(n) < 0.1).astype(np.uint8)})
        agent.memory.update_priorities(np.arange(n), np.abs(rng.randn(n)))
        losses = []
        for step in range(6):                       # 2 eager steps, capture, 3 replays
            random.seed(20 + step)
            np.random.seed(20 + step)
            agent.total_steps_counter += 4
            losses.append(agent.train() if step % 2 else agent.learn_from_batch(agent.sample_batch())[0])
        torch.cuda.synchronize()
        if graph:
            assert agent._graphs is not None and agent.graph_kernel_launches > 0
        results.append((losses, agent.net_def.store.theta.clone(), agent.memory.sum_tree.clone()))
    assert results[0][0] == results[1][0]
    assert torch.equal(results[0][1], results[1][1])
    assert torch.equal(results[0][2], results[1][2])
