"""GPU parity of the DDPG / TD3 learn steps against the torch-CPU oracle (oracle/actor_critic.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import actor_critic as oac     # noqa: E402  (checker only)
from test_learn_gpu import close           # noqa: E402


def _make(twin, B, D=17, A=6, seed=0):
    from coach_b200.agents.ddpg_agent import DDPGAgent, DDPGAgentParameters, TD3Agent, TD3AgentParameters
    from coach_b200.memories.memory import MemoryGranularity
    ap = TD3AgentParameters() if twin else DDPGAgentParameters()
    ap.memory.max_size = (MemoryGranularity.Transitions, 4096)
    ap.network_wrappers["actor"].batch_size = ap.network_wrappers["critic"].batch_size = B
    cls = TD3Agent if twin else DDPGAgent
    return cls(ap, observation_dim=D, action_dim=A, seed=seed)


@pytest.mark.parametrize("twin", [False, True])
def test_learn_from_batch_matches_oracle(twin):
    from coach_b200.core_types import DeviceBatch
    B, D, A = 256, 17, 6
    ag = _make(twin, B)
    rng = np.random.RandomState(4)
    dev = ag.device
    # de-synchronise targets from online nets
    ag.actor.target.copy_(ag.actor.store.theta * 0.95 + 0.002)
    ag.critic.target.copy_(ag.critic.store.theta * 0.9 - 0.001)
    cols = {"state:observation": rng.randn(B, D).astype(np.float32),
            "next_state:observation": rng.randn(B, D).astype(np.float32),
            "action": np.tanh(rng.randn(B, A)).astype(np.float32), "reward": rng.randn(B),
            "game_over": (rng.rand(B) < 0.1).astype(np.uint8)}
    batch = DeviceBatch({k: torch.from_numpy(v).to(dev) for k, v in cols.items()}, B)
    noise = rng.normal(0, 0.2, (B, A))
    pa, pc = ag.ap.network_wrappers["actor"], ag.ap.network_wrappers["critic"]
    for step in range(2):
        actor, actor_t = ag.actor.store.export_named(), ag.actor.store.export_named(ag.actor.target)
        critic, critic_t = ag.critic.store.export_named(), ag.critic.store.export_named(ag.critic.target)
        if step == 0:
            opt_a = oac.make_adam(actor, pa.learning_rate, 0.9, 0.999, 1e-8)
            opt_c = oac.make_adam(critic, pc.learning_rate, 0.9, 0.999, 1e-8)
        ag.training_iteration += 1
        if twin:
            loss, _, _ = ag.learn_from_batch(batch, noise=noise)
            upd = ag.training_iteration % 2 == 0
        else:
            loss, _, _ = ag.learn_from_batch(batch)
            upd = True
        torch.cuda.synchronize()
        ref = oac.ddpg_td3_step(actor, actor_t, critic, critic_t, opt_a, opt_c,
                                dict(states=cols["state:observation"], next_states=cols["next_state:observation"],
                                     actions=cols["action"], rewards=cols["reward"],
                                     game_overs=cols["game_over"].astype(bool)),
                                twin=twin, noise=noise, update_actor=upd)
        close(ag.td_targets.cpu().numpy(), ref["td_targets"], name="td targets")
        close(loss, ref["loss"], name="critic loss")
        gc = ag.critic.store.export_named(ag.critic.store.grad)
        for n in ref["critic_grads"]:
            close(gc[n], ref["critic_grads"][n].numpy(), name="critic grad " + n)
        pc_new = ag.critic.store.export_named()
        for n in ref["new_critic"]:
            # Adam with eps = 1e-8 turns every gradient, however tiny, into a step of ~lr: entries whose gradient is
            # at rounding-noise level legitimately differ by a small fraction of lr
            close(pc_new[n], ref["new_critic"][n].numpy(), name="critic param " + n, atol=1e-2 * pc.learning_rate)
        if ref["actor_grads"] is not None:
            ga = ag.actor.store.export_named(ag.actor.store.grad)
            for n in ref["actor_grads"]:
                close(ga[n], ref["actor_grads"][n].numpy(), name="actor grad " + n)
        pa_new = ag.actor.store.export_named()
        for n in ref["new_actor"]:
            close(pa_new[n], ref["new_actor"][n].numpy(), name="actor param " + n, atol=1e-2 * pa.learning_rate)


def test_train_driver_with_episodic_replay_and_polyak():
    B, D, A = 64, 17, 6
    ag = _make(True, B)
    rng = np.random.RandomState(0)
    n = 512
    done = np.zeros(n, np.uint8)
    done[99::100] = 1
    done[-1] = 1
    ag.memory.store_columns({"state:observation": rng.randn(n, D).astype(np.float32),
                             "next_state:observation": rng.randn(n, D).astype(np.float32),
                             "action": np.tanh(rng.randn(n, A)).astype(np.float32), "reward": rng.randn(n),
                             "game_over": done})
    ag.ap.algorithm.num_consecutive_training_steps = 4
    t0 = ag.actor.target.clone()
    np.random.seed(3)
    loss = ag.train()
    assert np.isfinite(loss) and ag.training_iteration == 4
    # TD3: TrainingSteps(2) cadence -> two polyak updates with rate 0.005 happened
    assert not torch.equal(t0, ag.actor.target)
    diff = (ag.actor.target - ag.actor.store.theta).abs().max().item()
    assert diff > 0          # soft update, not a hard copy


@pytest.mark.parametrize("kind", ["ddpg", "td3", "sac"])
def test_graph_replay_of_the_actor_critic_steps_matches_eager(kind, monkeypatch):
    """the CUDA-graph replay of the DDPG / TD3 / SAC learn step (agents/ddpg_agent.py: GraphedKernels) leaves exactly
    the parameters the eager launch sequence leaves"""
    from coach_b200.memories.memory import MemoryGranularity
    res = []
    for graph in (0, 1):
        monkeypatch.setenv("CB200_AC_GRAPH", str(graph))
        if kind == "sac":
            from coach_b200.agents.soft_actor_critic_agent import SoftActorCriticAgent as cls, \
                SoftActorCriticAgentParameters as P
        elif kind == "td3":
            from coach_b200.agents.ddpg_agent import TD3Agent as cls, TD3AgentParameters as P
        else:
            from coach_b200.agents.ddpg_agent import DDPGAgent as cls, DDPGAgentParameters as P
        ap = P()
        ap.memory.max_size = (MemoryGranularity.Transitions, 4096)
        for nw in ap.network_wrappers.values():
            nw.batch_size = 64
        ag = cls(ap, observation_dim=17, action_dim=6, seed=3)
        rng = np.random.RandomState(1)
        n = 2000
        cols = {"state:observation": rng.randn(n, 17).astype(np.float32),
                "next_state:observation": rng.randn(n, 17).astype(np.float32),
                "action": rng.uniform(-1, 1, (n, 6)).astype(np.float32), "reward": rng.randn(n)}
        done = np.zeros(n, np.uint8)
        done[99::100] = 1
        cols["game_over"] = done
        ag.memory.store_columns(cols)
        np.random.seed(5)
        losses = []
        for step in range(7):                       # 2 eager steps, capture, replays (TD3: both graph variants)
            ag.total_steps_counter += 1
            losses.append(ag.train(fetch=True))
        torch.cuda.synchronize()
        stores = [getattr(ag, t).store.theta.clone() for t in ("actor", "critic", "policy", "q", "v") if hasattr(ag, t)]
        res.append((losses, stores))
        if graph:
            assert ag._graph_step.graph is not None and ag._graph_step.launches > 20
    assert res[0][0] == res[1][0]
    for a, b in zip(res[0][1], res[1][1]):
        assert torch.equal(a, b)
