"""GPU parity of the Soft Actor-Critic learn step against the torch-CPU oracle (oracle/sac.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import actor_critic as oac     # noqa: E402  (checker only)
from oracle import sac as osac             # noqa: E402
from test_learn_gpu import close           # noqa: E402


def _make(B, D=17, A=6, seed=0):
    from coach_b200.agents.soft_actor_critic_agent import SoftActorCriticAgent, SoftActorCriticAgentParameters
    from coach_b200.memories.memory import MemoryGranularity
    ap = SoftActorCriticAgentParameters()
    ap.memory.max_size = (MemoryGranularity.Transitions, 4096)
    for k in ("policy", "q", "v"):
        ap.network_wrappers[k].batch_size = B
    return SoftActorCriticAgent(ap, observation_dim=D, action_dim=A, seed=seed)


def test_sac_learn_from_batch_matches_oracle():
    from coach_b200.core_types import DeviceBatch
    B, D, A = 256, 17, 6
    ag = _make(B)
    rng = np.random.RandomState(7)
    dev = ag.device
    ag.v.target.copy_(ag.v.store.theta * 0.9 + 0.003)
    cols = {"state:observation": rng.randn(B, D).astype(np.float32),
            "next_state:observation": rng.randn(B, D).astype(np.float32),
            "action": np.tanh(rng.randn(B, A)).astype(np.float32), "reward": rng.randn(B) * 5,
            "game_over": (rng.rand(B) < 0.1).astype(np.uint8)}
    batch = DeviceBatch({k: torch.from_numpy(v).to(dev) for k, v in cols.items()}, B)
    lr = ag.ap.network_wrappers["q"].learning_rate
    opts = None
    for step in range(2):
        pol, q, v, vt = (ag.policy.store.export_named(), ag.q.store.export_named(), ag.v.store.export_named(),
                         ag.v.store.export_named(ag.v.target))
        if opts is None:
            opts = [oac.make_adam(x, lr, 0.9, 0.99, 1e-4) for x in (pol, q, v)]
        # |eps| <= 2.5 keeps tanh away from saturation: log(1 - tanh^2 + 1e-6) loses all relative precision in fp32
        # once 1 - tanh^2 ~ 1e-6 (in the oracle as much as on the device), which would make a 1e-5 comparison of
        # log pi meaningless rather than strict.
        noise = [np.clip(rng.standard_normal((B, A)), -2.5, 2.5).astype(np.float32) for _ in range(3)]
        loss, _, _ = ag.learn_from_batch(batch, noise=noise)
        torch.cuda.synchronize()
        ref = osac.sac_step(pol, q, v, vt, opts[0], opts[1], opts[2],
                            dict(states=cols["state:observation"], next_states=cols["next_state:observation"],
                                 actions=cols["action"], rewards=cols["reward"],
                                 game_overs=cols["game_over"].astype(bool)), noise)
        opts64 = [oac.make_adam(x, lr, 0.9, 0.99, 1e-4, dtype=torch.float64) for x in (pol, q, v)]   # grads only
        ref64 = osac.sac_step(pol, q, v, vt, opts64[0], opts64[1], opts64[2],
                              dict(states=cols["state:observation"], next_states=cols["next_state:observation"],
                                   actions=cols["action"], rewards=cols["reward"],
                                   game_overs=cols["game_over"].astype(bool)), noise, dtype=torch.float64)
        close(ag.sampled.cpu().numpy(), ref["sampled"], name="sampled actions")
        # log pi contains log(1 - a^2 + 1e-6): one fp32 ulp of a^2 (6e-8) moves it by 6e-8 / (1 - a^2 + 1e-6), in
        # the fp32 oracle exactly as on the device.  That forward-error bound (x4: a, a^2, the subtraction, the log)
        # is the only slack on top of the 1e-5 rule.
        cond = (4 * 6e-8 / (1.0 - ref["sampled"].astype(np.float64) ** 2 + 1e-6)).sum(1)
        close(ag.logp.cpu().numpy(), ref["logp"], name="log pi", atol=cond)
        close(ag.log_target.cpu().numpy(), ref["log_target"], name="min Q")
        close(ag.dq_da.cpu().numpy(), ref["dq_da"], name="dq_da")
        close(ag.td_targets.cpu().numpy(), ref["td_targets"], name="td targets")
        close(loss, ref["q_loss"], name="q loss")
        close(ag.v_loss.item(), ref["v_loss"], name="v loss")
        for net, gkey, pkey in ((ag.policy, "policy_grads", "new_policy"), (ag.q, "q_grads", "new_q"),
                                (ag.v, "v_grads", "new_v")):
            got_g = net.store.export_named(net.store.grad)
            for n in ref[gkey]:
                try:
                    close(got_g[n], ref[gkey][n].numpy(), name=gkey + " " + n)
                except AssertionError as exc:
                    # batch sums of terms of either sign: two fp32 evaluations differ by their rounding noise; then
                    # the device result must be at least as close to the fp64 evaluation as the fp32 oracle is
                    w64 = ref64[gkey][n].numpy()
                    e_ours, e_orc = np.abs(got_g[n] - w64).max(), np.abs(ref[gkey][n].numpy() - w64).max()
                    assert e_ours <= 2 * e_orc, "%s; vs fp64: ours %.3e, fp32 oracle %.3e" % (exc, e_ours, e_orc)
            got_p = net.store.export_named()
            for n in ref[pkey]:
                close(got_p[n], ref[pkey][n].numpy(), name=pkey + " " + n, atol=1e-3 * lr)


def test_sac_train_driver_uniform_replay():
    B, D, A = 64, 17, 6
    ag = _make(B)
    rng = np.random.RandomState(0)
    n = 1000
    ag.memory.store_columns({"state:observation": rng.randn(n, D).astype(np.float32),
                             "next_state:observation": rng.randn(n, D).astype(np.float32),
                             "action": np.tanh(rng.randn(n, A)).astype(np.float32), "reward": rng.randn(n),
                             "game_over": (rng.rand(n) < 0.05).astype(np.uint8)})
    ag.ap.algorithm.num_consecutive_training_steps = 3
    t0 = ag.v.target.clone()
    np.random.seed(1)
    ag.total_steps_counter = 1
    loss = ag.train()
    assert np.isfinite(loss) and ag.training_iteration == 3
    assert not torch.equal(t0, ag.v.target)            # polyak step with tau = 0.005
