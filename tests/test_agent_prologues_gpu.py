"""GPU replay of tests/golden/agent_prologues.npz: the arrays the UNMODIFIED reference agents handed to their
networks / memories (oracle/make_golden_agents.py) against the CUDA kernels that compute the same quantities inside
coach_b200's agents.  Everything here goes through the C ABI (ctypes); integer / fp64-then-rounded-once quantities are
compared bit for bit.

  DQN / DDQN   cb200_dqn_td_targets           <- dqn_agent.py:92-103, ddqn_agent.py:42-43
  priorities   PrioritizedExperienceReplay    <- value_optimization_agent.py:74-80
  ClippedPPO   rl_math.fill_advantages        <- clipped_ppo_agent.py:157-207
  DDPG / TD3   cb200_ac_td_targets, cb200_min2, cb200_td3_smooth_actions  <- ddpg_agent.py:152-161, td3_agent.py:161-179
  SAC          cb200_sub, cb200_ac_td_targets <- soft_actor_critic_agent.py:234-260
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _lib():
    from coach_b200 import _lib
    return _lib, _lib.load()


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def fx():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "agent_prologues.npz"))


@pytest.mark.parametrize("tag", ["dqn", "ddqn"])
def test_dqn_targets_and_priorities_match_reference_agent(fx, tag):
    L, lib = _lib()
    from coach_b200.memories.memory import MemoryGranularity
    from coach_b200.memories.prioritized_experience_replay import PrioritizedExperienceReplay
    B, A = fx[tag + "_q_next"].shape
    qn, qo = _dev(fx[tag + "_q_next"]), _dev(fx[tag + "_q_online"])
    qs = _dev(fx[tag + "_q_select"]) if tag == "ddqn" else qn
    act, rew, done = _dev(fx[tag + "_actions"]), _dev(fx[tag + "_rewards"]), _dev(fx[tag + "_game_overs"])
    out_t = torch.empty(B, A, device="cuda")
    out_e = torch.empty(B, dtype=torch.float64, device="cuda")
    L.check(lib.cb200_dqn_td_targets(qn.data_ptr(), qs.data_ptr(), qo.data_ptr(), act.data_ptr(), rew.data_ptr(),
                                     done.data_ptr(), 0.99, B, A, out_t.data_ptr(), out_e.data_ptr(), None))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out_t.cpu().numpy(), fx[tag + "_targets"])
    np.testing.assert_array_equal(out_e.cpu().numpy(), fx[tag + "_td_errors"])
    # the same TD errors through the memory: leaves == (err + eps) ** alpha evaluated like the reference does
    mem = PrioritizedExperienceReplay((MemoryGranularity.Transitions, 1024))
    idx = fx[tag + "_prio_idx"]
    mem.update_priorities(idx, out_e)
    leaves = mem.sum_tree.cpu().numpy()[mem.power_of_2_size - 1:]
    last = {int(i): k for k, i in enumerate(idx)}
    for i, k in last.items():
        assert leaves[i] == (float(fx[tag + "_td_errors"][k]) + 1e-6) ** 0.6


def test_ppo_fill_advantages_matches_reference_agent(fx):
    from coach_b200 import rl_math
    for k in range(int(fx["ppo_cases"])):
        r, v, d = _dev(fx["ppo_rewards_%d" % k]), _dev(fx["ppo_values_%d" % k][:, 0]), _dev(fx["ppo_game_overs_%d" % k])
        adv, tgt, n_valid = rl_math.fill_advantages(r, v, d, 0.99, 0.95)
        want_adv, want_tgt = fx["ppo_adv_%d" % k], fx["ppo_vtgt_%d" % k]
        nv = int(n_valid.item())
        assert nv == len(want_adv)
        np.testing.assert_allclose(adv.cpu().numpy()[:nv], want_adv, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(tgt.cpu().numpy()[:nv], want_tgt, rtol=1e-10, atol=1e-12)


def test_actor_critic_targets_match_reference_agents(fx):
    L, lib = _lib()
    for tag in ("ddpg0", "ddpg1", "ddpg2", "td3", "sac"):
        if tag == "sac":
            q, clip, nz = _dev(fx["sac_v_next"]), None, 0
        else:
            clip = tuple(float(c) for c in fx[tag + "_clip"]) if fx[tag + "_clip"].any() else None
            nz = int(fx[tag + "_nonzero_terminal"])
            q = _dev(fx[tag + "_q1"])
            if tag == "td3":
                q2, qm = _dev(fx["td3_q2"]), torch.empty_like(q)
                L.check(lib.cb200_min2(q.data_ptr(), q2.data_ptr(), q.numel(), qm.data_ptr(), None))
                q = qm
        B = q.shape[0]
        r, d = _dev(fx[tag + "_rewards"]), _dev(fx[tag + "_game_overs"])
        out = torch.empty(B, 1, device="cuda")
        L.check(lib.cb200_ac_td_targets(r.data_ptr(), d.data_ptr(), q.data_ptr(), 1, B, 0.99, nz, int(clip is not None),
                                        clip[0] if clip else 0.0, clip[1] if clip else 0.0, out.data_ptr(), None))
        torch.cuda.synchronize()
        # the reference hands fp64 targets to a float32 placeholder: one rounding
        np.testing.assert_array_equal(out.cpu().numpy(), fx[tag + "_td_targets"].astype(np.float32))
    # TD3 target-policy smoothing: fp64 draw, fp64 sum and clips, rounded once
    a = _dev(fx["td3_next_actions"]).clone()
    noise = _dev(fx["td3_noise"])
    lo, hi = (float(x) for x in fx["td3_space"])
    L.check(lib.cb200_td3_smooth_actions(a.data_ptr(), noise.data_ptr(), a.numel(), 0.5, lo, hi, None))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(a.cpu().numpy(), fx["td3_smoothed_actions"].astype(np.float32))
    # SAC value targets = min Q(s, a~pi) - log pi(a|s), fp32
    lt, lp = _dev(fx["sac_q_min"][:, 0].copy()), _dev(fx["sac_logp"])
    vt = torch.empty_like(lt)
    L.check(lib.cb200_sub(lt.data_ptr(), lp.data_ptr(), lt.numel(), vt.data_ptr(), None))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(vt.cpu().numpy()[:, None], fx["sac_value_targets"])


def test_device_batch_matches_reference_batch(fx):
    """DeviceBatch accessors after a ring store + gather: same values, shapes and expand_dims behaviour as
    rl_coach.core_types.Batch (core_types.py:488-623)"""
    from coach_b200.core_types import Transition
    from coach_b200.memories.experience_replay import ExperienceReplay
    from coach_b200.memories.memory import MemoryGranularity
    n = len(fx["batch_rewards"])
    mem = ExperienceReplay((MemoryGranularity.Transitions, 64))
    for i in range(n):
        mem.store(Transition(state={"observation": fx["batch_states"][i]}, action=int(fx["batch_actions"][i]),
                             reward=float(fx["batch_rewards"][i]),
                             next_state={"observation": fx["batch_next_states"][i]},
                             game_over=bool(fx["batch_game_overs"][i])))
    from coach_b200.core_types import DeviceBatch
    mem._flush()
    b = DeviceBatch(dict(mem.ring.gather(torch.arange(n, device="cuda"))), n)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(b.states(["observation"])["observation"].cpu().numpy(), fx["batch_states"])
    np.testing.assert_array_equal(b.next_states(["observation"])["observation"].cpu().numpy(), fx["batch_next_states"])
    np.testing.assert_array_equal(b.actions().cpu().numpy(), fx["batch_actions"])
    np.testing.assert_array_equal(b.rewards().cpu().numpy(), fx["batch_rewards"])
    np.testing.assert_array_equal(b.game_overs().cpu().numpy().astype(bool), fx["batch_game_overs"])
    assert tuple(b.rewards(True).shape) == fx["batch_rewards_x"].shape
    assert tuple(b.actions(True).shape) == fx["batch_actions_x"].shape
    b.slice(3, 11)
    assert b.size == int(fx["batch_slice_size"])
    np.testing.assert_array_equal(b.rewards().cpu().numpy(), fx["batch_slice_rewards"])


@pytest.mark.parametrize("case", [0, 1, 2])
def test_episodic_replay_matches_reference_session(fx, case):
    """EpisodicExperienceReplay on the HBM ring replays a session recorded from the reference
    (episodic_experience_replay.py:60-330, Episode core_types.py:771-820): store / store_episode, whole-episode
    eviction, n-step discounted returns, the n-step next_state relink + should_bootstrap_next_state, counters, and the
    uniform sample stream"""
    from types import SimpleNamespace
    from coach_b200.core_types import Transition
    from coach_b200.memories.episodic_experience_replay import EpisodicExperienceReplay
    from coach_b200.memories.memory import MemoryGranularity
    p = "epi%d_" % case
    n_step, cap = int(fx[p + "n_step"]), int(fx[p + "capacity"])
    mem = EpisodicExperienceReplay((MemoryGranularity.Transitions, cap), n_step=n_step, discount=0.99)
    st, rw, ac = fx[p + "in_state"], fx[p + "in_reward"], fx[p + "in_action"]
    script = fx[p + "script"]
    k = 0
    for e, (kind, T) in enumerate(script):
        last = e == len(script) - 1                       # the final entry is an episode left open
        ts = [Transition(state={"observation": np.array([st[k + i]], dtype=np.float32)}, action=int(ac[k + i]),
                         reward=float(rw[k + i]),
                         next_state={"observation": np.array([st[k + i] + 100000], dtype=np.float32)},
                         game_over=bool(i == T - 1 and not last)) for i in range(T)]
        k += T
        if kind == 1:
            mem.store_episode(SimpleNamespace(transitions=ts))
        else:
            for t in ts:
                mem.store(t)
    counts = fx[p + "counts"]
    assert [mem.num_transitions(), mem.num_transitions_in_complete_episodes(), mem.num_complete_episodes(),
            mem.length()] == list(counts)
    b = mem.transitions_batch()
    torch.cuda.synchronize()
    np.testing.assert_array_equal(b.states(["observation"])["observation"].cpu().numpy()[:, 0], fx[p + "state"])
    np.testing.assert_array_equal(b.next_states(["observation"])["observation"].cpu().numpy()[:, 0],
                                  fx[p + "next_state"])
    np.testing.assert_array_equal(b.rewards().cpu().numpy(), fx[p + "reward"])
    np.testing.assert_array_equal(b.game_overs().cpu().numpy(), fx[p + "game_over"])
    np.testing.assert_array_equal(b.n_step_discounted_rewards().cpu().numpy(), fx[p + "nstep"])
    if n_step > 1:
        np.testing.assert_array_equal(b.info("should_bootstrap_next_state").cpu().numpy(), fx[p + "bootstrap"])
    else:
        assert "should_bootstrap_next_state" not in b.columns and np.all(fx[p + "bootstrap"] == -1)
    np.random.seed(11 + case)
    s = mem.sample_batch(32)
    np.testing.assert_array_equal(s.states(["observation"])["observation"].cpu().numpy()[:, 0], fx[p + "sample_state"])


@pytest.mark.parametrize("tag", ["c51", "rainbow"])
def test_distributional_targets_match_reference_agents(fx, tag):
    """cb200_c51_head fed the fixture's network outputs (probabilities): TD_targets bit-exact with what the reference
    agents hand to the train op; target actions = the reference's argmax; loss / gradient against a torch evaluation"""
    L, lib = _lib()
    g = lambda k: fx[tag + "_" + k]                                                              # noqa: E731
    B, A, N = g("dist_next").shape
    d_online = g("dist_online")
    logits_online = np.log(d_online).astype(np.float32)
    nxt, onl = _dev(g("dist_next")), _dev(logits_online)
    sel = _dev(g("dist_select")) if tag == "rainbow" else None
    act, rew, done, boot, z = _dev(g("actions")), _dev(g("rewards")), _dev(g("game_overs")), _dev(g("bootstrap")), \
        _dev(g("z"))
    labels = torch.empty(B, A, N, device="cuda")
    dlog = torch.empty(B, A, N, device="cuda")
    rows = torch.empty(B, A, device="cuda")
    total = torch.empty(1, device="cuda")
    td = torch.empty(B, dtype=torch.float64, device="cuda")
    q = torch.empty(B, A, dtype=torch.float64, device="cuda")
    ta = torch.empty(B, dtype=torch.int64, device="cuda")
    L.check(lib.cb200_c51_head(nxt.data_ptr(), onl.data_ptr(), sel.data_ptr() if sel is not None else None,
                               act.data_ptr(), rew.data_ptr(), done.data_ptr(),
                               boot.data_ptr() if tag == "rainbow" else None, z.data_ptr(), float(g("gamma_n")),
                               B, A, N, 1, labels.data_ptr(), dlog.data_ptr(), rows.data_ptr(), total.data_ptr(),
                               td.data_ptr(), q.data_ptr(), ta.data_ptr(), None))
    torch.cuda.synchronize()
    want = g("targets")
    got = labels.cpu().numpy()
    actions = g("actions")
    ar = np.arange(B)
    # the taken action's row is the projected distribution: bit-exact
    np.testing.assert_array_equal(got[ar, actions], want[ar, actions])
    src = g("dist_select") if tag == "rainbow" else g("dist_next")
    np.testing.assert_array_equal(ta.cpu().numpy(), np.argmax(np.dot(src, g("z")), axis=1))
    # other rows: the online softmax (the fixture's was produced by numpy, the kernel's by expf: 1e-6)
    np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-8)
    # loss rows / gradient vs torch on the kernel's own labels
    lg = torch.from_numpy(logits_online).double()
    lab = torch.from_numpy(got).double()
    ref_rows = -(lab * torch.log_softmax(lg, dim=-1)).sum(-1)
    np.testing.assert_allclose(rows.cpu().numpy(), ref_rows.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(float(total.cpu()), float(ref_rows.sum()), rtol=1e-5)
    np.testing.assert_array_equal(td.cpu().numpy(), rows.cpu().numpy()[ar, actions].astype(np.float64))
    ref_g = torch.softmax(lg, dim=-1) - lab
    gk = dlog.cpu().numpy()
    np.testing.assert_allclose(gk[ar, actions], ref_g.numpy()[ar, actions], atol=2e-7)
    mask = np.ones((B, A), dtype=bool)
    mask[ar, actions] = False
    assert (gk[mask] == 0).all()
    np.testing.assert_allclose(q.cpu().numpy(), np.dot(torch.softmax(lg, -1).numpy(), g("z")), rtol=1e-6, atol=1e-6)
