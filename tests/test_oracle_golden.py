"""The oracle restatement (both backends) replayed against fixtures produced by the unmodified reference
(oracle/make_golden.py) and against the reference's own SegmentTree known-answer tests
(rl_coach/tests/memories/test_prioritized_experience_replay.py:12-87).  CPU only."""
import os

import numpy as np
import pytest

from oracle import memory as om
from oracle import rl_math


@pytest.mark.parametrize("backend", ["python", "c"])
def test_segment_tree_known_answers(backend):
    # values from the reference's test_sum_tree / test_min_tree / test_max_tree
    st = om.OracleSegmentTree(4, "sum", backend)
    for v, tot in [(10, 10), (20, 30), (5, 35), (7.5, 42.5), (2.5, 35), (5, 20)]:
        st.add(v)
        assert st.total_value() == tot
    assert st.retrieve(2) == (0, 2.5)
    assert st.retrieve(3) == (1, 5.0)
    assert st.retrieve(10) == (2, 5.0)
    assert st.retrieve(13) == (3, 7.5)
    st.update(2, 10)
    assert st.levels_str() == "[25.]\n[ 7.5 17.5]\n[ 2.5  5.  10.   7.5]\n"
    with pytest.raises(ValueError):
        om.OracleSegmentTree(5, "sum", backend)

    mn = om.OracleSegmentTree(4, "min", backend)
    for v, tot in [(10, 10), (20, 10), (5, 5), (7.5, 5), (2, 2)]:
        mn.add(v)
        assert mn.total_value() == tot
    for v in (3, 3, 3, 5):
        mn.add(v)
    assert mn.total_value() == 3

    mx = om.OracleSegmentTree(4, "max", backend)
    for v, tot in [(10, 10), (20, 20), (5, 20), (7.5, 20), (2, 20)]:
        mx.add(v)
        assert mx.total_value() == tot
    for v in (3, 3, 3, 5):
        mx.add(v)
    assert mx.total_value() == 5
    mx.update(1, 10)
    assert mx.total_value() == 10
    assert mx.levels_str() == "[10.]\n[10.  3.]\n[ 5. 10.  3.  3.]\n"
    mx.update(1, 2)
    assert mx.levels_str() == "[5.]\n[5. 3.]\n[5. 2. 3. 3.]\n"


def replay_per_fixture(fx, mem, sample_fn, update_fn, store_fn, snapshot_fn):
    """Drives any PER implementation through a recorded reference session; shared with the GPU tests."""
    si = ui = ci = 0
    for step, op in enumerate(fx["ops"]):
        if op == 0:
            store_fn(mem, int(fx["store_counts"][ci]))
            ci += 1
        elif op == 1:
            idx, w = sample_fn(mem, fx["uniforms"][si], float(fx["s_beta"][si]), int(fx["s_nt"][si]))
            np.testing.assert_array_equal(np.asarray(idx), fx["s_idx"][si], err_msg="indices, sample %d" % si)
            yield ("weights", np.asarray(w), fx["s_w"][si])
            si += 1
        else:
            update_fn(mem, fx["u_idx"][ui], fx["u_err"][ui])
            ui += 1
        s, mn, mx, maxp = snapshot_fn(mem)
        yield ("tree", (s, mn, mx, maxp),
               (fx["snaps_sum"][step], fx["snaps_min"][step], fx["snaps_max"][step], fx["snap_maxp"][step]))


@pytest.mark.parametrize("backend", ["python", "c"])
@pytest.mark.parametrize("name", ["per_small", "per_nonpow2", "per_medium"])
def test_per_session_matches_reference(golden_dir, name, backend):
    fx = np.load(os.path.join(golden_dir, name + ".npz"))
    beta = om.OracleLinearSchedule(*fx["beta"])
    mem = om.OraclePrioritizedExperienceReplay(int(fx["max_size"]), alpha=float(fx["alpha"]), beta=beta,
                                               epsilon=float(fx["epsilon"]), backend=backend)
    assert mem.power_of_2_size == int(fx["size"])

    def sample_fn(m, u, beta_expected, nt_expected):
        assert float(m.beta.current_value) == beta_expected      # LinearSchedule recurrence, bit-exact
        assert m.num_transitions() == nt_expected                 # doubled count (quirk Q1)
        return m.sample_indices(len(u), uniforms=list(u))

    def store_fn(m, n):
        if backend == "c":
            m.store_many([None] * n)
        else:
            for _ in range(n):
                m.store(None)

    for kind, got, want in replay_per_fixture(
            fx, mem, sample_fn, lambda m, i, e: m.update_priorities(list(i), list(e)), store_fn,
            lambda m: (m.sum_tree.tree, m.min_tree.tree, m.max_tree.tree, m.maximal_priority)):
        if kind == "weights":
            np.testing.assert_array_equal(got, want)              # same libm -> bit-exact
        else:
            for g, w in zip(got[:3], want[:3]):
                np.testing.assert_array_equal(g, w)
            assert got[3] == want[3]


def test_er_uniform_indices(golden_dir):
    fx = np.load(os.path.join(golden_dir, "er_uniform.npz"))
    for dup in (1, 0):
        mem = om.OracleExperienceReplay(int(fx["max_size"]), bool(dup))
        for i in range(int(fx["n_store"])):
            mem.store(i)
        assert mem.num_transitions() == int(fx["num_transitions"])
        np.random.seed(int(fx["seed"]))
        for row in fx["idx_dup%d" % dup]:
            np.testing.assert_array_equal(mem.sample_indices(int(fx["batch"])), row)


def test_linear_schedule(golden_dir):
    fx = np.load(os.path.join(golden_dir, "linear_schedule.npz"))
    s = om.OracleLinearSchedule(0.4, 1.0, 1000)
    for v in fx["vals"]:
        assert float(s.current_value) == v
        s.step()


def test_gae_nstep_runningstats(golden_dir):
    fx = np.load(os.path.join(golden_dir, "rl_math.npz"))
    for k in range(len(fx["gae_lens"])):
        adv, tgt = rl_math.gae(fx["gae_r_%d" % k], fx["gae_v_%d" % k], 0.99, 0.95)
        np.testing.assert_allclose(adv, fx["gae_adv_%d" % k], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(tgt, fx["gae_tgt_%d" % k][:, 0], rtol=1e-12, atol=1e-12)
    for k in range(int(fx["nstep_cases"])):
        out = rl_math.n_step_returns(fx["nstep_r_%d" % k], 0.99, int(fx["nstep_n_%d" % k]))
        np.testing.assert_array_equal(out, fx["nstep_out_%d" % k])
    rs = rl_math.RunningStats([17])
    for k in range(3):
        rs.push(fx["rs_push%d" % k])
        np.testing.assert_array_equal(rs.mean, fx["rs_means"][k])
        np.testing.assert_array_equal(rs.std, fx["rs_stds"][k])
    assert rs.count == float(fx["rs_count"])
    np.testing.assert_array_equal(rs.normalize(fx["rs_query"]), fx["rs_norm"])


# ---- agent prologues (oracle/make_golden_agents.py: the reference agents' own learn_from_batch, stub networks) -------
def test_dqn_ddqn_targets_match_reference_agents(golden_dir):
    from oracle import nets as on
    fx = np.load(os.path.join(golden_dir, "agent_prologues.npz"))
    for tag in ("dqn", "ddqn"):
        q_sel = fx[tag + "_q_select"] if tag == "ddqn" else fx[tag + "_q_next"]     # ddqn_agent.py:42-43
        targets, td = on.dqn_targets(fx[tag + "_q_next"], q_sel, fx[tag + "_q_online"], fx[tag + "_actions"],
                                     fx[tag + "_rewards"], fx[tag + "_game_overs"].astype(bool), 0.99)
        np.testing.assert_array_equal(targets, fx[tag + "_targets"])           # what train_and_sync_networks was fed
        np.testing.assert_array_equal(td, fx[tag + "_td_errors"])              # what update_priorities was handed
        assert targets.dtype == np.float32 and fx[tag + "_weights"] is not None


def test_ppo_fill_advantages_matches_reference_agent(golden_dir):
    fx = np.load(os.path.join(golden_dir, "agent_prologues.npz"))
    for k in range(int(fx["ppo_cases"])):
        adv, tgt, n_valid = rl_math.ppo_fill_advantages(fx["ppo_rewards_%d" % k], fx["ppo_values_%d" % k][:, 0],
                                                        fx["ppo_game_overs_%d" % k].astype(bool), 0.99, 0.95)
        want_adv, want_tgt = fx["ppo_adv_%d" % k], fx["ppo_vtgt_%d" % k]
        assert n_valid == len(want_adv)                        # trailing open episode gets nothing (zip(), :203)
        np.testing.assert_allclose(adv[:n_valid], want_adv, rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(tgt[:n_valid], want_tgt, rtol=1e-12, atol=1e-13)
        assert np.all(np.isnan(adv[n_valid:]))


def test_actor_critic_targets_match_reference_agents(golden_dir):
    fx = np.load(os.path.join(golden_dir, "agent_prologues.npz"))
    for tag in ("ddpg0", "ddpg1", "ddpg2", "td3"):
        clip = tuple(fx[tag + "_clip"]) if fx[tag + "_clip"].any() else None
        q = np.minimum(fx[tag + "_q1"], fx[tag + "_q2"]) if tag == "td3" else fx[tag + "_q1"]
        y = rl_math.ac_td_targets(fx[tag + "_rewards"], fx[tag + "_game_overs"], q, 0.99, clip,
                                  bool(fx[tag + "_nonzero_terminal"]))
        np.testing.assert_array_equal(y, fx[tag + "_td_targets"])
        # the actor is fed -dQ/da (ddpg_agent.py:181, td3_agent.py:196); the critic trains on the batch's own actions
        np.testing.assert_array_equal(fx[tag + "_actor_feed"], -fx[tag + "_action_grads"])
        np.testing.assert_array_equal(fx[tag + "_critic_train_action"], fx[tag + "_batch_actions"])
    # DDPG takes dQ/da from critic output 1, TD3 from output 3 (mean of Q1)
    assert str(fx["ddpg0_grad_fetch"][0]) == "g1" and str(fx["td3_grad_fetch"][0]) == "g3"
    sm = rl_math.td3_smooth_actions(fx["td3_next_actions"], fx["td3_noise"], 0.5, *fx["td3_space"])
    np.testing.assert_array_equal(sm, fx["td3_smoothed_actions"])
    y = rl_math.ac_td_targets(fx["sac_rewards"], fx["sac_game_overs"], fx["sac_v_next"], 0.99)
    np.testing.assert_array_equal(y, fx["sac_td_targets"])
    np.testing.assert_array_equal((fx["sac_q_min"][:, 0] - fx["sac_logp"])[:, None], fx["sac_value_targets"])
    assert float(fx["sac_dlogp_feed"]) == 1.0
    np.testing.assert_array_equal(fx["sac_dq_feed"], fx["sac_dq_da"])
    for i in range(2):       # policy gradient = d(mean log pi) - dq_dphi (soft_actor_critic_agent.py:231)
        np.testing.assert_array_equal(fx["sac_pgrad_%d" % i], fx["sac_dlogp_%d" % i] - fx["sac_dq_%d" % i])
    np.testing.assert_array_equal(fx["sac_q_action_input"], fx["sac_sampled_actions"])


def test_batch_matches_reference_batch(golden_dir):
    """core_types.py:405-649: AoS -> SoA, expand_dims, info, slice -- oracle restatement and the host Batch class"""
    from coach_b200.core_types import Batch, Transition
    fx = np.load(os.path.join(golden_dir, "agent_prologues.npz"))
    n = len(fx["batch_rewards"])
    ts = [Transition(state={"observation": fx["batch_states"][i]}, action=int(fx["batch_actions"][i]),
                     reward=float(fx["batch_rewards"][i]), next_state={"observation": fx["batch_next_states"][i]},
                     game_over=bool(fx["batch_game_overs"][i]),
                     info={"idx": int(fx["batch_idx"][i]), "weight": float(fx["batch_weight"][i])}) for i in range(n)]
    for i, t in enumerate(ts):
        t.n_step_discounted_rewards = float(fx["batch_nstep"][i])
    s, s2, a, r, d = om.batch_columns(ts)
    b = Batch(ts)
    for got, got2, key in ((s, b.states(["observation"])["observation"], "batch_states"),
                           (s2, b.next_states(["observation"])["observation"], "batch_next_states"),
                           (a, b.actions(), "batch_actions"), (r, b.rewards(), "batch_rewards"),
                           (d, b.game_overs(), "batch_game_overs")):
        for g in (got, got2):
            assert g.dtype == fx[key].dtype and g.shape == fx[key].shape
            np.testing.assert_array_equal(g, fx[key])
    for key, got in (("batch_actions_x", b.actions(True)), ("batch_rewards_x", b.rewards(True)),
                     ("batch_game_overs_x", b.game_overs(True)), ("batch_nstep", b.n_step_discounted_rewards()),
                     ("batch_idx", b.info("idx")), ("batch_weight", b.info("weight"))):
        assert got.shape == fx[key].shape
        np.testing.assert_array_equal(got, fx[key])
    b.slice(3, 11)
    assert b.size == int(fx["batch_slice_size"])
    np.testing.assert_array_equal(b.rewards(), fx["batch_slice_rewards"])
    np.testing.assert_array_equal(b.states(["observation"])["observation"], fx["batch_slice_states"])


@pytest.mark.parametrize("tag", ["c51", "rainbow"])
def test_distributional_targets_match_reference_agents(golden_dir, tag):
    """oracle/c51.py c51_targets == what CategoricalDQNAgent / RainbowDQNAgent.learn_from_batch hand to the train op
    (categorical_dqn_agent.py:120-152, rainbow_dqn_agent.py:107-131), bit for bit, and the PER errors (:160-163)"""
    from oracle import c51
    fx = np.load(os.path.join(golden_dir, "agent_prologues.npz"))
    g = lambda k: fx[tag + "_" + k]                                                              # noqa: E731
    sel = g("dist_select") if tag == "rainbow" else None
    targets, _, m = c51.c51_targets(g("dist_next"), g("dist_online"), sel, g("actions"), g("rewards"), g("bootstrap"),
                                    float(g("gamma_n")), g("z"))
    np.testing.assert_array_equal(targets, g("targets"))
    assert targets.dtype == np.float32
    # integral b_j loses its mass in the reference's arithmetic: at least one row of m does not sum to 1
    assert (np.abs(m.sum(axis=1) - 1.0) > 1e-3).any()
    B = targets.shape[0]
    np.testing.assert_array_equal(g("loss_rows")[np.arange(B), g("actions")].astype(np.float64), g("prio_errors"))
    np.testing.assert_array_equal(c51.z_values(-10.0, 10.0, 51), g("z"))
