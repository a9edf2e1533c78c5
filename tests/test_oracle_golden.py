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
