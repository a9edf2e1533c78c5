"""GPU parity of the scalar RL recurrences and filters against fixtures from the unmodified reference
(tests/golden/rl_math.npz) and the numpy oracle (oracle/rl_math.py)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import rl_math as orm        # noqa: E402  (checker only)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_gae_matches_reference_fixture(golden_dir):
    from coach_b200 import rl_math
    fx = np.load(os.path.join(golden_dir, "rl_math.npz"))
    for k in range(len(fx["gae_lens"])):
        r, v = fx["gae_r_%d" % k], fx["gae_v_%d" % k]
        T = len(r)
        done = np.zeros(T, np.uint8)
        done[-1] = 1                      # single complete episode; the fixture's bootstrap value v[-1] is 0
        adv, tgt, nv = rl_math.gae(dev(r), dev(v[:-1].astype(np.float32)), dev(done), 0.99, 0.95)
        assert int(nv.item()) == T
        np.testing.assert_allclose(adv.cpu().numpy(), fx["gae_adv_%d" % k], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(tgt.cpu().numpy(), fx["gae_tgt_%d" % k][:, 0], rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("n,p_done", [(131072, 1 / 500.0), (2048, 0.01), (5000, 0.3), (7, 0.5), (1, 1.0)])
def test_fill_advantages_matches_oracle(n, p_done):
    """BASELINE config 3 shape (64 envs x 2048 steps = 131072 transitions) and edge cases: many short episodes, a
    trailing unfinished segment, a single transition."""
    from coach_b200 import rl_math
    rng = np.random.RandomState(n)
    r = rng.randn(n)
    v = rng.randn(n).astype(np.float32)
    done = (rng.rand(n) < p_done)
    if n > 10:
        done[-3:] = False                 # trailing segment without game_over -> gets nothing
        done[n // 2] = True
    else:
        done[-1] = True
    adv, tgt, nv = rl_math.fill_advantages(dev(r), dev(v), dev(done.astype(np.uint8)), 0.99, 0.95)
    o_adv, o_tgt, o_nv = orm.ppo_fill_advantages(r, v, done, 0.99, 0.95)
    assert int(nv.item()) == o_nv
    np.testing.assert_allclose(adv.cpu().numpy()[:o_nv], o_adv[:o_nv], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(tgt.cpu().numpy()[:o_nv], o_tgt[:o_nv], rtol=1e-9, atol=1e-11)
    assert np.all(np.isnan(adv.cpu().numpy()[o_nv:]))


def test_nstep_returns_bit_exact(golden_dir):
    from coach_b200 import rl_math
    fx = np.load(os.path.join(golden_dir, "rl_math.npz"))
    for k in range(int(fx["nstep_cases"])):
        r = fx["nstep_r_%d" % k]
        out = rl_math.nstep_returns(dev(r), [len(r)], 0.99, int(fx["nstep_n_%d" % k]))
        np.testing.assert_array_equal(out.cpu().numpy(), fx["nstep_out_%d" % k])
    # several episodes back to back
    rng = np.random.RandomState(0)
    lens = [5, 1, 300, 17]
    r = rng.randn(sum(lens))
    out = rl_math.nstep_returns(dev(r), lens, 0.97, 4).cpu().numpy()
    o = 0
    for L in lens:
        np.testing.assert_array_equal(out[o:o + L], orm.n_step_returns(r[o:o + L], 0.97, 4))
        o += L


def test_observation_normalization_filter(golden_dir):
    from coach_b200.filters.filter import ObservationNormalizationFilter
    fx = np.load(os.path.join(golden_dir, "rl_math.npz"))
    f = ObservationNormalizationFilter()
    f.set_device("cuda")
    f.set_shape([17])
    for k in range(3):
        f.filter(dev(fx["rs_push%d" % k]), update_internal_state=True)
        np.testing.assert_allclose(f.running_observation_stats.mean.cpu().numpy(), fx["rs_means"][k], rtol=1e-12,
                                   atol=1e-13)
        np.testing.assert_allclose(f.running_observation_stats.std.cpu().numpy(), fx["rs_stds"][k], rtol=1e-11)
    assert f.running_observation_stats.n == float(fx["rs_count"])
    out = f.filter(dev(fx["rs_query"]), update_internal_state=False)
    np.testing.assert_allclose(out.cpu().numpy(), fx["rs_norm"].astype(np.float32), rtol=1e-6, atol=1e-7)
    assert out.cpu().numpy().max() <= 5.0 and out.cpu().numpy().min() >= -5.0
    # PPO-sized push (131072 x 17) against the numpy oracle
    rng = np.random.RandomState(1)
    big = (rng.randn(131072, 17) * 2 + 0.5).astype(np.float32)
    rs = orm.RunningStats([17])
    for k in range(3):
        rs.push(fx["rs_push%d" % k])
    rs.push(big)
    f.filter(dev(big), update_internal_state=True)
    np.testing.assert_allclose(f.running_observation_stats.mean.cpu().numpy(), rs.mean, rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(f.running_observation_stats.std.cpu().numpy(), rs.std, rtol=1e-10)
