"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported through oracle/ref_loader.py).

Run in the build container only:   python -m oracle.make_golden
The fixtures are committed; the GPU box (which has no /root/reference) replays them against the oracle restatement
and against the CUDA path.  TEST INFRASTRUCTURE ONLY.
"""
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _ref():
    from oracle import ref_loader
    ref_loader.load()


def golden_per(name, max_size, n_ops, batch, seed, alpha=0.6, beta=(0.4, 1.0, 50), obs_dim=6):
    """A scripted store / sample / update_priorities session on the reference PER.  Everything the session
    consumed (uniform draws, error values) and produced (indices, weights, tree snapshots) is recorded."""
    _ref()
    from rl_coach.memories.non_episodic.prioritized_experience_replay import PrioritizedExperienceReplay
    from rl_coach.memories.memory import MemoryGranularity
    from rl_coach.core_types import Transition
    from rl_coach.schedules import LinearSchedule

    rng = np.random.RandomState(seed)
    mem = PrioritizedExperienceReplay((MemoryGranularity.Transitions, max_size), alpha=alpha,
                                      beta=LinearSchedule(*beta), epsilon=1e-6)
    size = mem.power_of_2_size
    ops = []            # op codes: 0 store(n), 1 sample, 2 update
    store_counts, obs_log = [], []
    uniforms, s_idx, s_w, s_beta, s_nt = [], [], [], [], []
    u_idx, u_err = [], []
    snaps_sum, snaps_min, snaps_max, snap_maxp = [], [], [], []
    stored = 0
    last_idx = None

    def snapshot():
        snaps_sum.append(mem.sum_tree.tree.copy())
        snaps_min.append(mem.min_tree.tree.copy())
        snaps_max.append(mem.max_tree.tree.copy())
        snap_maxp.append(mem.maximal_priority)

    for op_i in range(n_ops):
        r = rng.rand()
        can_sample = mem.num_transitions() >= batch
        if not can_sample or r < 0.34:
            n = int(rng.randint(1, max(2, size // 3)))
            for _ in range(n):
                obs = rng.randint(0, 256, size=obs_dim).astype(np.uint8)
                nobs = rng.randint(0, 256, size=obs_dim).astype(np.uint8)
                t = Transition(state={'observation': obs}, action=int(rng.randint(0, 4)),
                               reward=float(rng.randint(-1, 2)), next_state={'observation': nobs},
                               game_over=bool(rng.rand() < 0.1))
                obs_log.append(np.concatenate([obs, nobs, [t.action, int(t.reward) + 1, int(t.game_over)]]))
                mem.store(t)
                stored += 1
            ops.append(0)
            store_counts.append(n)
        elif r < 0.67 or last_idx is None:
            # sample: intercept the raw random() draws that random.uniform consumes
            state = random.getstate()
            us = [random.random() for _ in range(batch)]
            random.setstate(state)
            s_beta.append(float(mem.beta.current_value))
            s_nt.append(mem.num_transitions())
            b = mem.sample(batch)
            uniforms.append(us)
            s_idx.append([t.info['idx'] for t in b])
            s_w.append([t.info['weight'] for t in b])
            last_idx = s_idx[-1]
            ops.append(1)
        else:
            # update with |N(0,1)| errors; sprinkle exact duplicates and zeros
            idx = list(last_idx)
            if rng.rand() < 0.5:
                idx[1] = idx[0]
                idx[-1] = idx[0]
            err = np.abs(rng.randn(batch))
            err[rng.rand(batch) < 0.05] = 0.0
            mem.update_priorities(idx, list(err))
            u_idx.append(idx)
            u_err.append(err)
            ops.append(2)
        snapshot()

    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        max_size=max_size, size=size, batch=batch, alpha=alpha, beta=np.array(beta, dtype=np.float64), epsilon=1e-6,
        ops=np.array(ops, dtype=np.int64), store_counts=np.array(store_counts, dtype=np.int64),
        obs_log=np.array(obs_log, dtype=np.int64),
        uniforms=np.array(uniforms, dtype=np.float64).reshape(-1, batch),
        s_idx=np.array(s_idx, dtype=np.int64).reshape(-1, batch),
        s_w=np.array(s_w, dtype=np.float64).reshape(-1, batch),
        s_beta=np.array(s_beta, dtype=np.float64), s_nt=np.array(s_nt, dtype=np.int64),
        u_idx=np.array(u_idx, dtype=np.int64).reshape(-1, batch),
        u_err=np.array(u_err, dtype=np.float64).reshape(-1, batch),
        snaps_sum=np.array(snaps_sum), snaps_min=np.array(snaps_min), snaps_max=np.array(snaps_max),
        snap_maxp=np.array(snap_maxp, dtype=np.float64))
    print(name, "ops", len(ops), "stores", stored, "samples", len(s_idx), "updates", len(u_idx))


def golden_er(name, max_size, n_store, batch, n_samples, seed):
    """Uniform ExperienceReplay.sample index streams (legacy np.random global state), with and without dups."""
    _ref()
    from rl_coach.memories.non_episodic.experience_replay import ExperienceReplay
    from rl_coach.memories.memory import MemoryGranularity
    from rl_coach.core_types import Transition
    out = {}
    for dup in (True, False):
        mem = ExperienceReplay((MemoryGranularity.Transitions, max_size), allow_duplicates_in_batch_sampling=dup)
        for i in range(n_store):
            mem.store(Transition(state={'observation': np.array([i])}, action=0, reward=0.0,
                                 next_state={'observation': np.array([i + 1])}, game_over=False))
        np.random.seed(seed)
        got = []
        for _ in range(n_samples):
            b = mem.sample(batch)
            got.append([int(t.state['observation'][0]) for t in b])
        # stored ids are i = n_store-len .. ; position p in the list holds id (n_store - len + p)
        first_id = n_store - mem.num_transitions()
        out["idx_dup%d" % int(dup)] = np.array(got, dtype=np.int64) - first_id
        out["num_transitions"] = mem.num_transitions()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), max_size=max_size, n_store=n_store, batch=batch,
                        seed=seed, **out)
    print(name, "ok")


def golden_schedule(name):
    _ref()
    from rl_coach.schedules import LinearSchedule
    s = LinearSchedule(0.4, 1.0, 1000)
    vals = []
    for _ in range(1200):
        vals.append(float(s.current_value))
        s.step()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), vals=np.array(vals))
    print(name, "ok")


def golden_rl_math(name, seed=5):
    """GAE (actor_critic_agent.py:111-125), n-step returns (core_types.py:771-801), running stats
    (utilities/shared_running_stats.py:115-164) from the reference's own code."""
    _ref()
    from types import SimpleNamespace
    from rl_coach.agents.actor_critic_agent import ActorCriticAgent
    from rl_coach.core_types import Episode, Transition
    from rl_coach.utilities.shared_running_stats import NumpySharedRunningStats
    rng = np.random.RandomState(seed)
    out = {}
    # GAE on episodes of assorted lengths
    fake = SimpleNamespace(ap=SimpleNamespace(algorithm=SimpleNamespace(discount=0.99, gae_lambda=0.95,
                                                                 estimate_state_value_using_gae=True)))
    fake.discount = lambda x, gamma: ActorCriticAgent.discount(fake, x, gamma)
    lens = [1, 2, 7, 64, 333, 2048]
    for k, T in enumerate(lens):
        r = rng.randn(T)
        v = rng.randn(T + 1).astype(np.float32)
        v[-1] = 0.0
        adv, tgt = ActorCriticAgent.get_general_advantage_estimation_values(fake, r, v.astype(np.float64) * 1.0)
        out["gae_r_%d" % k] = r
        out["gae_v_%d" % k] = v
        out["gae_adv_%d" % k] = adv
        out["gae_tgt_%d" % k] = tgt
    out["gae_lens"] = np.array(lens)
    # n-step discounted rewards
    for k, (T, n) in enumerate([(10, -1), (10, 3), (57, 5), (200, -1), (33, 1)]):
        ep = Episode(discount=0.99, n_step=n, bootstrap_total_return_from_old_policy=False)
        rew = rng.randn(T)
        for t in range(T):
            ep.insert(Transition(state={'observation': np.zeros(1)}, action=0, reward=float(rew[t]),
                                 next_state={'observation': np.zeros(1)}, game_over=(t == T - 1)))
        ep.update_discounted_rewards()
        out["nstep_r_%d" % k] = rew
        out["nstep_n_%d" % k] = n
        out["nstep_out_%d" % k] = np.array([t.n_step_discounted_rewards for t in ep.transitions])
    out["nstep_cases"] = 5
    # running stats: three pushes, then normalize
    st = NumpySharedRunningStats(name="x", epsilon=1e-2)
    st.set_params(shape=[17], clip_values=(-5.0, 5.0))
    pushes = [rng.randn(n, 17).astype(np.float32) * 3 + 1 for n in (1, 64, 1000)]
    means, stds = [], []
    for p in pushes:
        st.push(p)
        means.append(st.mean.copy())
        stds.append(st.std.copy())
    q = (rng.randn(32, 17) * 10).astype(np.float32)
    out["rs_push0"], out["rs_push1"], out["rs_push2"] = pushes
    out["rs_means"] = np.array(means)
    out["rs_stds"] = np.array(stds)
    out["rs_query"] = q
    out["rs_norm"] = st.normalize(q)
    out["rs_count"] = st._count
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "ok")


def main():
    os.makedirs(OUT, exist_ok=True)
    random.seed(1234)
    golden_per("per_small", max_size=16, n_ops=60, batch=4, seed=1)
    golden_per("per_nonpow2", max_size=100, n_ops=80, batch=16, seed=2)          # rounds up to 128
    golden_per("per_medium", max_size=1024, n_ops=60, batch=64, seed=3, obs_dim=16)
    golden_er("er_uniform", max_size=500, n_store=800, batch=32, n_samples=20, seed=7)
    golden_schedule("linear_schedule")
    golden_rl_math("rl_math")


if __name__ == "__main__":
    main()
