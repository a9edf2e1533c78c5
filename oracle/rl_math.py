"""numpy fp64 restatements of the scalar RL recurrences on the path.  TEST INFRASTRUCTURE ONLY.

  gae()                <- rl_coach/agents/actor_critic_agent.py:108-125 (lfilter([1],[1,-g*l]) on reversed deltas)
  n_step_returns()     <- rl_coach/core_types.py:771-801
  RunningStats         <- rl_coach/utilities/shared_running_stats.py:115-164
  ppo_fill_advantages()<- rl_coach/agents/clipped_ppo_agent.py:157-207 (episode split, zero bootstrap, standardise)
  ac_td_targets()      <- rl_coach/agents/ddpg_agent.py:152-161, td3_agent.py:170-179, soft_actor_critic_agent.py:259-260
  td3_smooth_actions() <- rl_coach/agents/td3_agent.py:161-164 (+ spaces.py BoxActionSpace.clip_action_to_space)

Pinned: every function here reproduces tests/golden/{rl_math,agent_prologues}.npz, which the UNMODIFIED reference
generated (oracle/make_golden.py, oracle/make_golden_agents.py; tests/test_oracle_golden.py).
"""
import numpy as np


def discount_reverse(x, g):
    """y_t = x_t + g*y_{t+1}; the operation order of scipy's lfilter direct-form-II-transposed on the reversed
    sequence: y[n] = x[n] + g*y[n-1] evaluated as (g*y_prev) + x ... lfilter computes y = b0*x + z; z = -a1*y,
    i.e. y_t = x_t + (g*y_{t+1}) -- one multiply then one add, fp64."""
    x = np.asarray(x, dtype=np.float64)
    y = np.empty_like(x)
    acc = 0.0
    for t in range(len(x) - 1, -1, -1):
        acc = x[t] + g * acc
        y[t] = acc
    return y


def gae(rewards, values, discount, lam):
    """values has T+1 entries.  Returns (advantages[T], value_targets[T]) with value target = A + V[:-1]
    (estimate_state_value_using_gae=True, actor_critic_agent.py:120-121)."""
    rewards = np.asarray(rewards, dtype=np.float64)
    values = np.asarray(values, dtype=np.float64)
    deltas = rewards + discount * values[1:] - values[:-1]
    adv = discount_reverse(deltas, discount * lam)
    return adv, adv + values[:-1]


def n_step_returns(rewards, discount, n_step):
    """core_types.py:779-790: R_t = sum_{k<n} g^k r_{t+k}, accumulated in the reference's order
    (k ascending, running power of the discount)."""
    r = np.asarray(rewards, dtype=np.float64)
    T = len(r)
    n = T if (n_step == -1 or n_step > T) else n_step
    out = r.copy()
    cur = discount
    for i in range(1, n):
        shifted = np.zeros(T)
        shifted[:T - i] = r[i:]
        out += cur * shifted
        cur *= discount
    return out


class RunningStats:
    """NumpySharedRunningStats (shared_running_stats.py:115-164)."""

    def __init__(self, shape, epsilon=1e-2, clip_values=(-5.0, 5.0)):
        self.epsilon = epsilon
        self.count = epsilon
        self.sum = np.zeros(shape)
        self.sum_squares = epsilon * np.ones(shape)
        self.mean = np.zeros(shape)
        self.std = np.sqrt(epsilon) * np.ones(shape)
        self.clip_values = clip_values

    def push(self, samples):
        s = np.asarray(samples).astype(np.float64)
        self.sum += s.sum(axis=0)
        self.sum_squares += np.square(s).sum(axis=0)
        self.count += s.shape[0]
        self.mean = self.sum / self.count
        self.std = np.sqrt(np.maximum(
            (self.sum_squares - self.count * np.square(self.mean)) / np.maximum(self.count - 1, 1), self.epsilon))

    def normalize(self, batch):
        return np.clip((batch - self.mean) / (self.std + 1e-15), *self.clip_values)


def ppo_fill_advantages(rewards, values, game_overs, discount, lam):
    """clipped_ppo_agent.py:170-207 for policy_gradient_rescaler == GAE.

    rewards fp64[N], values fp32/64[N] (V(s_t) from the online net), game_overs bool[N].  Episodes are split at
    game_over flags, the bootstrap value appended at every episode end is 0 (:188), a trailing segment without
    game_over gets nothing (the reference's zip() truncates, :203) -- here its entries are returned as NaN and
    ``n_valid`` tells how many leading transitions were filled.  Advantages are standardised with the population
    std over all filled entries (:201)."""
    rewards = np.asarray(rewards, dtype=np.float64)
    values = np.asarray(values, dtype=np.float64)
    N = len(rewards)
    adv = np.full(N, np.nan)
    tgt = np.full(N, np.nan)
    start = 0
    n_valid = 0
    for i in range(N):
        if game_overs[i]:
            v = np.append(values[start:i + 1], 0.0)
            a, t = gae(rewards[start:i + 1], v, discount, lam)
            adv[start:i + 1] = a
            tgt[start:i + 1] = t
            start = i + 1
            n_valid = i + 1
    a = adv[:n_valid]
    adv[:n_valid] = (a - np.mean(a)) / np.std(a)
    return adv, tgt, n_valid


def ac_td_targets(rewards, game_overs, q_next, discount, clip=None, use_non_zero_discount_for_terminal_states=False):
    """Bootstrapped TD targets of the actor-critic agents: fp64 numpy broadcasting on the fp32 network output,
    [B, 1].  The train op's float32 placeholder rounds the result once (np.float32(result))."""
    r = np.asarray(rewards, dtype=np.float64).reshape(-1, 1)
    d = np.asarray(game_overs).astype(bool).reshape(-1, 1)
    q = np.asarray(q_next).reshape(-1, 1)
    if use_non_zero_discount_for_terminal_states:
        y = r + discount * q
    else:
        y = r + (1.0 - d) * discount * q
    if clip:
        y = np.clip(y, *clip)
    return y


def td3_smooth_actions(next_actions, noise, noise_clip, low, high):
    """td3_agent.py:162-164: fp64 draw clipped, added to the fp32 target-actor output (promoted to fp64), clipped to
    the action space; fp64 result (the critic's float32 placeholder rounds it once)."""
    nz = np.asarray(noise, dtype=np.float64).clip(-noise_clip, noise_clip)
    return np.clip(np.asarray(next_actions) + nz, low, high)
