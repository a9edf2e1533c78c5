"""Device wrappers of the scalar RL recurrences on the path (kernels in csrc/rl_math.cu).

  gae / fill_advantages   <- agents/actor_critic_agent.py:108-125, agents/clipped_ppo_agent.py:157-207
  nstep_returns           <- core_types.py:771-801 (Episode.update_discounted_rewards)
"""
import numpy as np
import torch

from coach_b200 import _lib


def gae(rewards, values, game_overs, discount, gae_lambda):
    """rewards fp64 [N], values fp32 [N] (V(s_t)), game_overs uint8 [N] -- CUDA tensors, episodes back to back.
    Returns (advantages fp64 [N], value_targets fp64 [N], n_valid int64 [1]); only the first n_valid entries (up to
    the last game_over) are meaningful, exactly the transitions the reference fills."""
    lib = _lib.load()
    n = rewards.shape[0]
    adv = torch.empty(n, dtype=torch.float64, device=rewards.device)
    tgt = torch.empty_like(adv)
    n_valid = torch.zeros(1, dtype=torch.int64, device=rewards.device)
    _lib.check(lib.cb200_gae_scan(rewards.data_ptr(), values.data_ptr(), game_overs.data_ptr(), n, float(discount),
                                  float(gae_lambda), adv.data_ptr(), tgt.data_ptr(), n_valid.data_ptr(),
                                  _lib.current_stream()))
    return adv, tgt, n_valid


def standardize_(x, n_valid=None):
    """in place: x[:n_valid] = (x - mean) / population std; returns device tensor [mean, std]"""
    lib = _lib.load()
    ms = torch.empty(2, dtype=torch.float64, device=x.device)
    _lib.check(lib.cb200_standardize(x.data_ptr(), x.shape[0], n_valid.data_ptr() if n_valid is not None else None,
                                     ms.data_ptr(), _lib.current_stream()))
    return ms


def fill_advantages(rewards, values, game_overs, discount, gae_lambda):
    """ClippedPPOAgent.fill_advantages for policy_gradient_rescaler == GAE: returns (standardised advantages,
    value targets, n_valid)."""
    from coach_b200 import parallel
    adv, tgt, n_valid = gae(rewards, values, game_overs, discount, gae_lambda)
    if parallel.is_distributed():
        # one rollout shard per rank: the moments are those of the whole distributed rollout
        parallel.global_standardize_(adv, int(n_valid.item()))
    else:
        standardize_(adv, n_valid)
    return adv, tgt, n_valid


def nstep_returns(rewards, episode_lengths, discount, n_step):
    """rewards fp64 CUDA [N] (episodes back to back), episode_lengths: host ints summing to N."""
    lib = _lib.load()
    n = rewards.shape[0]
    lens = np.asarray(episode_lengths, dtype=np.int64)
    assert lens.sum() == n
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]])
    ep_start = torch.from_numpy(np.repeat(starts, lens)).to(rewards.device)
    ep_end = torch.from_numpy(np.repeat(starts + lens, lens)).to(rewards.device)
    out = torch.empty_like(rewards)
    _lib.check(lib.cb200_nstep_returns(rewards.data_ptr(), ep_start.data_ptr(), ep_end.data_ptr(), n, float(discount),
                                       int(n_step), out.data_ptr(), _lib.current_stream()))
    return out
