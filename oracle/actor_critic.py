"""torch-CPU fp32/fp64 restatement of the actor-critic learn steps (ClippedPPO, DDPG, TD3, SAC).
TEST INFRASTRUCTURE ONLY.  **Parity unpinned** (TensorFlow semantics restated; see oracle/nets.py header).

Sources restated
  ClippedPPO  agents/clipped_ppo_agent.py:157-308; heads/ppo_head.py:52-144; heads/v_head.py:35-51;
              presets/Mujoco_ClippedPPO.py:30-45 (two separate sub-networks Dense(64)-tanh x2)
  DDPG        agents/ddpg_agent.py:137-195; heads/ddpg_actor_head.py:46-63; heads/ddpg_v_head.py
  TD3         agents/td3_agent.py:148-209; heads/td3_v_head.py:40-70
  SAC         agents/soft_actor_critic_agent.py:168-280; heads/sac_head.py:49-97; heads/sac_q_head.py:59-119
"""
import math
from collections import OrderedDict

import numpy as np
import torch

from oracle.nets import AdamTF

EPS = 1e-15


def mlp(params, x, acts):
    """params: flat list [W0, b0, W1, b1, ...]; acts: list of 'tanh' | 'relu' | None per layer."""
    h = x
    for i, act in enumerate(acts):
        h = h @ params[2 * i] + params[2 * i + 1]
        if act == "tanh":
            h = torch.tanh(h)
        elif act == "relu":
            h = torch.relu(h)
    return h


# ---- ClippedPPO -------------------------------------------------------------------------------------------------------
def ppo_logp(mu, logstd, actions):
    sigma = torch.exp(logstd) + EPS
    z = (actions - mu) / sigma
    k = actions.shape[1]
    return -0.5 * (z * z).sum(1) - torch.log(sigma).sum() - 0.5 * k * math.log(2 * math.pi)


def ppo_losses(v_params, p_params, logstd, old_mu, old_logstd, states, actions, advantages, value_targets, clip_eps,
               beta_entropy):
    """returns (total, value_loss, surrogate(+entropy reg), extras)"""
    v = mlp(v_params, states, ["tanh", "tanh", None])                    # [B,1]
    value_loss = ((v[:, 0] - value_targets) ** 2).mean()                 # VHead MSE, head.py:172-177
    mu = mlp(p_params, states, ["tanh", "tanh", None])
    logp = ppo_logp(mu, logstd, actions)
    logp_old = ppo_logp(old_mu, old_logstd, actions)
    ratio = torch.exp(logp - logp_old)
    clipped = torch.clamp(ratio, 1 - clip_eps, 1 + clip_eps)
    surrogate = -torch.min(ratio * advantages, clipped * advantages).mean()
    k = actions.shape[1]
    entropy = 0.5 * k * (1 + math.log(2 * math.pi)) + torch.log(torch.exp(logstd) + EPS).sum()
    policy_loss = surrogate - beta_entropy * entropy
    return value_loss + policy_loss, value_loss, policy_loss, dict(ratio=ratio, clipped=clipped, entropy=entropy,
                                                                   v=v, mu=mu)


def ppo_minibatch_step(named, old_named, opt, mb, clip_eps, beta_entropy, dtype=torch.float32):
    """named: OrderedDict of ALL online parameters in creation order:
         v: W0 b0 W1 b1 Wv bv rescaler | p: W0 b0 W1 b1 Wmu bmu logstd rescaler
    mb: dict(states [B,D], actions [B,A], advantages [B], value_targets [B]).  Returns loss terms, grads, new params."""
    names = list(named.keys())
    params = [torch.as_tensor(named[n]).to(dtype).clone().requires_grad_(True) for n in names]
    old = [torch.as_tensor(old_named[n]).to(dtype) for n in names]
    v_params, p_params, logstd = params[0:6], params[7:13], params[13].reshape(-1)
    t = lambda a: torch.as_tensor(np.asarray(a)).to(dtype)      # noqa: E731
    states = t(mb["states"])
    with torch.no_grad():
        old_mu = mlp(old[7:13], states, ["tanh", "tanh", None])
    total, vl, pl, ex = ppo_losses(v_params, p_params, logstd, old_mu, old[13].reshape(-1), states, t(mb["actions"]),
                                   t(mb["advantages"]), t(mb["value_targets"]), clip_eps, beta_entropy)
    grads = torch.autograd.grad(total, params, allow_unused=True)
    grads = [g if g is not None else torch.zeros_like(p) for g, p in zip(grads, params)]
    gnorm = torch.sqrt(sum((g * g).sum() for g in grads))
    new_params = opt.step([p.detach() for p in params], grads)
    return dict(total=float(total.detach()), value_loss=float(vl.detach()), policy_loss=float(pl.detach()),
                grad_norm=float(gnorm), grads=OrderedDict(zip(names, [g.detach() for g in grads])),
                new_params=OrderedDict(zip(names, new_params)), mean_ratio=float(ex["ratio"].mean()),
                entropy=float(ex["entropy"].detach()), old_mu=old_mu.numpy(), v=ex["v"].detach().numpy())


def make_adam(named, lr, beta1, beta2, eps, dtype=torch.float32):
    return AdamTF([torch.as_tensor(v).to(dtype) for v in named.values()], lr, beta1, beta2, eps, dtype=dtype)
