"""torch-CPU fp32/fp64 restatement of the actor-critic learn steps (ClippedPPO, DDPG, TD3, SAC).
TEST INFRASTRUCTURE ONLY.  The numpy prologues (TD targets, target-policy smoothing, advantage filling) are PINNED to the unmodified reference
(tests/golden/agent_prologues.npz); the network arithmetic (layers, losses, gradients, Adam) is **parity unpinned**
(TensorFlow semantics restated; see oracle/nets.py header).

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
from oracle.rl_math import ac_td_targets, td3_smooth_actions

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
                new_params=OrderedDict(zip(names, new_params)), mean_ratio=float(ex["ratio"].mean().detach()),
                entropy=float(ex["entropy"].detach()), old_mu=old_mu.numpy(), v=ex["v"].detach().numpy())


def make_adam(named, lr, beta1, beta2, eps, dtype=torch.float32):
    return AdamTF([torch.as_tensor(v).to(dtype) for v in named.values()], lr, beta1, beta2, eps, dtype=dtype)


# ---- DDPG / TD3 ---------------------------------------------------------------------------------------------------------
def actor_forward(p, s, scale):
    """obs -> 400 relu -> 300 relu -> A tanh, * scale (ddpg_actor_head.py:46-63)"""
    return torch.tanh(mlp(p[0:6], s, ["relu", "relu", None])) * scale


def ddpg_critic_forward(p, s, a):
    """concat[action, relu(Dense400(obs))] -> Dense300 relu -> Dense1 (general_network.py:252-279: action first)"""
    e = torch.relu(s @ p[0] + p[1])
    h = torch.relu(torch.cat([a, e], dim=1) @ p[2] + p[3])
    return [h @ p[4] + p[5]]


def td3_critic_forward(p, s, a):
    """concat[action, obs] -> two streams (Dense400 relu, Dense300 relu) -> Dense1 each (td3_v_head.py:40-62)"""
    x = torch.cat([a, s], dim=1)
    outs = []
    for k in range(2):
        h = torch.relu(x @ p[4 * k] + p[4 * k + 1])
        h = torch.relu(h @ p[4 * k + 2] + p[4 * k + 3])
        outs.append(h @ p[8 + 2 * k] + p[9 + 2 * k])
    return outs


def ddpg_td3_step(actor, actor_t, critic, critic_t, opt_a, opt_c, batch, discount=0.99, scale=1.0, twin=False,
                  noise=None, noise_clip=0.5, low=-1.0, high=1.0, update_actor=True, dtype=torch.float32):
    """One learn_from_batch of DDPG (twin=False; ddpg_agent.py:137-195) or TD3 (twin=True; td3_agent.py:148-209).
    actor / critic (+ *_t targets): OrderedDict name -> array in creation order.  Returns new params, grads, losses."""
    t = lambda a: torch.as_tensor(np.asarray(a)).to(dtype)     # noqa: E731
    an, cn = list(actor.keys()), list(critic.keys())
    A = [t(actor[n]).clone().requires_grad_(True) for n in an]
    C = [t(critic[n]).clone().requires_grad_(True) for n in cn]
    At, Ct = [t(actor_t[n]) for n in an], [t(critic_t[n]) for n in cn]
    cf = td3_critic_forward if twin else ddpg_critic_forward
    s, s2, a = t(batch["states"]), t(batch["next_states"]), t(batch["actions"])
    with torch.no_grad():
        next_actions = actor_forward(At, s2, scale)
        if twin:
            # pinned restatement (oracle/rl_math.py): fp64 sum / clips, rounded once to the compute dtype
            next_actions = t(td3_smooth_actions(next_actions.numpy(), noise, noise_clip, low, high))
        qn = cf(Ct, s2, next_actions)
        q_next = torch.min(qn[0], qn[1]) if twin else qn[0]
    y = ac_td_targets(batch["rewards"], batch["game_overs"], q_next.numpy(), discount)
    y = t(y.astype(np.float32) if dtype == torch.float32 else y)

    def action_grad(Cp):
        mu = actor_forward(A, s, scale)
        q1 = cf(Cp, s, mu)[0]
        return torch.autograd.grad(-q1.mean(), A, allow_unused=True)

    actor_grads = None
    if not twin:
        actor_grads = action_grad([c.detach() for c in C])          # critic BEFORE its update
    qs = cf(C, s, a)
    loss = sum(((y - q) ** 2).mean() for q in qs)
    cg = torch.autograd.grad(loss, C, allow_unused=True)
    cg = [g if g is not None else torch.zeros_like(p) for g, p in zip(cg, C)]
    newC = opt_c.step([p.detach() for p in C], cg)
    if twin and update_actor:
        actor_grads = action_grad(newC)                              # critic AFTER its update
    newA = [p.detach() for p in A]
    ag = None
    if actor_grads is not None:
        ag = [g if g is not None else torch.zeros_like(p) for g, p in zip(actor_grads, A)]
        newA = opt_a.step([p.detach() for p in A], ag)
    return dict(loss=float(loss.detach()), td_targets=y.numpy(), critic_grads=OrderedDict(zip(cn, cg)),
                actor_grads=OrderedDict(zip(an, ag)) if ag is not None else None,
                new_critic=OrderedDict(zip(cn, newC)), new_actor=OrderedDict(zip(an, newA)))
