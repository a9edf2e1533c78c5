"""torch-CPU restatement of the Soft Actor-Critic learn step.  TEST INFRASTRUCTURE ONLY.  TD / value targets are pinned to the
unmodified reference (tests/golden/agent_prologues.npz); the network arithmetic is **parity unpinned** (TensorFlow
semantics restated, see oracle/nets.py).

Sources: agents/soft_actor_critic_agent.py:168-280; heads/sac_head.py:49-97 (policy: [mu | log sigma], clip [-20, 2],
reparameterised sample, tanh squash + log-prob correction with eps 1e-6); heads/sac_q_head.py:59-119 (two Q heads,
obs/action embeddings summed, output min(Q1,Q2), loss 0.5*mean((Q_k-y)^2) each); heads/v_head.py:35-51.
The three policy-network evaluations of the reference re-sample the Gaussian noise: the three samples are
explicit inputs here.
"""
import math
from collections import OrderedDict

import numpy as np
import torch

from oracle.actor_critic import mlp
from oracle.rl_math import ac_td_targets

LOG_SIG_MIN, LOG_SIG_MAX, EPS = -20.0, 2.0, 1e-6


def policy_sample(p, s, eps):
    z = mlp(p[0:6], s, ["relu", "relu", None])
    A = z.shape[1] // 2
    mu, ls = z[:, :A], torch.clamp(z[:, A:], LOG_SIG_MIN, LOG_SIG_MAX)
    raw = mu + torch.exp(ls) * eps
    act = torch.tanh(raw)
    # MultivariateNormalDiag.log_prob(raw) with raw = mu + sigma*eps, minus the squash correction
    zz = (raw - mu) / torch.exp(ls)
    logp = (-0.5 * zz * zz - ls - 0.5 * math.log(2 * math.pi)).sum(1) - torch.log(1 - act ** 2 + EPS).sum(1)
    return act, logp


def q_heads(p, s, a):
    outs = []
    for k in range(2):
        o = 8 * k
        e = torch.relu(s @ p[o] + p[o + 1]) + torch.relu(a @ p[o + 2] + p[o + 3])
        h = torch.relu(e @ p[o + 4] + p[o + 5])
        outs.append(h @ p[o + 6] + p[o + 7])
    return outs


def sac_step(policy, q, v, v_target, opt_p, opt_q, opt_v, batch, noise, discount=0.99, dtype=torch.float32):
    """policy / q / v / v_target: OrderedDict name -> array in creation order; noise: three [B, A] arrays."""
    t = lambda a: torch.as_tensor(np.asarray(a)).to(dtype)     # noqa: E731
    pn, qn, vn = list(policy.keys()), list(q.keys()), list(v.keys())
    P = [t(policy[n]).clone().requires_grad_(True) for n in pn]
    Q = [t(q[n]).clone().requires_grad_(True) for n in qn]
    V = [t(v[n]).clone().requires_grad_(True) for n in vn]
    Vt = [t(v_target[n]) for n in vn]
    s, s2, a = t(batch["states"]), t(batch["next_states"]), t(batch["actions"])
    e1, e2, e3 = (t(n) for n in noise)
    Qd = [x.detach() for x in Q]
    # 1./2. sampled action, its log-prob, min-Q and dq_da (all with noise sample 1)
    with torch.no_grad():
        a1, logp1 = policy_sample(P, s, e1)
    a1v = a1.clone().requires_grad_(True)
    q1, q2 = q_heads(Qd, s, a1v)
    qmin = torch.min(q1, q2)
    dq_da = torch.autograd.grad(qmin.mean(), a1v)[0]
    log_target = qmin.detach()[:, 0]
    # 3. policy gradients: d mean(logp)(noise 2) - sum dq_da * d a(noise 3)
    _, logp2 = policy_sample(P, s, e2)
    a3, _ = policy_sample(P, s, e3)
    g_logp = torch.autograd.grad(logp2.mean(), P, allow_unused=True)
    g_q = torch.autograd.grad(a3, P, grad_outputs=dq_da, allow_unused=True)
    pg = [(gl if gl is not None else torch.zeros_like(p)) - (gq if gq is not None else torch.zeros_like(p))
          for gl, gq, p in zip(g_logp, g_q, P)]
    newP = opt_p.step([p.detach() for p in P], pg)
    # 4. V
    v_targets = (log_target - logp1).reshape(-1, 1)
    vv = mlp(V[0:6], s, ["relu", "relu", None])
    v_loss = ((vv - v_targets) ** 2).mean()
    vg = torch.autograd.grad(v_loss, V, allow_unused=True)
    vg = [g if g is not None else torch.zeros_like(p) for g, p in zip(vg, V)]
    newV = opt_v.step([p.detach() for p in V], vg)
    # 5. Q
    with torch.no_grad():
        v_next = mlp(Vt[0:6], s2, ["relu", "relu", None])
    y = ac_td_targets(batch["rewards"], batch["game_overs"], v_next.numpy(), discount)
    y = t(y.astype(np.float32) if dtype == torch.float32 else y)
    qa1, qa2 = q_heads(Q, s, a)
    q_loss = 0.5 * ((qa1 - y) ** 2).mean() + 0.5 * ((qa2 - y) ** 2).mean()
    qg = torch.autograd.grad(q_loss, Q, allow_unused=True)
    qg = [g if g is not None else torch.zeros_like(p) for g, p in zip(qg, Q)]
    newQ = opt_q.step([p.detach() for p in Q], qg)
    return dict(q_loss=float(q_loss.detach()), v_loss=float(v_loss.detach()), logp=logp1.numpy(),
                sampled=a1.numpy(), log_target=log_target.numpy(), dq_da=dq_da.numpy(), td_targets=y.numpy(),
                policy_grads=OrderedDict(zip(pn, pg)), q_grads=OrderedDict(zip(qn, qg)),
                v_grads=OrderedDict(zip(vn, vg)), new_policy=OrderedDict(zip(pn, newP)),
                new_q=OrderedDict(zip(qn, newQ)), new_v=OrderedDict(zip(vn, newV)))
