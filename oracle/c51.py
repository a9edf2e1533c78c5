"""CPU restatement of the distributional (C51) learn step.  TEST INFRASTRUCTURE ONLY -- never imported by coach_b200.

  rl_coach/agents/categorical_dqn_agent.py:75-165     z_values, target action, projection loop, PER errors
  rl_coach/agents/rainbow_dqn_agent.py:93-140         same with n-step rewards, should_bootstrap_next_state, double-Q
  rl_coach/architectures/tensorflow_components/heads/categorical_q_head.py:41-57   softmax head, cross entropy, q_values
  rl_coach/architectures/tensorflow_components/general_network.py:352-360          total_loss = reduce_sum(losses)

The numpy part (``c51_targets``) is pinned bit for bit against the unmodified reference agents
(tests/golden/agent_prologues.npz, keys c51_* / rainbow_*, written by oracle/make_golden_agents.py).  The TensorFlow part
(softmax cross entropy and its gradient) is restated in torch fp32 -- parity unpinned, like oracle/nets.py.
"""
from collections import OrderedDict

import numpy as np
import torch

from oracle.nets import _t


def z_values(v_min, v_max, atoms):
    """categorical_dqn_agent.py:77"""
    return np.linspace(v_min, v_max, atoms)


def distribution_prediction_to_q_values(prediction, z):
    """categorical_dqn_agent.py:83-84"""
    return np.dot(prediction, z)


def c51_targets(dist_next, dist_online, dist_select, actions, rewards, bootstrap, gamma_n, z):
    """What learn_from_batch does between the predictions and the train op (categorical_dqn_agent.py:120-152;
    rainbow_dqn_agent.py:107-131 when ``dist_select`` is the online prediction on s' and ``bootstrap`` / ``rewards``
    are the n-step quantities).  dist_*: float32 [B, A, N] softmax outputs; bootstrap: float64 [B] ((1.0 - game_overs)
    for C51).  Returns (TD_targets float32 [B, A, N], target_actions, m float64 [B, N]).

    Written sample by sample: for one sample the reference's vectorised statements reduce to the scalar sequence below
    (atom j ascending; the floor bin is credited before the ceil bin; an integral position credits nothing), every
    operation an IEEE double operation -- hence the same bits (tests/test_oracle_golden.py pins it to the reference)."""
    n_samples, n_atoms = dist_next.shape[0], z.size
    chooser = dist_next if dist_select is None else dist_select
    target_actions = np.argmax(distribution_prediction_to_q_values(chooser, z), axis=1)
    z_lo, z_hi = z[0], z[-1]
    step = z[1] - z[0]
    m = np.zeros((n_samples, n_atoms))
    for b in range(n_samples):
        scale = bootstrap[b] * gamma_n
        probs = dist_next[b, target_actions[b]]
        row = m[b]
        for j in range(n_atoms):
            shifted = rewards[b] + scale * z[j]
            shifted = z_lo if shifted < z_lo else (z_hi if shifted > z_hi else shifted)
            pos = (shifted - z_lo) / step
            below, above = int(np.floor(pos)), int(np.ceil(pos))
            mass = np.float64(probs[j])
            row[below] += mass * (above - pos)
            row[above] += mass * (pos - below)
    targets = np.array(dist_online, dtype=np.float32, copy=True)
    targets[np.arange(n_samples), actions] = m
    return targets, target_actions, m


def categorical_loss(logits, labels):
    """tf.nn.softmax_cross_entropy_with_logits(labels, logits) per (sample, action) -> [B, A]; the network's total
    loss is the SUM over the tensor (no importance weights enter: the head defines its loss itself,
    categorical_q_head.py:53-54, head.py:152-158 only creates the placeholder)."""
    logp = torch.log_softmax(logits, dim=-1)
    return -(labels * logp).sum(dim=-1)


def c51_learn_step(net, online, target, opt, batch, discount, z, n_actions, double_q=False, bootstrap=None,
                   gamma_n=None, clip=None, kink=None):
    """One learn_from_batch step of CategoricalDQNAgent (RainbowDQNAgent's target rule with double_q / bootstrap /
    gamma_n) on a QNetOracle whose head has n_actions * atoms outputs.  TF's gradient of the cross entropy treats the
    labels as constants and is `softmax - labels` whatever the labels sum to (the xent kernel's backprop output; the
    projection loses the mass of integral b_j, so m need not sum to 1): the surrogate below has exactly that gradient."""
    names = list(online.keys())
    params = [online[n].clone().requires_grad_(True) for n in names]
    pd = OrderedDict(zip(names, params))
    N = z.size
    with torch.no_grad():
        ln = net.forward(target, batch["next_states"]).reshape(-1, n_actions, N)
        lo = net.forward(online, batch["states"]).reshape(-1, n_actions, N)
        dist_next = torch.softmax(ln, dim=-1).numpy()
        dist_online = torch.softmax(lo, dim=-1).numpy()
        dist_select = None
        if double_q:
            ls = net.forward(online, batch["next_states"]).reshape(-1, n_actions, N)
            dist_select = torch.softmax(ls, dim=-1).numpy()
    boot = (1.0 - np.asarray(batch["game_overs"], dtype=np.float64)) if bootstrap is None else bootstrap
    g = discount if gamma_n is None else gamma_n
    targets, target_actions, m = c51_targets(dist_next, dist_online, dist_select, batch["actions"],
                                             np.asarray(batch["rewards"], dtype=np.float64), boot, g, z)
    logits = net.forward(pd, batch["states"], kink=kink).reshape(-1, n_actions, N)
    labels = _t(targets, net.dtype)
    loss_rows = categorical_loss(logits, labels)
    total = loss_rows.sum()
    # gradient surrogate: d/dlogits [ logsumexp(logits) - sum(labels * logits) ] = softmax - labels
    surrogate = (torch.logsumexp(logits, dim=-1) - (labels * logits).sum(dim=-1)).sum()
    grads = torch.autograd.grad(surrogate, params, allow_unused=True)
    grads = [gr if gr is not None else torch.zeros_like(p) for gr, p in zip(grads, params)]
    gnorm = torch.sqrt(sum((gr * gr).sum() for gr in grads))
    if clip:
        scale = clip / max(float(gnorm), clip)
        grads = [gr * scale for gr in grads]
    new_params = opt.step([p.detach() for p in params], grads)
    rows = loss_rows.detach().numpy()
    B = rows.shape[0]
    return dict(loss=float(total.detach()), loss_rows=rows, td_errors=rows[np.arange(B), batch["actions"]],
                targets=targets, target_actions=target_actions, m=m,
                grads=OrderedDict(zip(names, [gr.detach() for gr in grads])), grad_norm=float(gnorm),
                new_params=OrderedDict(zip(names, new_params)),
                q_online=distribution_prediction_to_q_values(dist_online, z))
