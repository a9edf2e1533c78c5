"""Pins the numpy prologues of the reference AGENTS: tests/golden/agent_prologues.npz.

TensorFlow is absent, but everything an agent's ``learn_from_batch`` does *around* its network calls is numpy and
imports under oracle/ref_loader.py.  Each function below calls the UNMODIFIED reference method with a stand-in
``self`` whose ``networks`` return given arrays and record what they are handed:

  DQNAgent.learn_from_batch / DDQNAgent.select_actions      rl_coach/agents/dqn_agent.py:81-113, ddqn_agent.py:42-43
  ValueOptimizationAgent.update_transition_priorities_...   rl_coach/agents/value_optimization_agent.py:74-80
  ClippedPPOAgent.fill_advantages                           rl_coach/agents/clipped_ppo_agent.py:157-207
  DDPGAgent.learn_from_batch                                rl_coach/agents/ddpg_agent.py:137-195
  TD3Agent.learn_from_batch                                 rl_coach/agents/td3_agent.py:148-209
  SoftActorCriticAgent.learn_from_batch                     rl_coach/agents/soft_actor_critic_agent.py:168-280
  Batch                                                     rl_coach/core_types.py:405-649

The recorded arrays (TD targets handed to ``train_and_sync_networks``, priorities handed to the memory, advantages
written into ``transition.info`` ...) are what the oracle restatement AND the CUDA kernels must reproduce bit for bit
from the same inputs (tests/test_oracle_golden.py, tests/test_agent_prologues_gpu.py).

Run in the build container only:   python -m oracle.make_golden_agents          TEST INFRASTRUCTURE ONLY.
"""
import os
import sys
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


class _Sig(object):
    def add_sample(self, *_a, **_k):
        pass


def _ref():
    from oracle import ref_loader
    ref_loader.load()


def _transitions(rng, n, obs_shape, action, with_info=True, obs_dtype=np.float32):
    from rl_coach.core_types import Transition
    ts = []
    for i in range(n):
        if obs_dtype == np.uint8:
            s = rng.randint(0, 256, obs_shape).astype(np.uint8)
            s2 = rng.randint(0, 256, obs_shape).astype(np.uint8)
        else:
            s = rng.randn(*obs_shape).astype(obs_dtype)
            s2 = rng.randn(*obs_shape).astype(obs_dtype)
        a = int(rng.randint(0, action)) if isinstance(action, int) else rng.uniform(-1, 1, action[0]).astype(np.float32)
        t = Transition(state={'observation': s}, action=a, reward=float(rng.randint(-1, 2) if isinstance(action, int)
                                                                       else rng.randn()),
                       next_state={'observation': s2}, game_over=bool(rng.rand() < 0.15))
        if with_info:
            t.info['idx'] = int(rng.randint(0, 1024))
            t.info['weight'] = float(rng.rand() + 0.1)
        ts.append(t)
    return ts


def _ap(keys=("observation",), **alg):
    nw = SimpleNamespace(input_embedders_parameters={k: None for k in keys}, batch_size=alg.pop("batch_size", 64))
    return nw, SimpleNamespace(**alg)


# ---- DQN / DDQN ------------------------------------------------------------------------------------------------------
def golden_dqn(out, rng, B=64, A=6):
    from rl_coach.agents.dqn_agent import DQNAgent
    from rl_coach.agents.ddqn_agent import DDQNAgent
    from rl_coach.agents.value_optimization_agent import ValueOptimizationAgent
    from rl_coach.core_types import Batch
    from rl_coach.memories.non_episodic.prioritized_experience_replay import PrioritizedExperienceReplay
    from rl_coach.memories.memory import MemoryGranularity
    for tag, cls in (("dqn", DQNAgent), ("ddqn", DDQNAgent)):
        ts = _transitions(rng, B, (4,), A)
        batch = Batch(ts)
        q_next = rng.randn(B, A).astype(np.float32)
        q_online = rng.randn(B, A).astype(np.float32)
        q_select = rng.randn(B, A).astype(np.float32)
        q_select[5, 1] = q_select[5, 3] = q_select[5].max() + 1.0            # a tie: np.argmax takes the first
        rec = {}
        nw, alg = _ap(discount=0.99)
        mem = PrioritizedExperienceReplay((MemoryGranularity.Transitions, 1024))
        net = SimpleNamespace(
            target_network="T", online_network=SimpleNamespace(predict=lambda states: q_select.copy()),
            parallel_prediction=lambda pairs: (q_next.copy(), q_online.copy()),
            train_and_sync_networks=lambda states, targets, importance_weights=None: (
                rec.update(states=states, targets=np.array(targets), w=None if importance_weights is None
                           else np.array(importance_weights)) or (0.0, [0.0], 0.0)))
        calls = []
        fake = SimpleNamespace(ap=SimpleNamespace(network_wrappers={'main': nw}, algorithm=alg),
                               networks={'main': net}, q_values=_Sig(), memory=mem,
                               call_memory=lambda f, args: calls.append((f, args)))
        fake.select_actions = lambda ns, q: cls.select_actions(fake, ns, q)
        fake.update_transition_priorities_and_get_weights = \
            lambda e, b: ValueOptimizationAgent.update_transition_priorities_and_get_weights(fake, e, b)
        cls.learn_from_batch(fake, batch)
        assert calls[0][0] == 'update_priorities'
        pidx, perr = calls[0][1]
        out[tag + "_q_next"], out[tag + "_q_online"], out[tag + "_q_select"] = q_next, q_online, q_select
        out[tag + "_actions"] = batch.actions().astype(np.int64)
        out[tag + "_rewards"] = batch.rewards().astype(np.float64)
        out[tag + "_game_overs"] = batch.game_overs().astype(np.uint8)
        out[tag + "_targets"] = rec["targets"]                       # fp32 [B, A], what the train op is fed
        out[tag + "_td_errors"] = np.array(perr, dtype=np.float64)   # what update_priorities is handed
        out[tag + "_prio_idx"] = np.array(pidx, dtype=np.int64)
        out[tag + "_weights"] = rec["w"]
        assert rec["targets"].dtype == np.float32


# ---- Categorical DQN / Rainbow targets -------------------------------------------------------------------------------
def golden_c51(out, rng, B=64, A=6, N=51):
    """categorical_dqn_agent.py:105-165 and rainbow_dqn_agent.py:93-140 with stand-in networks: the distributions the
    networks return are given, the TD_targets handed to the train op and the errors handed to the memory are recorded"""
    from rl_coach.agents.categorical_dqn_agent import CategoricalDQNAgent
    from rl_coach.agents.rainbow_dqn_agent import RainbowDQNAgent
    from rl_coach.core_types import Batch
    from rl_coach.memories.non_episodic.prioritized_experience_replay import PrioritizedExperienceReplay
    from rl_coach.memories.memory import MemoryGranularity

    def softmax(x):
        e = np.exp(x - x.max(axis=-1, keepdims=True))
        return (e / e.sum(axis=-1, keepdims=True)).astype(np.float32)

    for tag, cls in (("c51", CategoricalDQNAgent), ("rainbow", RainbowDQNAgent)):
        ts = _transitions(rng, B, (4,), A)
        for i, t in enumerate(ts):
            t.reward = float(rng.choice([-1.0, 0.0, 1.0, 0.37, -12.5, 11.0]))    # integral b_j, clipped ends, generic
            t.n_step_discounted_rewards = float(rng.randn() * 2)
            t.info['should_bootstrap_next_state'] = bool(rng.rand() < 0.8)
        batch = Batch(ts)
        z = np.linspace(-10.0, 10.0, N)
        d_next = softmax(rng.randn(B, A, N) * 2)
        d_online = softmax(rng.randn(B, A, N))
        d_select = softmax(rng.randn(B, A, N) * 2)
        loss_rows = rng.rand(B, A).astype(np.float32)
        rec = {}
        nw, alg = _ap(discount=0.99, n_step=3)
        mem = PrioritizedExperienceReplay((MemoryGranularity.Transitions, 1024))
        net = SimpleNamespace(
            target_network="T", online_network=SimpleNamespace(predict=lambda states: d_select.copy()),
            parallel_prediction=lambda pairs: (d_next.copy(), d_online.copy()),
            train_and_sync_networks=lambda states, targets, importance_weights=None: (
                rec.update(targets=np.array(targets), w=None if importance_weights is None
                           else np.array(importance_weights)) or (0.0, [loss_rows], 0.0)))
        calls = []
        fake = SimpleNamespace(ap=SimpleNamespace(network_wrappers={'main': nw}, algorithm=alg),
                               networks={'main': net}, q_values=_Sig(), memory=mem, z_values=z,
                               call_memory=lambda f, args: calls.append((f, args)))
        fake.distribution_prediction_to_q_values = lambda pred: cls.distribution_prediction_to_q_values(fake, pred)
        cls.learn_from_batch(fake, batch)
        assert calls[0][0] == 'update_priorities'
        out[tag + "_z"] = z
        out[tag + "_dist_next"], out[tag + "_dist_online"], out[tag + "_dist_select"] = d_next, d_online, d_select
        out[tag + "_actions"] = batch.actions().astype(np.int64)
        out[tag + "_rewards"] = (batch.n_step_discounted_rewards() if tag == "rainbow" else batch.rewards()) \
            .astype(np.float64)
        out[tag + "_game_overs"] = batch.game_overs().astype(np.uint8)
        out[tag + "_bootstrap"] = np.asarray(batch.info('should_bootstrap_next_state'), dtype=np.float64) \
            if tag == "rainbow" else (1.0 - batch.game_overs()).astype(np.float64)
        out[tag + "_gamma_n"] = np.float64(alg.discount ** alg.n_step if tag == "rainbow" else alg.discount)
        out[tag + "_targets"] = rec["targets"]                       # float32 [B, A, N], fed to the train op
        out[tag + "_loss_rows"] = loss_rows
        out[tag + "_prio_errors"] = np.array(calls[0][1][1], dtype=np.float64)
        assert rec["targets"].dtype == np.float32 and rec["targets"].shape == (B, A, N)


# ---- ClippedPPO fill_advantages --------------------------------------------------------------------------------------
def golden_ppo(out, rng):
    from rl_coach.agents.actor_critic_agent import ActorCriticAgent
    from rl_coach.agents.clipped_ppo_agent import ClippedPPOAgent
    from rl_coach.agents.policy_optimization_agent import PolicyGradientRescaler
    from rl_coach.core_types import Batch
    for k, (N, mb, tail) in enumerate([(300, 64, True), (256, 64, False), (97, 32, True)]):
        ts = _transitions(rng, N, (17,), (6,), with_info=False)
        for i, t in enumerate(ts):
            t.game_over = bool(rng.rand() < 0.03)
            t.n_step_discounted_rewards = 0.0
        ts[-1].game_over = not tail            # tail=True: the rollout ends inside an episode (zip() truncation, :203)
        if tail:
            ts[N // 2].game_over = True
        values = rng.randn(N, 1).astype(np.float32)
        batch = Batch(ts)
        nw, alg = _ap(discount=0.99, gae_lambda=0.95, estimate_state_value_using_gae=True, batch_size=mb)
        seen = []

        def predict(d):
            start = sum(seen)
            n = len(d['observation'])
            seen.append(n)
            return [values[start:start + n]]
        fake = SimpleNamespace(ap=SimpleNamespace(network_wrappers={'main': nw}, algorithm=alg),
                               networks={'main': SimpleNamespace(online_network=SimpleNamespace(predict=predict))},
                               state_values=_Sig(), action_advantages=_Sig(),
                               policy_gradient_rescaler=PolicyGradientRescaler.GAE)
        fake.discount = lambda x, g: ActorCriticAgent.discount(fake, x, g)
        fake.get_general_advantage_estimation_values = \
            lambda r, v: ActorCriticAgent.get_general_advantage_estimation_values(fake, r, v)
        ClippedPPOAgent.fill_advantages(fake, batch)
        filled = [i for i, t in enumerate(ts) if 'advantage' in t.info]
        assert filled == list(range(len(filled)))
        out["ppo_values_%d" % k] = values
        out["ppo_rewards_%d" % k] = batch.rewards().astype(np.float64)
        out["ppo_game_overs_%d" % k] = batch.game_overs().astype(np.uint8)
        out["ppo_adv_%d" % k] = np.array([ts[i].info['advantage'] for i in filled], dtype=np.float64)
        out["ppo_vtgt_%d" % k] = np.array([ts[i].info['gae_based_value_target'] for i in filled], dtype=np.float64)
        out["ppo_minibatch_%d" % k] = mb
    out["ppo_cases"] = 3


# ---- DDPG / TD3 ------------------------------------------------------------------------------------------------------
def _critic_actor_stubs(rec, next_actions, actions_mean, q_outputs, action_grads):
    class Critic(object):
        def __init__(self):
            self.target_network = SimpleNamespace(predict=self._target_predict)
            self.online_network = SimpleNamespace(predict=self._online_predict,
                                                  gradients_wrt_inputs=[{'action': 'g0'}, {'action': 'g1'},
                                                                        {'action': 'g2'}, {'action': 'g3'}])

        def _target_predict(self, inputs):
            rec["target_critic_action"] = np.array(inputs['action'])
            return q_outputs

        def _online_predict(self, inputs, outputs=None):
            rec["action_grad_fetch"] = outputs
            rec["action_grad_action_input"] = np.array(inputs['action'])
            return action_grads

        def train_and_sync_networks(self, inputs, targets, **kw):
            rec["critic_train_action"] = np.array(inputs['action'])
            rec["td_targets"] = np.array(targets)
            return 0.0, [0.0], 0.0

    class Actor(object):
        has_global = False
        target_network = "T"

        def __init__(self):
            self.online_network = SimpleNamespace(predict=self._predict, gradients_weights_ph=["ph0"],
                                                  weighted_gradients=["wg0"])

        def parallel_prediction(self, pairs):
            return next_actions.copy(), actions_mean.copy()

        def _predict(self, states, outputs=None, initial_feed_dict=None):
            rec["actor_feed"] = np.array(initial_feed_dict["ph0"])
            return ["grads"]

        def apply_gradients_to_online_network(self, grads, additional_inputs=None):
            rec["actor_applied"] = True
    return Critic(), Actor()


def golden_ddpg_td3(out, rng, B=64, D=17, A=6):
    from rl_coach.agents.ddpg_agent import DDPGAgent
    from rl_coach.agents.td3_agent import TD3Agent
    from rl_coach.core_types import Batch
    from rl_coach.spaces import BoxActionSpace
    cases = [("ddpg0", DDPGAgent, dict(clip_critic_targets=None, use_non_zero_discount_for_terminal_states=False)),
             ("ddpg1", DDPGAgent, dict(clip_critic_targets=(-1.5, 1.5), use_non_zero_discount_for_terminal_states=False)),
             ("ddpg2", DDPGAgent, dict(clip_critic_targets=None, use_non_zero_discount_for_terminal_states=True)),
             ("td3", TD3Agent, dict(clip_critic_targets=None, use_non_zero_discount_for_terminal_states=False,
                                    policy_noise=0.2, noise_clipping=0.5, update_policy_every_x_episode_steps=2))]
    for tag, cls, alg_kw in cases:
        ts = _transitions(rng, B, (D,), (A,), with_info=False)
        batch = Batch(ts)
        next_actions = rng.uniform(-1, 1, (B, A)).astype(np.float32)
        actions_mean = rng.uniform(-1, 1, (B, A)).astype(np.float32)
        q1 = (rng.randn(B, 1) * 3).astype(np.float32)
        q2 = (rng.randn(B, 1) * 3).astype(np.float32)
        q_outputs = [q1, q2, np.minimum(q1, q2), np.float32(q1.mean())] if cls is TD3Agent else [q1, np.float32(q1.mean())]
        action_grads = rng.randn(B, A).astype(np.float32)
        rec = {}
        critic, actor = _critic_actor_stubs(rec, next_actions, actions_mean, q_outputs, action_grads)
        nw, alg = _ap(discount=0.99, **alg_kw)
        fake = SimpleNamespace(ap=SimpleNamespace(network_wrappers={'actor': nw, 'critic': nw}, algorithm=alg),
                               networks={'actor': actor, 'critic': critic}, TD_targets_signal=_Sig(),
                               training_iteration=0,
                               spaces=SimpleNamespace(action=BoxActionSpace(A, -0.9, 0.9)))
        seed = int(rng.randint(0, 2 ** 31 - 1))
        np.random.seed(seed)
        cls.learn_from_batch(fake, batch)
        out[tag + "_next_actions"], out[tag + "_q1"], out[tag + "_q2"] = next_actions, q1, q2
        out[tag + "_rewards"] = batch.rewards().astype(np.float64)
        out[tag + "_game_overs"] = batch.game_overs().astype(np.uint8)
        out[tag + "_td_targets"] = rec["td_targets"]                   # fp64 [B, 1] as handed to the train op
        out[tag + "_actor_feed"] = rec["actor_feed"]                   # -action_gradients
        out[tag + "_action_grads"] = action_grads
        out[tag + "_critic_train_action"] = rec["critic_train_action"]
        out[tag + "_batch_actions"] = batch.actions()
        out[tag + "_grad_fetch"] = np.array([rec["action_grad_fetch"]])
        out[tag + "_clip"] = np.array(alg_kw["clip_critic_targets"] or (0.0, 0.0), dtype=np.float64)
        out[tag + "_nonzero_terminal"] = int(alg_kw["use_non_zero_discount_for_terminal_states"])
        if cls is TD3Agent:
            np.random.seed(seed)
            out[tag + "_noise"] = np.random.normal(0, 0.2, next_actions.shape)       # the draw the agent consumed
            out[tag + "_smoothed_actions"] = rec["target_critic_action"]              # after noise + clip to the space
            out[tag + "_space"] = np.array([-0.9, 0.9])


# ---- SAC -------------------------------------------------------------------------------------------------------------
def golden_sac(out, rng, B=64, D=17, A=6):
    from rl_coach.agents.soft_actor_critic_agent import SoftActorCriticAgent
    from rl_coach.core_types import Batch
    ts = _transitions(rng, B, (D,), (A,), with_info=False)
    batch = Batch(ts)
    rec = {}
    mu, std = rng.randn(B, A).astype(np.float32), rng.rand(B, A).astype(np.float32)
    raw, act = rng.randn(B, A).astype(np.float32), np.tanh(rng.randn(B, A)).astype(np.float32)
    logp = rng.randn(B).astype(np.float32)
    q_min = rng.randn(B, 1).astype(np.float32)
    v_next = rng.randn(B, 1).astype(np.float32)
    dlogp = [rng.randn(5, 3).astype(np.float32), rng.randn(3).astype(np.float32)]
    dq = [rng.randn(5, 3).astype(np.float32), rng.randn(3).astype(np.float32)]
    dq_da = rng.randn(B, A).astype(np.float32)

    def policy_predict(inputs, outputs=None, initial_feed_dict=None):
        if outputs is None:
            return [mu, std, raw, act, logp, np.float32(logp.mean())]
        if outputs == "wg5":
            rec["dlogp_feed"] = np.array(initial_feed_dict["ph5"])
            return dlogp
        rec["dq_feed"] = np.array(initial_feed_dict["ph3"])
        return dq
    policy_net = SimpleNamespace(predict=policy_predict, gradients_weights_ph=["ph%d" % i for i in range(6)],
                                 weighted_gradients=["wg%d" % i for i in range(6)],
                                 apply_gradients=lambda g: rec.update(policy_grads=[np.array(x) for x in g]))
    q_head = SimpleNamespace(q1_output="q1", q2_output="q2", q1_loss="l1", q2_loss="l2")

    def q_predict(inputs, outputs=None):
        if outputs is None:
            rec["q_action_input"] = np.array(inputs['output_0_0'])
            return [q_min]
        if isinstance(outputs, list):
            return q_min, q_min
        return dq_da

    def q_train(inputs, targets, additional_fetches=None):
        rec["q_train_action"] = np.array(inputs['output_0_0'])
        rec["td_targets"] = np.array(targets)
        return 0.0, [0.0], 0.0, (0.0, 0.0)
    q_net = SimpleNamespace(predict=q_predict, output_heads=[q_head], train_on_batch=q_train,
                            gradients_wrt_inputs=[{}, {'output_0_0': 'gq'}])
    v_wrap = SimpleNamespace(
        online_network=SimpleNamespace(train_on_batch=lambda i, t: (rec.update(value_targets=np.array(t)) or [0.0])),
        target_network=SimpleNamespace(predict=lambda i: v_next))
    nw, alg = _ap(discount=0.99)
    sig = _Sig()
    fake = SimpleNamespace(ap=SimpleNamespace(network_wrappers={'v': nw, 'q': nw, 'policy': nw}, algorithm=alg),
                           networks={'v': v_wrap, 'q': SimpleNamespace(online_network=q_net),
                                     'policy': SimpleNamespace(online_network=policy_net)},
                           policy_means=sig, policy_logsig=sig, policy_logprob_sampled=sig, q1_values=sig,
                           q2_values=sig, policy_grads=sig, v_onl_ys=sig, v_tgt_ns=sig, TD_err1=sig, TD_err2=sig)
    SoftActorCriticAgent.learn_from_batch(fake, batch)
    out["sac_q_min"], out["sac_logp"], out["sac_v_next"] = q_min, logp, v_next
    out["sac_rewards"] = batch.rewards().astype(np.float64)
    out["sac_game_overs"] = batch.game_overs().astype(np.uint8)
    out["sac_value_targets"] = rec["value_targets"]                      # [B, 1] = min Q(s, a~pi) - log pi
    out["sac_td_targets"] = rec["td_targets"]                            # fp64 [B, 1]
    out["sac_dlogp_feed"] = rec["dlogp_feed"]                            # scalar 1.0 (sum over the batch of mean log pi)
    out["sac_dq_feed"], out["sac_dq_da"] = rec["dq_feed"], dq_da
    for i in range(2):
        out["sac_dlogp_%d" % i], out["sac_dq_%d" % i], out["sac_pgrad_%d" % i] = dlogp[i], dq[i], rec["policy_grads"][i]
    out["sac_sampled_actions"], out["sac_q_action_input"] = act, rec["q_action_input"]


# ---- Batch -----------------------------------------------------------------------------------------------------------
def golden_batch(out, rng, n=24):
    from rl_coach.core_types import Batch
    ts = _transitions(rng, n, (3, 3, 2), 4, obs_dtype=np.uint8)
    for i, t in enumerate(ts):
        t.n_step_discounted_rewards = float(rng.randn())
    b = Batch(ts)
    out["batch_states"] = b.states(["observation"])["observation"]
    out["batch_next_states"] = b.next_states(["observation"])["observation"]
    out["batch_actions"], out["batch_actions_x"] = b.actions(), b.actions(expand_dims=True)
    out["batch_rewards"], out["batch_rewards_x"] = b.rewards(), b.rewards(expand_dims=True)
    out["batch_game_overs"], out["batch_game_overs_x"] = b.game_overs(), b.game_overs(expand_dims=True)
    out["batch_nstep"] = b.n_step_discounted_rewards()
    out["batch_idx"], out["batch_weight"] = b.info("idx"), b.info("weight")
    b.slice(3, 11)
    out["batch_slice_rewards"], out["batch_slice_states"] = b.rewards(), b.states(["observation"])["observation"]
    out["batch_slice_size"] = b.size


def main():
    _ref()
    rng = np.random.RandomState(20240)
    out = {}
    golden_dqn(out, rng)
    golden_ppo(out, rng)
    golden_ddpg_td3(out, rng)
    golden_sac(out, rng)
    golden_batch(out, rng)
    golden_episodic(out, np.random.RandomState(77))
    golden_filters(out, np.random.RandomState(78))
    golden_c51(out, np.random.RandomState(79))
    golden_frame_stream(out, np.random.RandomState(80))
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "agent_prologues.npz"), **out)
    print("agent_prologues", len(out), "arrays")



# ---- EpisodicExperienceReplay: store / store_episode / n-step relink / eviction / sample ----------------------------
def golden_episodic(out, rng):
    """rl_coach/memories/episodic/episodic_experience_replay.py:60-330 + Episode (core_types.py:771-820): scripted
    sessions; every transition carries a unique observation id so that the n-step ``next_state`` relink, the
    ``should_bootstrap_next_state`` flags and whole-episode eviction are visible in the recorded contents."""
    from rl_coach.core_types import Episode, Transition
    from rl_coach.memories.episodic.episodic_experience_replay import EpisodicExperienceReplay
    from rl_coach.memories.memory import MemoryGranularity
    for k, (n_step, cap) in enumerate([(3, 150), (-1, 150), (1, 10000)]):
        mem = EpisodicExperienceReplay((MemoryGranularity.Transitions, cap), n_step=n_step)
        uid = [0]
        script = []        # (kind, length): kind 0 = store() transition by transition, 1 = store_episode()

        def make(T, last_done=True):
            ts = []
            for i in range(T):
                s, s2 = uid[0], uid[0] + 100000
                uid[0] += 1
                ts.append(Transition(state={'observation': np.array([s], dtype=np.float32)}, action=int(rng.randint(0, 3)),
                                     reward=float(np.round(rng.randn(), 3)),
                                     next_state={'observation': np.array([s2], dtype=np.float32)},
                                     game_over=bool(last_done and i == T - 1)))
            return ts
        all_ts = []
        for e in range(14):
            T = int(rng.randint(1, 40))
            kind = int(e in (4, 9))
            ts = make(T)
            all_ts.append(ts)
            if kind == 0:
                for t in ts:
                    mem.store(t)
            else:
                ep = Episode(n_step=n_step)
                for t in ts:
                    ep.insert(t)
                mem.store_episode(ep)
            script.append((kind, T))
        tail = make(int(rng.randint(1, 8)), last_done=False)      # an open episode at the end
        all_ts.append(tail)
        for t in tail:
            mem.store(t)
        script.append((0, len(tail)))
        out["epi%d_script" % k] = np.array(script, dtype=np.int64)
        out["epi%d_n_step" % k], out["epi%d_capacity" % k] = n_step, cap
        flat = [t for ts in all_ts for t in ts]
        out["epi%d_in_state" % k] = np.array([t.state['observation'][0] for t in flat], dtype=np.float32)
        out["epi%d_in_reward" % k] = np.array([t.reward for t in flat], dtype=np.float64)
        out["epi%d_in_action" % k] = np.array([t.action for t in flat], dtype=np.int64)
        # contents after the session (complete episodes only = what the agents train on)
        nc = mem.num_transitions_in_complete_episodes()
        kept = mem.transitions[:nc]
        out["epi%d_state" % k] = np.array([t.state['observation'][0] for t in kept], dtype=np.float32)
        out["epi%d_next_state" % k] = np.array([t.next_state['observation'][0] for t in kept], dtype=np.float32)
        out["epi%d_reward" % k] = np.array([t.reward for t in kept], dtype=np.float64)
        out["epi%d_game_over" % k] = np.array([t.game_over for t in kept], dtype=np.uint8)
        out["epi%d_nstep" % k] = np.array([t.n_step_discounted_rewards for t in kept], dtype=np.float64)
        out["epi%d_bootstrap" % k] = np.array([int(t.info.get('should_bootstrap_next_state', -1)) for t in kept],
                                              dtype=np.int64)
        out["epi%d_counts" % k] = np.array([mem.num_transitions(), nc, mem.num_complete_episodes(), mem.length()],
                                           dtype=np.int64)
        np.random.seed(11 + k)
        got = mem.sample(32)
        out["epi%d_sample_state" % k] = np.array([t.state['observation'][0] for t in got], dtype=np.float32)
    out["epi_cases"] = 3


# ---- InputFilter: observe-time chain (to-uint8 -> frame stacking, reward clipping / rescale) + Transition lists ---------

def golden_frame_stream(out, rng, H=16, W=16, K=4):
    """An episode stream through the reference's ObservationStackingFilter (observation_stacking_filter.py:27-115): raw
    frames in, the stacked state / next_state every transition carries (LazyStack materialised with np.array, i.e.
    np.stack(..., axis=-1)) out -- what a frame-deduplicated ring must reproduce byte for byte when it gathers."""
    from rl_coach.filters.observation.observation_stacking_filter import ObservationStackingFilter
    flt = ObservationStackingFilter(K)
    lengths = [5, 1, 9, 2, 14, 3, 11]
    frames, starts, states, next_states, actions, rewards, dones = [], [], [], [], [], [], []
    for L in lengths:
        flt.reset()
        f0 = rng.randint(0, 256, (H, W)).astype(np.uint8)
        frames.append(f0)
        starts.append(len(frames) - 1)
        s = flt.filter(f0)
        for t in range(L):
            f = rng.randint(0, 256, (H, W)).astype(np.uint8)
            frames.append(f)
            s2 = flt.filter(f)
            states.append(np.array(s))
            next_states.append(np.array(s2))
            actions.append(int(rng.randint(0, 4)))
            rewards.append(float(rng.randn()))
            dones.append(t == L - 1)
            s = s2
    out["fs_frames"] = np.stack(frames)
    out["fs_episode_lengths"] = np.array(lengths, dtype=np.int64)
    out["fs_states"] = np.stack(states)
    out["fs_next_states"] = np.stack(next_states)
    out["fs_actions"] = np.array(actions, dtype=np.int64)
    out["fs_rewards"] = np.array(rewards, dtype=np.float64)
    out["fs_game_overs"] = np.array(dones, dtype=np.uint8)
    assert out["fs_states"].shape == (sum(lengths), H, W, K) and out["fs_states"].dtype == np.uint8


def golden_filters(out, rng):
    """rl_coach/filters/filter.py:295-350 driven like Agent.observe does (agents/agent.py:905-973): one EnvResponse per
    environment step through the input filter, reset at episode ends; then a list of Transitions (the pre-network-filter
    call of Agent.train, agent.py:735).  Filters: observation_to_uint8_filter.py, observation_stacking_filter.py (LazyStack,
    first-frame replication), reward_clipping_filter.py (truthiness quirk Q13), reward_rescale_filter.py."""
    from rl_coach.core_types import EnvResponse, Transition
    from rl_coach.filters.filter import InputFilter
    from rl_coach.filters.observation.observation_stacking_filter import ObservationStackingFilter
    from rl_coach.filters.observation.observation_to_uint8_filter import ObservationToUInt8Filter
    from rl_coach.filters.reward.reward_clipping_filter import RewardClippingFilter
    from rl_coach.filters.reward.reward_rescale_filter import RewardRescaleFilter
    f = InputFilter(is_a_reference_filter=False)
    f.add_observation_filter('observation', 'to_uint8', ObservationToUInt8Filter(0, 1))
    f.add_observation_filter('observation', 'stacking', ObservationStackingFilter(4))
    f.add_reward_filter('rescale', RewardRescaleFilter(2.0))
    f.add_reward_filter('clipping', RewardClippingFilter(-1.0, 1.0))
    T = 23
    frames = rng.rand(T, 6, 5)
    rewards = rng.randn(T) * 2
    ends = {8, 15}                                        # the filter is reset after these steps (episode ends)
    stacked, filt_r = [], []
    for t in range(T):
        er = EnvResponse(next_state={'observation': frames[t]}, reward=float(rewards[t]), game_over=t in ends)
        res = f.filter(er)[0]
        stacked.append(np.array(res.next_state['observation']))
        filt_r.append(res.reward)
        if t in ends:
            f.reset()
    out["flt_frames"], out["flt_rewards"], out["flt_ends"] = frames, rewards, np.array(sorted(ends))
    out["flt_stacked"], out["flt_filtered_rewards"] = np.array(stacked), np.array(filt_r, dtype=np.float64)
    # peek without updating the stack (update_internal_state=False), then continue
    peek = f.filter(EnvResponse(next_state={'observation': frames[0]}, reward=0.5, game_over=False),
                    update_internal_state=False)[0]
    out["flt_peek"] = np.array(peek.next_state['observation'])
    # Transition list through a reward-only + to-uint8 filter (stateless): states first, then next states
    g = InputFilter(is_a_reference_filter=False)
    g.add_observation_filter('observation', 'to_uint8', ObservationToUInt8Filter(0, 2))
    g.add_reward_filter('clipping', RewardClippingFilter(-1.0, 0))          # upper bound 0 is ignored (quirk Q13)
    ts = [Transition(state={'observation': rng.rand(3) * 2}, action=0, reward=float(rng.randn() * 3),
                     next_state={'observation': rng.rand(3) * 2}, game_over=False) for _ in range(7)]
    out["flt_t_states"] = np.array([t.state['observation'] for t in ts])
    out["flt_t_next"] = np.array([t.next_state['observation'] for t in ts])
    out["flt_t_rewards"] = np.array([t.reward for t in ts])
    res = g.filter(ts)
    out["flt_t_states_out"] = np.array([t.state['observation'] for t in res])
    out["flt_t_next_out"] = np.array([t.next_state['observation'] for t in res])
    out["flt_t_rewards_out"] = np.array([t.reward for t in res])


if __name__ == "__main__":
    main()
