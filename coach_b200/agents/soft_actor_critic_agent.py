"""Soft Actor-Critic learn step on the GPU.  Drop-in for
``rl_coach/agents/soft_actor_critic_agent.py:168-280`` with the heads of
``architectures/tensorflow_components/heads/{sac_head.py:49-97, sac_q_head.py:59-119, v_head.py:35-51}``.

Networks (agents/soft_actor_critic_agent.py:40-100, presets/Mujoco_SAC.py:38-46), one flat buffer + Adam each:
  policy : obs -> Dense(256) relu -> Dense(256) relu -> Dense(2A) = [mu | log sigma]     (log sigma clipped to [-20, 2])
  q      : two heads, each  relu(Dense(256)(obs)) + relu(Dense(256)(action)) -> Dense(256) relu -> Dense(1)
  v      : obs -> Dense(256) relu -> Dense(256) relu -> Dense(1), plus a polyak target (tau = 0.005)

One learn step (the reference spends 7-9 ``sess.run`` calls on it):
  1. policy forward; sample a~1 = tanh(mu + sigma*eps1), log pi(a~1)
  2. Q(s, a~1): log_target = min(Q1, Q2); dq_da = d mean_b(min(Q1,Q2)) / da                 (critic data-gradient pass)
  3. policy gradient = d mean log pi (noise eps2) / dphi  -  sum dq_da * d a~(noise eps3) / dphi   -> Adam(policy)
     (three separate evaluations in the reference => three independent noise samples)
  4. V: targets = log_target - log pi(a~1); MSE                                               -> Adam(v)
  5. Q: y = r + (1 - done) * gamma * V_target(s'); 0.5*mean((Q1-y)^2) + 0.5*mean((Q2-y)^2)    -> Adam(q)
  6. V target polyak update by the train() driver (agents/agent.py:755-761)
"""
import numpy as np
import torch

from coach_b200 import _lib, parallel
from coach_b200.agents.ddpg_agent import GraphedKernels, _Net
from coach_b200.architectures.layers import Dense, Workspace
from coach_b200.architectures.network import ParamStore, Sequential
from coach_b200.base_parameters import AgentParameters, AlgorithmParameters, EnvironmentSteps, NetworkParameters
from coach_b200.memories.experience_replay import ExperienceReplayParameters
from coach_b200.utils import dynamic_import_and_instantiate_module_from_params

RELU = 1


class SACNetworkParameters(NetworkParameters):
    def __init__(self):
        super().__init__()
        self.batch_size = 256
        self.learning_rate = 0.0003
        self.hidden_units = 256


class SoftActorCriticAlgorithmParameters(AlgorithmParameters):
    def __init__(self):
        super().__init__()
        self.num_steps_between_copying_online_weights_to_target = EnvironmentSteps(1)
        self.rate_for_copying_weights_to_target = 0.005
        self.use_deterministic_for_evaluation = True


class SoftActorCriticAgentParameters(AgentParameters):
    def __init__(self):
        super().__init__(algorithm=SoftActorCriticAlgorithmParameters(), memory=ExperienceReplayParameters(),
                         networks={"policy": SACNetworkParameters(), "q": SACNetworkParameters(),
                                   "v": SACNetworkParameters()})
        # only the state-value network has a target copy (soft_actor_critic_agent.py:46)
        self.network_wrappers["v"].create_target_network = True

    @property
    def path(self):
        return 'coach_b200.agents.soft_actor_critic_agent:SoftActorCriticAgent'


class _QHeadBinding(object):
    """one of the two Q heads of SACQHead on fixed (obs, action) input buffers"""

    def __init__(self, agent, k, obs, act, train, need_action_grad, dq_da, accumulate_da):
        lib, ws, B, H = agent.lib, agent.ws, agent.B, agent.H
        st = agent.q.store
        theta, grad = st.theta, st.grad
        dev = agent.device
        self.agent, self.B, self.H = agent, B, H
        want_bwd = train or need_action_grad
        self.obs_emb = agent.q_obs[k].instantiate(lib, ws, B, obs, theta, grad, train=train)
        self.act_emb = agent.q_act[k].instantiate(lib, ws, B, act, theta, grad, train=want_bwd,
                                                  need_input_grad=need_action_grad, input_act=0, dx_in=dq_da,
                                                  dx_accumulate=accumulate_da)
        self.e = torch.zeros((B, H), dtype=torch.float32, device=dev)
        self.d_e = torch.zeros((B, H), dtype=torch.float32, device=dev) if want_bwd else None
        self.tail = agent.q_tail[k].instantiate(lib, ws, B, self.e, theta, grad, train=want_bwd,
                                                need_input_grad=want_bwd, input_act=0, dx_in=self.d_e)
        self.q = self.tail.out
        self.train = train

    def forward(self):
        lib, st, B, H = self.agent.lib, _lib.current_stream(), self.B, self.H
        eo = self.obs_emb.forward()
        ea = self.act_emb.forward()
        _lib.check(lib.cb200_axpby_2d(eo.data_ptr(), H, B, H, 1.0, 0.0, self.e.data_ptr(), H, st))    # e = eo
        _lib.check(lib.cb200_axpby_2d(ea.data_ptr(), H, B, H, 1.0, 1.0, self.e.data_ptr(), H, st))    # e += ea
        return self.tail.forward()

    def backward(self, weights=True):
        """expects d(loss)/dq in tail.d_out"""
        lib, st, B, H = self.agent.lib, _lib.current_stream(), self.B, self.H
        self.tail.backward(weights)
        _lib.check(lib.cb200_act_backward(self.d_e.data_ptr(), H, self.act_emb.out.data_ptr(), H, B, H, RELU,
                                          self.act_emb.d_out.data_ptr(), H, st))
        self.act_emb.backward(weights)
        if weights:
            _lib.check(lib.cb200_act_backward(self.d_e.data_ptr(), H, self.obs_emb.out.data_ptr(), H, B, H, RELU,
                                              self.obs_emb.d_out.data_ptr(), H, st))
            self.obs_emb.backward()


class SoftActorCriticAgent(object):
    def __init__(self, agent_parameters, parent=None, observation_dim=None, action_dim=None, device=None, seed=None):
        self.ap = agent_parameters
        self.lib = _lib.load()
        self.device = dev = torch.device(device if device is not None else "cuda")
        self.D, self.A = D, A = int(observation_dim), int(action_dim)
        pp, pq, pv = (self.ap.network_wrappers[k] for k in ("policy", "q", "v"))
        self.B = B = int(pq.batch_size)
        self.H = H = int(getattr(pq, "hidden_units", 256))
        self.memory = dynamic_import_and_instantiate_module_from_params(self.ap.memory, extra_kwargs={"device": dev})
        self.ws = Workspace(dev)
        # ---- layouts (TF creation order inside each network) ----
        sp = ParamStore(dev)
        self.policy_seq = Sequential([Dense(D, H, "relu"), Dense(H, H, "relu"), Dense(H, 2 * A, None)], sp,
                                     "policy/online/network_0")
        sp.add("policy/online/network_0/gradients_from_head_0-0_rescalers", ())
        sp.finalize()
        sq = ParamStore(dev)
        self.q_obs, self.q_act, self.q_tail = [], [], []
        for k in range(2):
            pre = "q/online/network_0/sac_q_head_0/q%d_head" % (k + 1)
            self.q_obs.append(Sequential([Dense(D, H, "relu")], sq, pre + "/obs"))
            self.q_act.append(Sequential([Dense(A, H, "relu")], sq, pre + "/act"))
            self.q_tail.append(Sequential([Dense(H, H, "relu"), Dense(H, 1, None)], sq, pre + "/tail"))
        sq.add("q/online/network_0/gradients_from_head_0-0_rescalers", ())
        sq.finalize()
        sv = ParamStore(dev)
        self.v_seq = Sequential([Dense(D, H, "relu"), Dense(H, H, "relu"), Dense(H, 1, None)], sv,
                                "v/online/network_0")
        sv.add("v/online/network_0/gradients_from_head_0-0_rescalers", ())
        sv.finalize()
        gen = torch.Generator().manual_seed(int(seed)) if seed is not None else None
        for s in (sp, sq, sv):
            s.init_glorot(gen)
        self.policy, self.q, self.v = _Net(self.lib, sp, pp, dev), _Net(self.lib, sq, pq, dev), _Net(self.lib, sv, pv, dev)
        self.v.sync()
        # ---- buffers and bindings ----
        f32 = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)      # noqa: E731
        self.batch_buffers = {"state:observation": f32(B, D), "next_state:observation": f32(B, D),
                              "action": f32(B, A), "reward": torch.zeros(B, dtype=torch.float64, device=dev),
                              "game_over": torch.zeros(B, dtype=torch.uint8, device=dev)}
        if hasattr(self.memory, "declare_schema") and self.memory.ring.specs is None:
            self.memory.declare_schema(self.batch_buffers)       # store(Transition) casts gym's float64 to these dtypes
        s, s2 = self.batch_buffers["state:observation"], self.batch_buffers["next_state:observation"]
        self.policy_inst = self.policy_seq.instantiate(self.lib, self.ws, B, s, sp.theta, sp.grad, train=True)
        self.sampled = f32(B, A)
        self.logp = f32(B)
        self.dq_da = f32(B, A)
        self.eps = [f32(B, A) for _ in range(3)]
        # Q on (s, a~): forward + data gradient wrt the action; Q on (s, a_batch): training
        self.q_pi = [_QHeadBinding(self, k, s, self.sampled, False, True, self.dq_da, k > 0) for k in range(2)]
        self.q_train = [_QHeadBinding(self, k, s, self.batch_buffers["action"], True, False, None, False)
                        for k in range(2)]
        self.v_train = self.v_seq.instantiate(self.lib, self.ws, B, s, sv.theta, sv.grad, train=True)
        self.v_target_s2 = self.v_seq.instantiate(self.lib, self.ws, B, s2, self.v.target)
        self.log_target = f32(B)
        self.v_targets = f32(B, 1)
        self.td_targets = f32(B, 1)
        self.v_loss, self.q_loss = f32(1), [f32(1), f32(1)]
        self.training_iteration = 0
        self.total_steps_counter = 0
        self.last_target_network_update_step = 0
        self._graph_step = None

    @property
    def is_on_policy(self) -> bool:
        return False

    def learn_from_batch(self, batch, fetch=True, noise=None):
        """``noise``: optional three [B, A] arrays standing in for the three independent samples TensorFlow draws in
        the three policy-network runs; by default drawn from numpy's global generator."""
        B, A = self.B, self.A
        cols = batch.columns
        for k, buf in self.batch_buffers.items():          # the kernels (and their CUDA graph) read the agent's buffers
            if k in cols and cols[k].data_ptr() != buf.data_ptr():
                buf.copy_(cols[k].reshape(buf.shape))
        if noise is None:
            noise = [np.random.standard_normal((B, A)) for _ in range(3)]
        for e, n in zip(self.eps, noise):
            e.copy_(torch.as_tensor(np.asarray(n), dtype=torch.float32))
        if self._graph_step is None:
            self._graph_step = GraphedKernels(self._sac_kernels, self.device)
        self._graph_step()
        total = self.q_loss[0] + self.q_loss[1]
        if fetch:
            l = float(total.item())
            return l, [l], float(torch.sqrt(self.q.sumsq).item())
        return total, [total], self.q.sumsq

    def _sac_kernels(self):
        lib, st, B, A = self.lib, _lib.current_stream(), self.B, self.A
        cols = self.batch_buffers
        # 1. policy forward + sample (eps1)
        z = self.policy_inst.forward()
        _lib.check(lib.cb200_sac_policy_sample(z.data_ptr(), self.eps[0].data_ptr(), B, A, None,
                                               self.sampled.data_ptr(), self.logp.data_ptr(), st))
        # 2. Q(s, a~1): min and d mean(min) / da
        q1 = self.q_pi[0].forward()
        q2 = self.q_pi[1].forward()
        _lib.check(lib.cb200_sac_min_seed(q1.data_ptr(), q2.data_ptr(), B, self.q_pi[0].tail.d_out.data_ptr(),
                                          self.q_pi[1].tail.d_out.data_ptr(), self.log_target.data_ptr(), st))
        self.q_pi[0].backward(weights=False)       # writes dq_da
        self.q_pi[1].backward(weights=False)       # accumulates into dq_da
        # 3. policy gradient and step
        _lib.check(lib.cb200_sac_policy_grad(z.data_ptr(), self.eps[1].data_ptr(), self.eps[2].data_ptr(),
                                             self.dq_da.data_ptr(), B, A, self.policy_inst.d_out.data_ptr(), st))
        self.policy_inst.backward()
        self.policy.apply(self.ws)
        # 4. V network: targets = log_target - log pi(a~1)   (soft_actor_critic_agent.py:244-249)
        _lib.check(lib.cb200_sub(self.log_target.data_ptr(), self.logp.data_ptr(), B, self.v_targets.data_ptr(), st))
        v = self.v_train.forward()
        _lib.check(lib.cb200_regression_head_loss_grad(v.data_ptr(), self.v_targets.data_ptr(), None, B, 1, 0, 1.0,
                                                       self.v_train.d_out.data_ptr(), self.v_loss.data_ptr(), st))
        self.v_train.backward()
        self.v.apply(self.ws)
        # 5. Q networks: y = r + (1 - done) * gamma * V_target(s')   (:265-269); loss 0.5 * mean((Q_k - y)^2) each
        v_next = self.v_target_s2.forward()
        _lib.check(lib.cb200_ac_td_targets(cols["reward"].data_ptr(), cols["game_over"].data_ptr(), v_next.data_ptr(),
                                           1, B, float(self.ap.algorithm.discount), 0, 0, 0.0, 0.0,
                                           self.td_targets.data_ptr(), st))
        for k in range(2):
            q = self.q_train[k].forward()
            _lib.check(lib.cb200_regression_head_loss_grad(q.data_ptr(), self.td_targets.data_ptr(), None, B, 1, 0,
                                                           0.5, self.q_train[k].tail.d_out.data_ptr(),
                                                           self.q_loss[k].data_ptr(), st))
            self.q_train[k].backward()
        self.q.apply(self.ws)

    def sample_batch(self):
        return self.memory.sample_batch(self.B, out=self.batch_buffers)

    def train(self, fetch=True):
        loss = 0
        if self.memory.num_transitions() < 1:
            return loss
        for _ in range(self.ap.algorithm.num_consecutive_training_steps):
            self.training_iteration += 1
            batch = self.sample_batch()
            total_loss, _, _ = self.learn_from_batch(batch, fetch=fetch)
            loss = loss + total_loss if fetch else total_loss
            steps = self.ap.algorithm.num_steps_between_copying_online_weights_to_target.num_steps
            if (self.total_steps_counter - self.last_target_network_update_step) >= steps:
                self.last_target_network_update_step = self.total_steps_counter
                self.v.sync(self.ap.algorithm.rate_for_copying_weights_to_target)
        return loss
