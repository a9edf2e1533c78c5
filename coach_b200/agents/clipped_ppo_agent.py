"""Clipped PPO learn step on the GPU.  Drop-in for

  rl_coach/agents/clipped_ppo_agent.py:157-207   fill_advantages   (V(s) over the rollout, GAE per episode, standardise)
  rl_coach/agents/clipped_ppo_agent.py:209-308   train_network     (epochs x minibatches: old policy from the frozen
                                                                    target network, surrogate + value losses, Adam)
  rl_coach/agents/clipped_ppo_agent.py:314-344   train             (normalise observations, sync target, truncate to
                                                                    num_consecutive_playing_steps, shuffle, epochs)
  rl_coach/agents/actor_critic_agent.py:108-125  GAE

Network (presets/Mujoco_ClippedPPO.py:30-37, use_separate_networks_per_head): value net obs->64->64->1 and policy net
obs->64->64->A, tanh, plus the state-independent ``policy_log_std`` variable; variables in TF creation order inside
ONE flat buffer, so both sub-networks share a single gradient norm / Adam / all-reduce launch.

The minibatch step (row gather at a device-side cursor -> value fwd/bwd -> policy fwd -> PPO head -> policy bwd ->
global norm -> Adam with device-side step state -> cursor += B) has launch parameters that never change, so it is
captured once in a CUDA graph and replayed for every minibatch of every epoch: 0 host work per minibatch.
The old policy's mean is evaluated once per training phase for the whole rollout (the target network is frozen during
``train_network`` -- the reference recomputes identical values every minibatch, clipped_ppo_agent.py:238-240 TODO-perf).
"""
import random

import numpy as np
import torch

from coach_b200 import _lib, parallel, rl_math
from coach_b200.architectures.layers import Dense, Workspace
from coach_b200.architectures.network import ParamStore, Sequential
from coach_b200.base_parameters import AgentParameters, AlgorithmParameters, EnvironmentSteps, NetworkParameters
from coach_b200.core_types import DeviceBatch
from coach_b200.filters.filter import InputFilter, ObservationNormalizationFilter
from coach_b200.memories.episodic_experience_replay import EpisodicExperienceReplayParameters
from coach_b200.schedules import ConstantSchedule
from coach_b200.utils import dynamic_import_and_instantiate_module_from_params


class ClippedPPONetworkParameters(NetworkParameters):
    def __init__(self):
        super().__init__()
        self.batch_size = 64
        self.optimizer_type = 'Adam'
        self.clip_gradients = None
        self.use_separate_networks_per_head = True
        self.create_target_network = True
        # (learning rate, Adam epsilon / beta2 keep the NetworkParameters defaults like the reference class does; the
        # Mujoco preset sets 3e-4 / 1e-5 / 0.999: coach_b200/presets/Mujoco_ClippedPPO.py)
        self.hidden_units = 64


class ClippedPPOAlgorithmParameters(AlgorithmParameters):
    def __init__(self):
        super().__init__()
        self.gae_lambda = 0.95
        self.clip_likelihood_ratio_using_epsilon = 0.2
        self.estimate_state_value_using_gae = True
        self.beta_entropy = 0.01  # should be 0 for mujoco
        self.num_consecutive_playing_steps = EnvironmentSteps(2048)
        self.optimization_epochs = 10
        self.clipping_decay_schedule = ConstantSchedule(1)
        self.act_for_full_episodes = True
        self.update_pre_network_filters_state_on_train = True
        self.update_pre_network_filters_state_on_inference = False
        # the reference trains on dataset[:num_consecutive_playing_steps] only (clipped_ppo_agent.py:330-331); set to
        # False to train on the whole rollout (the 64-env synthetic configuration of BASELINE config 3)
        self.truncate_dataset_to_playing_steps = True


class ClippedPPOAgentParameters(AgentParameters):
    def __init__(self):
        super().__init__(algorithm=ClippedPPOAlgorithmParameters(), memory=EpisodicExperienceReplayParameters(),
                         networks={"main": ClippedPPONetworkParameters()})
        self.pre_network_filter = InputFilter()
        self.pre_network_filter.add_observation_filter('observation', 'normalize_observation',
                                                       ObservationNormalizationFilter(name='normalize_observation'))

    @property
    def path(self):
        return 'coach_b200.agents.clipped_ppo_agent:ClippedPPOAgent'


class PPONetworkDef(object):
    """flat parameter layout of the two sub-networks (general_network.py:244-349 creation order)"""

    def __init__(self, device, obs_dim, action_dim, hidden=64):
        self.device = torch.device(device)
        self.D, self.A, self.Hd = int(obs_dim), int(action_dim), int(hidden)
        s = self.store = ParamStore(self.device)
        self.v_seq = Sequential([Dense(self.D, hidden, "tanh"), Dense(hidden, hidden, "tanh"), Dense(hidden, 1, None)],
                                s, "main/online/network_0")
        s.add("main/online/network_0/gradients_from_head_0-0_rescalers", ())
        self.p_seq = Sequential([Dense(self.D, hidden, "tanh"), Dense(hidden, hidden, "tanh"),
                                 Dense(hidden, self.A, None)], s, "main/online/network_1")
        self.logstd_name = s.add("main/online/network_1/ppo_head_0/policy_log_std", (self.A,))
        s.add("main/online/network_1/gradients_from_head_1-0_rescalers", ())
        s.finalize()

    def init(self, generator=None):
        self.store.init_glorot(generator)
        self.store.view(self.store.theta, self.logstd_name).zero_()          # np.zeros((1, num_actions)), :129-133
        # policy mean layer: normalized_columns_initializer(0.01) (ppo_head.py:121, head.py:28-33)
        name = self.p_seq.names[2][0]
        w = torch.randn(self.Hd, self.A, generator=generator)
        w *= 0.01 / torch.sqrt((w * w).sum(dim=0, keepdim=True))
        self.store.view(self.store.theta, name).copy_(w)


class ClippedPPOAgent(object):
    def __init__(self, agent_parameters, parent=None, observation_dim=None, action_dim=None, device=None, seed=None):
        self.ap = agent_parameters
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else "cuda")
        self.D, self.A = int(observation_dim), int(action_dim)
        net_p = self.ap.network_wrappers["main"]
        self.B = int(net_p.batch_size)
        self.memory = dynamic_import_and_instantiate_module_from_params(
            self.ap.memory, extra_kwargs={"device": self.device, "discount": self.ap.algorithm.discount})
        self.pre_network_filter = self.ap.pre_network_filter
        if self.pre_network_filter is not None:
            self.pre_network_filter.set_device(self.device)
            for flt in self.pre_network_filter._observation_filters.values():
                for f in flt.values():
                    if hasattr(f, "set_shape"):
                        f.set_shape([self.D])
        self.net = PPONetworkDef(self.device, self.D, self.A, getattr(net_p, "hidden_units", 64))
        gen = torch.Generator().manual_seed(int(seed)) if seed is not None else None
        self.net.init(gen)
        st = self.net.store
        self.theta_target = st.new_buffer()
        self.ws = Workspace(self.device)
        dev, B = self.device, self.B
        f32 = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)     # noqa: E731
        # persistent minibatch buffers
        self.mb = dict(states=f32(B, self.D), actions=f32(B, self.A), advantages=f32(B), value_targets=f32(B, 1),
                       old_mu=f32(B, self.A))
        self.v_inst = self.net.v_seq.instantiate(self.lib, self.ws, B, self.mb["states"], st.theta, st.grad,
                                                 train=True)
        self.p_inst = self.net.p_seq.instantiate(self.lib, self.ws, B, self.mb["states"], st.theta, st.grad,
                                                 train=True)
        self.scalars = f32(5)
        self.v_loss = f32(1)
        self.sumsq = f32(1)
        self.adam_state = torch.tensor([net_p.adam_optimizer_beta1, net_p.adam_optimizer_beta2], dtype=torch.float32,
                                       device=dev)
        self.cursor = torch.zeros(1, dtype=torch.int64, device=dev)
        self._graph = None
        self._graph_key = None
        self._full = {}            # rollout-sized forward instances, keyed by N
        self.training_iteration = 0
        self.total_steps_counter = 0
        self.last_training_phase_step = 0
        self.use_cuda_graph = True
        self.last_losses = None

    @property
    def is_on_policy(self) -> bool:
        return True

    def sync(self):
        """online -> target: the frozen "old policy" (network_wrapper.py:94-107, clipped_ppo_agent.py:326)"""
        _lib.check(self.lib.cb200_polyak(self.theta_target.data_ptr(), self.net.store.theta.data_ptr(),
                                         self.net.store.size, 1.0, _lib.current_stream()))

    # ---- rollout-sized forward passes ------------------------------------------------------------------------------
    def _full_instances(self, N):
        if N not in self._full:
            st = self.net.store
            x = torch.zeros((N, self.D), dtype=torch.float32, device=self.device)
            v = self.net.v_seq.instantiate(self.lib, self.ws, N, x, st.theta)
            p_old = self.net.p_seq.instantiate(self.lib, self.ws, N, x, self.theta_target)
            self._full = {N: (x, v, p_old)}           # keep only the latest size
        return self._full[N]

    def fill_advantages(self, states_norm, rewards, game_overs):
        """clipped_ppo_agent.py:157-207.  Returns (advantages fp64 standardised, value targets fp64, n_valid)."""
        N = states_norm.shape[0]
        x, v_full, _ = self._full_instances(N)
        x.copy_(states_norm)
        values = v_full.forward().reshape(-1)                                   # V(s_t), fp32
        return rl_math.fill_advantages(rewards, values, game_overs, self.ap.algorithm.discount,
                                       self.ap.algorithm.gae_lambda)

    # ---- one minibatch ---------------------------------------------------------------------------------------------
    def _minibatch_kernels(self, data, perm, n_rows):
        """all launches of one minibatch step; parameters independent of the minibatch index"""
        lib, st = self.lib, _lib.current_stream()
        store = self.net.store
        alg, net_p = self.ap.algorithm, self.ap.network_wrappers["main"]
        arr, cnt = _lib.make_columns([(data[k].data_ptr(), self.mb[k].data_ptr(),
                                       self.mb[k].element_size() * int(np.prod(self.mb[k].shape[1:])))
                                      for k in ("states", "actions", "advantages", "value_targets", "old_mu")])
        _lib.check(lib.cb200_gather_at(arr, cnt, perm.data_ptr(), self.cursor.data_ptr(), self.B, st))
        v = self.v_inst.forward()
        mu = self.p_inst.forward()
        # VHead: MSE(v, gae_based_value_target), loss weight 1 (v_head.py:41-44, head.py:172-177)
        _lib.check(lib.cb200_regression_head_loss_grad(v.data_ptr(), self.mb["value_targets"].data_ptr(), None,
                                                       self.B, 1, 0, 1.0, self.v_inst.d_out.data_ptr(),
                                                       self.v_loss.data_ptr(), st))
        clip_eps = float(alg.clip_likelihood_ratio_using_epsilon) * float(alg.clipping_decay_schedule.current_value)
        logstd = store.view(store.theta, self.net.logstd_name)
        old_logstd = store.view(self.theta_target, self.net.logstd_name)
        d_logstd = store.view(store.grad, self.net.logstd_name)
        _lib.check(lib.cb200_ppo_continuous_head(mu.data_ptr(), logstd.data_ptr(), self.mb["actions"].data_ptr(),
                                                 self.mb["old_mu"].data_ptr(), old_logstd.data_ptr(),
                                                 self.mb["advantages"].data_ptr(), self.B, self.A, clip_eps,
                                                 float(alg.beta_entropy), self.p_inst.d_out.data_ptr(),
                                                 d_logstd.data_ptr(), self.scalars.data_ptr(), st))
        self.v_inst.backward()
        self.p_inst.backward()
        n = store.size
        _lib.check(lib.cb200_sumsq(store.grad.data_ptr(), n, self.sumsq.data_ptr(), self.ws.ptr(), st))
        if net_p.clip_gradients:
            _lib.check(lib.cb200_clip_by_global_norm(store.grad.data_ptr(), n, self.sumsq.data_ptr(),
                                                     float(net_p.clip_gradients), st))
        scaler = parallel.allreduce_gradients(store.grad,
                                              net_p.scale_down_gradients_by_number_of_workers_for_sync_training)
        if scaler != 1.0:
            _lib.check(lib.cb200_scale(store.grad.data_ptr(), n, float(scaler), st))
        _lib.check(lib.cb200_adam_tf_dev(store.theta.data_ptr(), store.m.data_ptr(), store.v.data_ptr(),
                                         store.grad.data_ptr(), n, float(net_p.learning_rate),
                                         float(net_p.adam_optimizer_beta1), float(net_p.adam_optimizer_beta2),
                                         float(net_p.optimizer_epsilon), self.adam_state.data_ptr(), st))
        _lib.check(lib.cb200_add_i64(self.cursor.data_ptr(), self.B, st))

    def train_network(self, data, n_rows, epochs):
        """clipped_ppo_agent.py:209-308.  data: dict of rollout-sized CUDA tensors (states, actions, advantages,
        value_targets, old_mu).  Returns the mean [value loss, policy loss] of the last epoch as device tensors."""
        B = self.B
        n_full = n_rows // B
        if n_rows % B:
            raise ValueError("the rollout length (%d) must be a multiple of the batch size (%d)" % (n_rows, B))
        perm_host = torch.zeros(n_rows, dtype=torch.int64, pin_memory=self.device.type == "cuda")
        perm = torch.zeros(n_rows, dtype=torch.int64, device=self.device)
        key = (tuple(t.data_ptr() for t in data.values()), perm.data_ptr(), n_rows,
               float(self.ap.algorithm.clipping_decay_schedule.current_value))
        world = parallel.world()[1]
        graphable = self.use_cuda_graph and self.device.type == "cuda" and world == 1
        if graphable:
            # warm-up launch outside capture (lazy module loading), then capture once
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._snapshot_then_restore(lambda: self._minibatch_kernels(data, perm, n_rows))
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._minibatch_kernels(data, perm, n_rows)
            self._restore_snapshot()
        order = list(range(n_rows))
        v_acc = torch.zeros(1, dtype=torch.float32, device=self.device)
        p_acc = torch.zeros(1, dtype=torch.float32, device=self.device)
        for epoch in range(epochs):
            random.shuffle(order)                                   # batch.shuffle(), core_types.py:452-468
            perm_host.copy_(torch.tensor(order, dtype=torch.int64))
            perm.copy_(perm_host, non_blocking=True)
            self.cursor.zero_()
            v_acc.zero_()
            p_acc.zero_()
            for i in range(n_full):
                if graphable:
                    graph.replay()
                else:
                    self._minibatch_kernels(data, perm, n_rows)
                v_acc += self.v_loss
                p_acc += self.scalars[0:1]
        self.last_losses = (v_acc / n_full, p_acc / n_full)
        return self.last_losses

    # snapshot / restore of everything a warm-up or capture pass mutates (weights, Adam slots + state, cursor)
    def _snapshot_then_restore(self, fn):
        self._take_snapshot()
        fn()
        self._restore_snapshot()

    def _take_snapshot(self):
        s = self.net.store
        self._snap = (s.theta.clone(), s.m.clone(), s.v.clone(), self.adam_state.clone(), self.cursor.clone())

    def _restore_snapshot(self):
        s = self.net.store
        th, m, v, ad, cur = self._snap
        s.theta.copy_(th)
        s.m.copy_(m)
        s.v.copy_(v)
        self.adam_state.copy_(ad)
        self.cursor.copy_(cur)

    # ---- driver ----------------------------------------------------------------------------------------------------
    def _should_train(self):
        steps = self.ap.algorithm.num_consecutive_playing_steps
        should = (self.total_steps_counter - self.last_training_phase_step) >= steps.num_steps
        should = should and self.memory.num_transitions_in_complete_episodes() > 0
        if should:
            self.last_training_phase_step = self.total_steps_counter
        return should

    def train(self):
        """clipped_ppo_agent.py:314-344"""
        if not self._should_train():
            return None
        alg = self.ap.algorithm
        batch = self.memory.transitions_batch()
        if self.pre_network_filter is not None:
            batch = self.pre_network_filter.filter(batch, deep_copy=False,
                                                   update_internal_state=alg.update_pre_network_filters_state_on_train)
        states = batch.states(["observation"])["observation"].to(torch.float32).contiguous()
        actions = batch.actions().to(torch.float32).reshape(batch.size, self.A).contiguous()
        for _ in range(alg.num_consecutive_training_steps):
            self.sync()
            adv, tgt, n_valid = self.fill_advantages(states, batch.rewards(), batch.game_overs())
            n_rows = batch.size
            if alg.truncate_dataset_to_playing_steps:
                n_rows = min(n_rows, alg.num_consecutive_playing_steps.num_steps)
            if n_rows < self.B:
                raise ValueError("the rollout holds %d transitions, fewer than one minibatch of %d: nothing to train on "
                                 "(the reference would train on one partial minibatch)" % (n_rows, self.B))
            # whole minibatches only: the reference also trains on the partial tail batch (clipped_ppo_agent.py:225,
            # ceil(size / batch_size)); with the presets' 2048-step rollouts and batch 64 there is none
            n_rows = (n_rows // self.B) * self.B
            _, _, p_old = self._full_instances(batch.size)
            old_mu = p_old.forward()                                   # frozen target network, whole rollout at once
            f32 = lambda t: t.to(torch.float32).contiguous()           # noqa: E731
            data = dict(states=states, actions=actions, advantages=f32(adv), value_targets=f32(tgt).reshape(-1, 1),
                        old_mu=old_mu)
            self.train_network(data, n_rows, alg.optimization_epochs)
        self.memory.clean()                                            # post_training_commands :310-312
        self.training_iteration += 1
        return None
