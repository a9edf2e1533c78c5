"""Q-value networks (DQN / DDQN / dueling) as flat-buffer device networks.

Structure and variable order follow the TF graph of the reference:
  embedder  -- image: Conv2d(32,8,4) Conv2d(64,4,2) Conv2d(64,3,1), ReLU, flatten (image_embedder.py:62-67, Medium);
               vector: Dense(256) ReLU (vector_embedder.py:58-61, Medium); input / 255 for images (embedder.py:103)
  middleware-- Dense(512) ReLU (fc_middleware.py:66-69, Medium)
  head      -- QHead: Dense(num_actions) (q_head.py:52-54);
               DuelingQHead: V: Dense(512) ReLU, Dense(1); A: Dense(512) ReLU, Dense(num_actions);
               Q = V + (A - mean_a A) (dueling_q_head.py:33-47)
  + one scalar ``gradients_from_head_0-0_rescalers`` variable (general_network.py:312-315; gradient identically 0)
"""
import numpy as np
import torch

from coach_b200 import _lib
from coach_b200.architectures import tiled as tl
from coach_b200.architectures.layers import Conv2d, Dense, Workspace
from coach_b200.architectures.network import ParamStore, Sequential, make_u8_lut


class QNetworkDef(object):
    """Parameter layout + layer chain; instances bind it to buffers (see QNetworkInstance)."""

    def __init__(self, device, observation_shape, num_actions, dueling=False, embedder="auto", middleware_units=512):
        self.device = torch.device(device)
        self.obs_shape = tuple(observation_shape)
        self.num_actions = int(num_actions)
        self.dueling = bool(dueling)
        self.store = ParamStore(self.device)
        self.is_image = len(self.obs_shape) == 3
        layers = []
        if self.is_image:
            h, w, c = self.obs_shape
            for (n, k, s) in ((32, 8, 4), (64, 4, 2), (64, 3, 1)):
                conv = Conv2d((h, w), c, n, k, s, "relu")
                layers.append(conv)
                h, w, c = conv.OH, conv.OW, n
            flat = h * w * c
        else:
            layers.append(Dense(self.obs_shape[0], 256, "relu"))
            flat = 256
        self.n_embedder = len(layers)
        layers.append(Dense(flat, middleware_units, "relu"))
        if not self.dueling:
            layers.append(Dense(middleware_units, self.num_actions, None))
            self.trunk = Sequential(layers, self.store, "main/online/network_0")
            self.v_tower = self.a_tower = None
        else:
            self.trunk = Sequential(layers, self.store, "main/online/network_0")
            self.v_tower = Sequential([Dense(middleware_units, 512, "relu"), Dense(512, 1, None)], self.store,
                                      "main/online/network_0/dueling_q_values_head_0/state_value")
            self.a_tower = Sequential([Dense(middleware_units, 512, "relu"), Dense(512, self.num_actions, None)],
                                      self.store, "main/online/network_0/dueling_q_values_head_0/action_advantage")
        self.store.add("main/online/network_0/gradients_from_head_0-0_rescalers", ())
        self.store.finalize()
        self.lut = make_u8_lut(self.device) if self.is_image else None

    def instantiate(self, lib, ws, B, x, theta, grad=None, train=False):
        return QNetworkInstance(self, lib, ws, B, x, theta, grad, train)


class QNetworkInstance(object):
    def __init__(self, net, lib, ws, B, x, theta, grad, train):
        self.net, self.lib, self.B = net, lib, B
        dev = net.device
        x_is_u8 = x.dtype == torch.uint8
        # Large batches run the trunk on pre-split bf16 operands (architectures/tiled.py).  The parameter planes are
        # re-derived from theta at the start of every forward (one launch), so any writer of theta -- Adam, a target
        # network copy, polyak, a checkpoint load -- is covered.
        # (CB200_GEMM_TILED=0 keeps every layer on the gather-GEMM of cb200_gemm: A/B runs, bench.py --no-tc)
        self.theta_planes = tl.ThetaPlanes(lib, net.store, theta) \
            if (B >= 128 and B % 32 == 0 and _lib.tune_default("gemm_tiled", 1)) else None
        self.trunk = net.trunk.instantiate(lib, ws, B, x, theta, grad, x_is_u8=x_is_u8, lut=net.lut, train=train,
                                           theta_planes=self.theta_planes)
        if not net.dueling:
            self.q = self.trunk.out
            self.dq = self.trunk.d_out
            return
        h = self.trunk.out                      # middleware output (post-ReLU)
        dh = self.trunk.d_out                   # gradient wrt its pre-activation
        relu = 1
        self.v = net.v_tower.instantiate(lib, ws, B, h, theta, grad, need_input_grad=train, input_act=relu,
                                         train=train, dx_in=dh, dx_accumulate=False)
        self.a = net.a_tower.instantiate(lib, ws, B, h, theta, grad, need_input_grad=train, input_act=relu,
                                         train=train, dx_in=dh, dx_accumulate=True)
        self.q = torch.empty((B, net.num_actions), dtype=torch.float32, device=dev)
        self.dq = torch.empty_like(self.q) if train else None

    def forward(self):
        if self.theta_planes is not None:
            self.theta_planes.refresh()
        self.trunk.forward()
        if self.net.dueling:
            self.v.forward()
            self.a.forward()
            _lib.check(self.lib.cb200_dueling_combine_fwd(self.v.out.data_ptr(), self.a.out.data_ptr(), self.B,
                                                          self.net.num_actions, self.q.data_ptr(),
                                                          _lib.current_stream()))
        return self.q

    def backward_top(self):
        """plain Q head only: the dense layers (middleware + head) of the trunk, which hold ~95 % of the parameters;
        their gradients are complete -- and can be all-reduced -- while ``backward_bottom`` still runs"""
        assert not self.net.dueling
        n = len(self.trunk.layers)
        self.trunk.backward(layers=(self.net.n_embedder, n))

    def backward_bottom(self):
        self.trunk.backward(layers=(0, self.net.n_embedder))

    def grad_split_offset(self):
        """element offset in the flat gradient buffer where the top (dense) layers' gradients start"""
        return self.net.store.entries[self.net.trunk.names[self.net.n_embedder][0]][0]

    def backward(self):
        """expects d(loss)/dq in self.dq; leaves all parameter gradients in the grad buffer"""
        if self.net.dueling:
            _lib.check(self.lib.cb200_dueling_combine_bwd(self.dq.data_ptr(), self.B, self.net.num_actions,
                                                          self.v.d_out.data_ptr(), self.a.d_out.data_ptr(),
                                                          _lib.current_stream()))
            self.v.backward()      # writes d(middleware pre-activation)
            self.a.backward()      # accumulates into it
        self.trunk.backward()
