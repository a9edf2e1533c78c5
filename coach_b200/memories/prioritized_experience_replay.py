"""Proportional prioritized experience replay with HBM-resident fp64 sum / min / max segment trees.  Drop-in for
``rl_coach/memories/non_episodic/prioritized_experience_replay.py:159-300``.

What is bit-exact with the reference (given the same stores, uniform draws and error values):
  * the three tree arrays after every store / update_priorities (``priority_mode='libm'``, the default),
  * the sampled leaf indices,
  * ``beta`` (same LinearSchedule recurrence), ``num_transitions()`` including the double-append quirk of
    ``store`` (:271 and :280 both append, so the count is min(2*stores, size)).
The importance weights use the device ``pow`` (<= 2 ulp in fp64; identical after the float32 cast the network feeds
on).  ``priority_mode='device'`` computes ``(err+eps)**alpha`` on the GPU as well (fully asynchronous; ~0.1% of
leaves then differ from glibc's ``pow`` in the last bit).

Host <-> device traffic per training step: 4 KB of uniforms down; in 'libm' mode 4 KB of TD errors up and 8 KB of
priorities down, overlapped by the agent with the network's backward pass.
"""
import ctypes
import random
from typing import List, Tuple

import numpy as np
import torch

from coach_b200 import _lib
from coach_b200.core_types import DeviceBatch, Transition
from coach_b200.memories.experience_replay import ExperienceReplay, ExperienceReplayParameters
from coach_b200.memories.memory import MemoryGranularity
from coach_b200.schedules import ConstantSchedule, Schedule


class PrioritizedExperienceReplayParameters(ExperienceReplayParameters):
    def __init__(self):
        super().__init__()
        self.max_size = (MemoryGranularity.Transitions, 1000000)
        self.alpha = 0.6
        self.beta = ConstantSchedule(0.4)
        self.epsilon = 1e-6

    @property
    def path(self):
        return 'coach_b200.memories.prioritized_experience_replay:PrioritizedExperienceReplay'


class _LazyColumns(object):
    """materialises an un-staged batch column from the drawn slots (DeviceBatch.column)"""

    def __init__(self, ring, idx, names):
        self.ring, self.idx, self.names = ring, idx, names

    def __call__(self, name):
        if name not in self.names:
            raise KeyError(name)
        return self.ring.gather_column(name, self.idx.clone())


class PrioritizedExperienceReplay(ExperienceReplay):
    def __init__(self, max_size: Tuple[MemoryGranularity, int], alpha: float = 0.6,
                 beta: Schedule = ConstantSchedule(0.4), epsilon: float = 1e-6,
                 allow_duplicates_in_batch_sampling: bool = True, device=None, priority_mode: str = "libm",
                 frame_dedup: bool = False, frame_slack: float = 0.25, frame_streams: int = 1):
        if max_size[0] != MemoryGranularity.Transitions:
            raise ValueError("Prioritized Experience Replay currently only support setting the memory size in "
                             "transitions granularity.")
        if priority_mode not in ("libm", "device"):
            raise ValueError("priority_mode must be 'libm' or 'device'")
        self.power_of_2_size = 1
        while self.power_of_2_size < max_size[1]:
            self.power_of_2_size *= 2                                                   # :176-178
        super().__init__((MemoryGranularity.Transitions, self.power_of_2_size), allow_duplicates_in_batch_sampling,
                         device=device, frame_dedup=frame_dedup, frame_slack=frame_slack, frame_streams=frame_streams)
        self.alpha = alpha
        self.beta = beta
        self.epsilon = epsilon
        self.priority_mode = priority_mode
        n = self.power_of_2_size
        dev = self.device
        self.sum_tree = torch.empty(2 * n - 1, dtype=torch.float64, device=dev)
        self.min_tree = torch.empty(2 * n - 1, dtype=torch.float64, device=dev)
        self.max_tree = torch.empty(2 * n - 1, dtype=torch.float64, device=dev)
        self._winner = torch.empty(n, dtype=torch.int32, device=dev)
        self._maxp_dev = torch.ones(1, dtype=torch.float64, device=dev)
        self._neg_flag = torch.zeros(1, dtype=torch.int32, device=dev)
        self._init_trees()
        self._maximal_priority = 1.0
        self._maxp_stale = False
        self._maxp_host, self._maxp_event = None, None   # pinned mirror of the max-tree root behind every update
        self._list_len = 0                  # len(self.transitions) of the reference (doubled, capped)
        self._u_ring = {}                   # batch size -> rotating pinned / device buffers of the uniform draws
        self._flag_host, self._flag_event = None, None
        self._update_side = None            # layers.SideStream an owner runs update_priorities_device on (DQNAgent)

    def _init_trees(self):
        _lib.check(self.lib.cb200_per_init(self.sum_tree.data_ptr(), self.min_tree.data_ptr(),
                                           self.max_tree.data_ptr(), self._winner.data_ptr(), self.power_of_2_size,
                                           _lib.current_stream()))

    # ---- reference-visible state ---------------------------------------------------------------------------------
    def _join_update(self):
        """a tree update an owner queued on its side stream is ordered before whatever the current stream does next"""
        if self._update_side is not None:
            self._update_side.join()

    @property
    def maximal_priority(self) -> float:
        if self._maxp_stale:
            if self._maxp_event is not None:
                # mirrored behind the update kernel on ITS stream: wait for that copy only, not for the learn step
                self._maxp_event.synchronize()
                self._maximal_priority = float(self._maxp_host[0])
            else:
                self._join_update()
                self._maximal_priority = float(self._maxp_dev.item())  # 8-byte D2H, only when a store needs it
            self._maxp_stale = False
        return self._maximal_priority

    @maximal_priority.setter
    def maximal_priority(self, v):
        self._maximal_priority = float(v)
        self._maxp_stale = False

    def num_transitions(self) -> int:
        return min(self._list_len + 2 * self.ring._pending, self.power_of_2_size)

    # ---- store ---------------------------------------------------------------------------------------------------
    def store(self, transition: Transition, lock=True) -> None:
        self.assert_not_frozen()
        if self.memory_backend:
            self.memory_backend.store(transition)
        if self.ring.stage_transition(transition):
            self._flush()

    def _flush(self):
        self._join_update()
        first, n = self.ring.flush()
        self._tree_store(first, n)
        return first, n

    def _tree_store(self, first, n):
        if n == 0:
            return
        p = self.maximal_priority                                                        # :276
        _lib.check(self.lib.cb200_per_store(self.sum_tree.data_ptr(), self.min_tree.data_ptr(),
                                            self.max_tree.data_ptr(), self._winner.data_ptr(), self.power_of_2_size,
                                            first, n, float(p ** self.alpha), float(p), _lib.current_stream()))
        self._list_len = min(self._list_len + 2 * n, self.power_of_2_size)              # :271 + :280

    def store_columns(self, columns: dict) -> None:
        self.assert_not_frozen()
        self._flush()
        first, n = self.ring.append_columns(columns)
        self._tree_store(first, n)

    # ---- priorities ----------------------------------------------------------------------------------------------
    def update_priorities(self, indices, error_values) -> None:
        """:203-217.  ``indices`` / ``error_values``: lists, numpy arrays or CUDA tensors."""
        if len(indices) != len(error_values):
            raise ValueError("The number of indexes requested for update don't match the number of error values given")
        n = len(indices)
        if n == 0:
            return
        self._flush()
        idx = self._as_device(indices, torch.int64, check_range=True)
        if self.priority_mode == "libm":
            if torch.is_tensor(error_values):
                err = error_values.detach().to(torch.float64).cpu().numpy()
            else:
                err = np.ascontiguousarray(error_values, dtype=np.float64)
            p_alpha_h, p_raw_h = self.host_priorities(err)
            p_alpha = torch.from_numpy(p_alpha_h).to(self.device, non_blocking=True)
            p_raw = torch.from_numpy(p_raw_h).to(self.device, non_blocking=True)
        else:
            err = self._as_device(error_values, torch.float64)
            p_alpha = torch.empty(n, dtype=torch.float64, device=self.device)
            p_raw = torch.empty(n, dtype=torch.float64, device=self.device)
            _lib.check(self.lib.cb200_per_priorities_device(err.data_ptr(), n, float(self.epsilon), float(self.alpha),
                                                            p_alpha.data_ptr(), p_raw.data_ptr(),
                                                            self._neg_flag.data_ptr(), _lib.current_stream()))
        self.update_priorities_device(idx, p_alpha, p_raw)

    def host_priorities(self, err: np.ndarray):
        """(err + eps) ** alpha with the host libm -- what ``priority ** self.alpha`` (:198) evaluates to."""
        err = np.ascontiguousarray(err, dtype=np.float64)
        p_alpha = np.empty_like(err)
        p_raw = np.empty_like(err)
        _lib.check(self.lib.cb200_host_priorities(err.ctypes.data, err.size, float(self.epsilon), float(self.alpha),
                                                  p_alpha.ctypes.data, p_raw.ctypes.data))
        return p_alpha, p_raw

    def update_priorities_device(self, idx, p_alpha, p_raw) -> None:
        """Tree update from device-resident leaf indices and ready-made priorities (p_alpha -> sum & min trees,
        p_raw -> max tree)."""
        _lib.check(self.lib.cb200_per_update(self.sum_tree.data_ptr(), self.min_tree.data_ptr(),
                                             self.max_tree.data_ptr(), self._winner.data_ptr(), self.power_of_2_size,
                                             idx.data_ptr(), p_alpha.data_ptr(), p_raw.data_ptr(), idx.shape[0],
                                             self._maxp_dev.data_ptr(), self._neg_flag.data_ptr(),
                                             _lib.current_stream()))
        if self.device.type == "cuda":
            if self._maxp_host is None:
                self._maxp_host = torch.zeros(1, dtype=torch.float64, pin_memory=True)
            self._maxp_host.copy_(self._maxp_dev, non_blocking=True)
            self._maxp_event = torch.cuda.Event()
            self._maxp_event.record()
        self._maxp_stale = True
        self._post_flag_check()

    # ---- device-side error flags (priority_mode='device', CUDA-tensor indices) ---------------------------------------
    # Invalid entries never reach the trees (the kernels skip them); the condition is reported like the reference
    # reports it -- a ValueError -- at the next host-side call that can observe it, without ever blocking the stream:
    # the flag word is copied to a pinned mirror behind every update and read once the copy's event has completed.
    def _post_flag_check(self):
        if self.device.type != "cuda":
            return
        if self._flag_host is None:
            self._flag_host = torch.zeros(1, dtype=torch.int32, pin_memory=True)
        self._flag_host.copy_(self._neg_flag, non_blocking=True)
        self._flag_event = torch.cuda.Event()
        self._flag_event.record()

    def check_device_errors(self, wait=False):
        """raises the reference's ValueError if a device-side priority update met a negative error or an out-of-range
        leaf since the last check (wait=True synchronises with the last update first)"""
        if self._flag_event is None:
            return
        if wait:
            self._flag_event.synchronize()
        if not self._flag_event.query():
            return
        self._flag_event = None
        flags = int(self._flag_host[0])
        if flags:
            self._neg_flag.zero_()
            if flags & 1:
                raise ValueError("The priorities must be non-negative values")                       # :195
            raise ValueError("The given left index can not be found in the tree. The available leaves are: 0-{}"
                             .format(self.power_of_2_size - 1))                                       # :124-126

    def _as_device(self, v, dtype, check_range=False):
        if torch.is_tensor(v):
            return v.to(device=self.device, dtype=dtype).contiguous()
        a = np.ascontiguousarray(v, dtype=np.int64 if dtype == torch.int64 else np.float64)
        if check_range and a.size and (a.min() < 0 or a.max() >= self.power_of_2_size):
            bad = int(a[(a < 0) | (a >= self.power_of_2_size)][0])
            raise ValueError("The given left index ({}) can not be found in the tree. The available leaves are: 0-{}"
                             .format(bad, self.power_of_2_size - 1))                     # :124-126
        return torch.from_numpy(a).to(self.device)

    # ---- sample --------------------------------------------------------------------------------------------------
    def sample_batch(self, size: int, out: dict = None, uniforms=None, s2d: dict = None) -> DeviceBatch:
        """:219-262 as ONE fused kernel launch (tree descent + importance weights + column gather).  ``uniforms``
        (optional) are the raw ``random.random()`` draws; by default they are drawn here from Python's global
        generator, one per sample, exactly the stream ``random.uniform`` consumes at :244.
        ``s2d`` ({"columns": {ring column: bf16 plane}, "geometry": (H, W, C, S)}): the image columns leave the kernel
        as the space-to-depth operand planes of the first convolution instead of a staged uint8 copy
        (cb200_per_sample_gather_s2d); the batch materialises them on demand (``DeviceBatch.column``)."""
        if not self.num_transitions() >= size:
            raise ValueError("The replay buffer cannot be sampled since there are not enough transitions yet. "
                             "There are currently {} transitions".format(self.num_transitions()))
        self._flush()
        self.check_device_errors()
        if uniforms is None:
            rnd = random.random
            uniforms = [rnd() for _ in range(size)]
        u = self._upload_uniforms(uniforms, size)
        if out is None:
            out = self.ring.alloc_batch(size)
        for k, dt in (("idx", torch.int64), ("weight", torch.float64), ("weight32", torch.float32)):
            if k not in out:
                out[k] = torch.empty(size, dtype=dt, device=self.device)
        ke = getattr(self, "kernel_events", None)      # (start, end) CUDA events recorded right around the launch
        if s2d is not None:
            ia, ni, sa, ns = self.ring.s2d_tables(s2d, out, size)
            H, W, C, S = s2d["geometry"]
            if ke:
                ke[0].record()
            _lib.check(self.lib.cb200_per_sample_gather_s2d(
                self.sum_tree.data_ptr(), self.min_tree.data_ptr(), self.power_of_2_size, u.data_ptr(), size,
                self.num_transitions(), float(self.beta.current_value), out["idx"].data_ptr(),
                out["weight"].data_ptr(), out["weight32"].data_ptr(), ia, ni, H, W, C, S, sa, ns,
                self.ring.frames_ptr(), self.ring.frame_capacity, _lib.current_stream()))
            if ke:
                ke[1].record()
            self.beta.step()
            cols = {k: v for k, v in out.items() if k not in s2d["columns"]}
            return DeviceBatch(cols, size, lazy=_LazyColumns(self.ring, cols["idx"], tuple(s2d["columns"])))
        arr, cnt = self.ring.column_table(out, size)
        if ke:
            ke[0].record()
        _lib.check(self.lib.cb200_per_sample_gather(
            self.sum_tree.data_ptr(), self.min_tree.data_ptr(), self.power_of_2_size, u.data_ptr(), size,
            self.num_transitions(), float(self.beta.current_value), out["idx"].data_ptr(), out["weight"].data_ptr(),
            out["weight32"].data_ptr(), arr, cnt, _lib.current_stream()))
        self.ring.assemble_stacks(out, out["idx"], size)       # frame-deduplicated ring: stacks built from the frame store
        if ke:
            ke[1].record()
        self.beta.step()                                                                 # :255
        return DeviceBatch(dict(out), size)

    def _upload_uniforms(self, uniforms, size):
        """host draws -> device through a small ring of persistent pinned buffers (no allocation per call; a buffer
        is reused only after the copy that read it has completed)"""
        if self.device.type != "cuda":
            return torch.tensor(uniforms, dtype=torch.float64)
        ring = self._u_ring.get(size)
        if ring is None:
            ring = self._u_ring[size] = {"i": 0, "slots": [
                (torch.empty(size, dtype=torch.float64, pin_memory=True),
                 torch.empty(size, dtype=torch.float64, device=self.device), [None]) for _ in range(4)]}
        host, dev, ev = ring["slots"][ring["i"]]
        ring["i"] = (ring["i"] + 1) % len(ring["slots"])
        if ev[0] is not None:
            ev[0].synchronize()
        host.numpy()[:] = uniforms
        dev.copy_(host, non_blocking=True)
        ev[0] = torch.cuda.Event()
        ev[0].record()
        return dev

    def sample_indices(self, size: int, uniforms=None):
        """Descent + weights only (no gather): returns (idx int64, weight float64) CUDA tensors."""
        if not self.num_transitions() >= size:
            raise ValueError("The replay buffer cannot be sampled since there are not enough transitions yet. "
                             "There are currently {} transitions".format(self.num_transitions()))
        self._flush()
        if uniforms is None:
            uniforms = [random.random() for _ in range(size)]
        u = torch.tensor(uniforms, dtype=torch.float64).to(self.device)
        idx = torch.empty(size, dtype=torch.int64, device=self.device)
        w = torch.empty(size, dtype=torch.float64, device=self.device)
        _lib.check(self.lib.cb200_per_sample(self.sum_tree.data_ptr(), self.min_tree.data_ptr(), self.power_of_2_size,
                                             u.data_ptr(), size, self.num_transitions(),
                                             float(self.beta.current_value), idx.data_ptr(), w.data_ptr(), None,
                                             _lib.current_stream()))
        self.beta.step()
        return idx, w

    def sample(self, size: int) -> List[Transition]:
        return self.sample_batch(size).to_transitions()

    def _draw_positions(self, size):       # uniform sampling is not how a PER is read
        raise NotImplementedError

    # ---- misc ----------------------------------------------------------------------------------------------------
    def clean(self, lock=True) -> None:
        self.assert_not_frozen()
        self._join_update()
        self.ring.clear()
        self._list_len = 0
        self._init_trees()                                                               # :294-296

    def get_transition(self, transition_index: int, lock: bool = True):
        raise NotImplementedError("list-position access is not meaningful for the prioritized ring; use the leaf "
                                  "indices returned by sample")
