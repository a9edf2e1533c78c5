"""Filters on the replay -> learn path, mirroring the reference's interfaces:

  InputFilter                      <- rl_coach/filters/filter.py:230-430 (the subset used at agent.py:735 and
                                      clipped_ppo_agent.py:321: ``filter(batch, update_internal_state, deep_copy)``)
  ObservationStackingFilter/LazyStack <- filters/observation/observation_stacking_filter.py:27-115 (host, observe time)
  ObservationNormalizationFilter   <- filters/observation/observation_normalization_filter.py:28-89 with the running
                                      statistics of utilities/shared_running_stats.py:115-164 held ON DEVICE
  RewardClippingFilter / RewardRescaleFilter / ObservationToUInt8Filter
                                   <- filters/reward/reward_clipping_filter.py:24-55, reward_rescale_filter.py:24-45,
                                      filters/observation/observation_to_uint8_filter.py:27-65

``InputFilter.filter`` accepts either the reference's ``List[Transition]`` / ``EnvResponse``-like objects (host path,
same element-wise semantics) or a :class:`~coach_b200.core_types.DeviceBatch` (device path: observation filters that
``support_device_batches`` are applied to the ``state:<key>`` / ``next_state:<key>`` columns in HBM).
"""
import copy
from collections import OrderedDict, deque

import numpy as np
import torch

from coach_b200 import _lib
from coach_b200.core_types import DeviceBatch


class Filter(object):
    def __init__(self, name=None):
        self.name = name

    def reset(self):
        pass

    def filter(self, data, update_internal_state=True):
        raise NotImplementedError("")

    def set_device(self, device, memory_backend_params=None, mode='numpy'):
        pass

    def set_session(self, sess):
        pass


class ObservationFilter(Filter):
    supports_batching = False
    supports_device_batches = False

    def get_filtered_observation_space(self, input_observation_space):
        return input_observation_space

    def validate_input_observation_space(self, input_observation_space):
        pass


class RewardFilter(Filter):
    supports_batching = False


# ---- observation filters ---------------------------------------------------------------------------------------------
class LazyStack(object):
    """np.stack deferred until the array is needed (frames are shared between neighbouring transitions)."""

    def __init__(self, history, axis=None):
        self.history = copy.copy(history)
        self.axis = axis

    def __array__(self, dtype=None, copy=None):
        array = np.stack(self.history, axis=self.axis)
        if dtype is not None:
            array = array.astype(dtype)
        return array


class ObservationStackingFilter(ObservationFilter):
    def __init__(self, stack_size: int, stacking_axis: int = -1):
        super().__init__()
        if stack_size <= 0:
            raise ValueError("The stack shape must be a positive number")
        if type(stack_size) != int:
            raise ValueError("The stack shape must be of int type")
        self.stack_size = stack_size
        self.stacking_axis = stacking_axis
        self.stack = []
        self.flatten_vectors = False        # VectorObservationSpace inputs are flattened eagerly (:97-99)

    def filter(self, observation, update_internal_state: bool = True):
        if len(self.stack) == 0:
            self.stack = deque([observation] * self.stack_size, maxlen=self.stack_size)     # first frame replicated
        elif update_internal_state:
            self.stack.append(observation)
        observation = LazyStack(self.stack, self.stacking_axis)
        if self.flatten_vectors:
            observation = np.array(observation).flatten()
        return observation

    def reset(self) -> None:
        self.stack = []


class DeviceRunningStats(object):
    """NumpySharedRunningStats (shared_running_stats.py:115-164) with sum / sum of squares / mean / std kept as fp64
    CUDA tensors; the (fractional) count stays on the host like the reference's ``_count``."""

    def __init__(self, device, epsilon=1e-2):
        self.device = torch.device(device)
        self.lib = _lib.load()
        self.epsilon = epsilon
        self._count = epsilon
        self.shape = None

    def set_params(self, shape, clip_values=None):
        self.shape = tuple(shape)
        d = int(np.prod(self.shape))
        dev = self.device
        self._sum = torch.zeros(d, dtype=torch.float64, device=dev)
        self._sum_squares = torch.full((d,), self.epsilon, dtype=torch.float64, device=dev)
        self._mean = torch.zeros(d, dtype=torch.float64, device=dev)
        self._std = torch.full((d,), float(np.sqrt(self.epsilon)), dtype=torch.float64, device=dev)
        self.clip_values = clip_values

    @property
    def n(self):
        return self._count

    @property
    def mean(self):
        return self._mean

    @property
    def std(self):
        return self._std

    def push(self, samples: torch.Tensor):
        """samples: fp32 CUDA tensor [rows, *shape].  With several ranks the statistics are SHARED (the reference's
        SharedRunningStats): every rank pushes its own rollout shard and the increments are merged by one all-reduce of
        (sum, sum of squares, rows) before mean / std are refreshed, so all ranks normalise with the same numbers."""
        from coach_b200 import parallel
        x = samples.reshape(samples.shape[0], -1).contiguous()
        st = _lib.current_stream()
        rows = x.shape[0]
        if parallel.is_distributed():
            d_sum, d_sq = torch.zeros_like(self._sum), torch.zeros_like(self._sum_squares)
            _lib.check(self.lib.cb200_running_stats_push(x.data_ptr(), rows, x.shape[1], d_sum.data_ptr(),
                                                         d_sq.data_ptr(), st))
            rows = parallel.allreduce_running_stats(d_sum, d_sq, rows)
            self._sum += d_sum
            self._sum_squares += d_sq
        else:
            _lib.check(self.lib.cb200_running_stats_push(x.data_ptr(), rows, x.shape[1], self._sum.data_ptr(),
                                                         self._sum_squares.data_ptr(), st))
        self._count += rows
        _lib.check(self.lib.cb200_running_stats_finalize(self._sum.data_ptr(), self._sum_squares.data_ptr(),
                                                         float(self._count), float(self.epsilon), x.shape[1],
                                                         self._mean.data_ptr(), self._std.data_ptr(), st))

    checkpoint_file_extension = 'srs'

    def save_state_to_checkpoint(self, checkpoint_dir: str, checkpoint_prefix):
        """utilities/shared_running_stats.py:170-179: the reference's pickle (same keys, same file name)"""
        from coach_b200 import checkpoint
        checkpoint.save_running_stats(self, checkpoint_dir, checkpoint_prefix, self.checkpoint_file_extension)

    def restore_state_from_checkpoint(self, checkpoint_dir: str, checkpoint_prefix):
        from coach_b200 import checkpoint
        checkpoint.restore_running_stats(self, checkpoint_dir, checkpoint_prefix, self.checkpoint_file_extension)

    def normalize(self, batch: torch.Tensor, out=None):
        x = batch.reshape(batch.shape[0], -1).contiguous()
        if out is None:
            out = torch.empty_like(x, dtype=torch.float32)
        lo, hi = self.clip_values
        _lib.check(self.lib.cb200_running_stats_normalize(x.data_ptr(), x.shape[0], x.shape[1], self._mean.data_ptr(),
                                                          self._std.data_ptr(), float(lo), float(hi), out.data_ptr(),
                                                          None, _lib.current_stream()))
        return out.view(batch.shape)


class ObservationNormalizationFilter(ObservationFilter):
    supports_batching = True
    supports_device_batches = True

    def __init__(self, clip_min: float = -5.0, clip_max: float = 5.0, name='observation_stats'):
        super().__init__()
        self.clip_min = clip_min
        self.clip_max = clip_max
        self.running_observation_stats = None
        self.name = name
        self.observation_space = None

    def set_device(self, device, memory_backend_params=None, mode='numpy') -> None:
        self.running_observation_stats = DeviceRunningStats(device)

    def set_shape(self, shape):
        """get_filtered_observation_space (:80-83) without the spaces object"""
        self.running_observation_stats.set_params(shape=shape, clip_values=(self.clip_min, self.clip_max))

    def filter(self, observations, update_internal_state: bool = True):
        """observations: CUDA tensor [rows, *shape] (or anything np.array() accepts: copied to the device)."""
        if not torch.is_tensor(observations):
            observations = torch.as_tensor(np.array(observations), dtype=torch.float32)
        x = observations.to(self.running_observation_stats.device, dtype=torch.float32)
        if update_internal_state:
            self.running_observation_stats.push(x)
            self.last_mean = self.running_observation_stats.mean
            self.last_stdev = self.running_observation_stats.std
        return self.running_observation_stats.normalize(x)


class ObservationToUInt8Filter(ObservationFilter):
    def __init__(self, input_low: float, input_high: float):
        super().__init__()
        if input_high <= input_low:
            raise ValueError("The input observation space high values can be less or equal to the input observation "
                             "space low values")
        self.input_low, self.input_high = input_low, input_high

    def filter(self, observation, update_internal_state: bool = True):
        observation = np.asarray(observation)
        observation = (observation - self.input_low) / (self.input_high - self.input_low) * 255
        return observation.astype('uint8')           # truncation, as in the reference


# ---- reward filters --------------------------------------------------------------------------------------------------
class RewardClippingFilter(RewardFilter):
    def __init__(self, clipping_low: float = -np.inf, clipping_high: float = np.inf):
        super().__init__()
        if clipping_low > clipping_high:
            raise ValueError("The reward clipping low must be lower than the reward clipping max")
        self.clipping_low, self.clipping_high = clipping_low, clipping_high

    def filter(self, reward, update_internal_state: bool = True):
        reward = float(reward)
        if self.clipping_high:                       # truthiness test: a bound of 0 is ignored (reference quirk Q13)
            reward = min(reward, self.clipping_high)
        if self.clipping_low:
            reward = max(reward, self.clipping_low)
        return reward


class RewardRescaleFilter(RewardFilter):
    def __init__(self, rescale_factor: float):
        super().__init__()
        if rescale_factor == 0:
            raise ValueError("The reward rescale value can not be set to 0")
        self.rescale_factor = rescale_factor

    def filter(self, reward, update_internal_state: bool = True):
        return float(reward) * self.rescale_factor


# ---- container -------------------------------------------------------------------------------------------------------
class InputFilter(object):
    def __init__(self, observation_filters=None, reward_filters=None, name='input_filter'):
        self.name = name
        self._observation_filters = OrderedDict()     # observation name -> OrderedDict(filter name -> filter)
        self._reward_filters = OrderedDict()
        for obs_name, flt in (observation_filters or {}).items():
            for fname, f in flt.items():
                self.add_observation_filter(obs_name, fname, f)
        for fname, f in (reward_filters or {}).items():
            self.add_reward_filter(fname, f)

    def add_observation_filter(self, observation_name, filter_name, filter, add_as_the_first_filter=False):
        d = self._observation_filters.setdefault(observation_name, OrderedDict())
        d[filter_name] = filter
        if add_as_the_first_filter:
            d.move_to_end(filter_name, last=False)

    def add_reward_filter(self, filter_name, filter, add_as_the_first_filter=False):
        self._reward_filters[filter_name] = filter
        if add_as_the_first_filter:
            self._reward_filters.move_to_end(filter_name, last=False)

    def set_device(self, device, memory_backend_params=None, mode='numpy'):
        for flt in self._observation_filters.values():
            for f in flt.values():
                f.set_device(device, memory_backend_params, mode)
        for f in self._reward_filters.values():
            f.set_device(device, memory_backend_params, mode)

    def save_state_to_checkpoint(self, checkpoint_dir, checkpoint_prefix):
        """filters/filter.py:452-465 of the reference: every filter with state saves it under its own prefix"""
        for obs_name, flts in self._observation_filters.items():
            for fname, f in flts.items():
                st = getattr(f, "running_observation_stats", None)
                if st is not None and getattr(st, "shape", None) is not None:
                    st.save_state_to_checkpoint(checkpoint_dir, "%s.%s.%s" % (checkpoint_prefix, obs_name, fname))

    def restore_state_from_checkpoint(self, checkpoint_dir, checkpoint_prefix):
        for obs_name, flts in self._observation_filters.items():
            for fname, f in flts.items():
                st = getattr(f, "running_observation_stats", None)
                if st is not None and getattr(st, "shape", None) is not None:
                    st.restore_state_from_checkpoint(checkpoint_dir, "%s.%s.%s" % (checkpoint_prefix, obs_name, fname))

    def reset(self):
        for flt in self._observation_filters.values():
            for f in flt.values():
                f.reset()
        for f in self._reward_filters.values():
            f.reset()

    def filter(self, unfiltered_data, update_internal_state: bool = True, deep_copy: bool = True):
        """filter.py:295-350.  DeviceBatch in -> DeviceBatch out (columns replaced by filtered CUDA tensors)."""
        if isinstance(unfiltered_data, DeviceBatch):
            return self._filter_device_batch(unfiltered_data, update_internal_state)
        is_list = isinstance(unfiltered_data, list)
        data = copy.deepcopy(unfiltered_data) if deep_copy else [copy.copy(t) for t in unfiltered_data]
        data = data if isinstance(data, list) else [data]
        # Transitions: all states through all filters, then all next states; environment responses (no ``state``) have
        # their next_state filtered -- the traversal order of filter.py:314-334, which stateful filters (frame
        # stacking) observe
        if hasattr(data[0], "state") and data[0].state is not None and hasattr(data[0], "next_state") and \
                type(data[0]).__name__ != "EnvResponse":
            state_lists = [[t.state for t in data], [t.next_state for t in data]]
        else:
            state_lists = [[t.next_state for t in data]]
        for states in state_lists:
            for obs_name, flt in self._observation_filters.items():
                if obs_name not in states[0]:
                    continue
                for f in flt.values():
                    vals = [st[obs_name] for st in states]
                    if f.supports_batching:
                        out = f.filter(vals, update_internal_state=update_internal_state)
                        out = out.cpu().numpy() if torch.is_tensor(out) else out
                    else:
                        out = [f.filter(v, update_internal_state=update_internal_state) for v in vals]
                    for st, v in zip(states, out):
                        st[obs_name] = v
        for f in self._reward_filters.values():
            for t in data:
                t.reward = f.filter(t.reward, update_internal_state=update_internal_state)
        return data          # always a list, like the reference (force_list, filter.py:311)

    def _filter_device_batch(self, batch, update_internal_state):
        cols = dict(batch.columns)
        for prefix in ("state:", "next_state:"):            # all states first, then all next states (filter.py:314-334)
            for obs_name, flt in self._observation_filters.items():
                key = prefix + obs_name
                if key not in cols:
                    continue
                for f in flt.values():
                    if not f.supports_device_batches:
                        raise ValueError("filter %s cannot run on device-resident batches" % type(f).__name__)
                    cols[key] = f.filter(cols[key], update_internal_state=update_internal_state)
        if self._reward_filters:
            raise ValueError("reward filters run at store time on the host (they are scalar, per transition)")
        return DeviceBatch(cols, batch.size)


class NoInputFilter(InputFilter):
    pass
