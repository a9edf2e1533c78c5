"""CPU restatement of the reference's replay memories.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows
  rl_coach/memories/non_episodic/experience_replay.py:41-150            (ExperienceReplay)
  rl_coach/memories/non_episodic/prioritized_experience_replay.py:43-300 (SegmentTree, PrioritizedExperienceReplay)
  rl_coach/core_types.py:488-623                                         (Batch column extraction)
  rl_coach/schedules.py:40-63                                            (LinearSchedule)

Two execution modes for the tree arithmetic:
  * ``backend='python'`` -- per-sample Python loops, the way the reference itself runs (one interpreter thread).
    This is what ``bench.py`` times as the ``"port"`` CPU baseline.
  * ``backend='c'``      -- the same arithmetic in plain C (oracle/segment_tree.c), used by the tests so that
    2^20-leaf cases finish in seconds.  Both are checked against fixtures written by the imported reference
    (tests/golden/, oracle/make_golden.py) in tests/test_oracle_golden.py, and against each other.
"""
import ctypes
import os
import random

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def clib():
    """Loads (building on first use if gcc is around) oracle/_build/liboracle.so."""
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.check_call(["make", "-C", _HERE, "-s"])
        lib = ctypes.CDLL(path)
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int64)
        lib.ost_init.argtypes = [dp, ctypes.c_int64, ctypes.c_int]
        lib.ost_update.argtypes = [dp, ctypes.c_int64, ctypes.c_int, ctypes.c_int64, ctypes.c_double]
        lib.ost_retrieve.argtypes = [dp, ctypes.c_int64, ctypes.c_double]
        lib.ost_retrieve.restype = ctypes.c_int64
        lib.oper_update_priorities.argtypes = [dp, dp, dp, ctypes.c_int64, ip, dp, ctypes.c_int64,
                                               ctypes.c_double, ctypes.c_double, dp]
        lib.oper_update_priorities.restype = ctypes.c_int
        lib.oper_store.argtypes = [dp, dp, dp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                   ctypes.c_double, ctypes.c_double]
        lib.oper_store.restype = ctypes.c_int64
        lib.oper_sample.argtypes = [dp, dp, ctypes.c_int64, ctypes.c_int64, dp, ctypes.c_int64, ctypes.c_double,
                                    ip, dp, dp]
        _LIB = lib
    return _LIB


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))


# ----------------------------------------------------------------------------------------------------------------
# schedules.py:40-63
class OracleLinearSchedule:
    """LinearSchedule: repeated subtraction + np.clip (accumulating fp error is part of the behaviour, quirk Q5)."""

    def __init__(self, initial_value, final_value, decay_steps):
        self.initial_value = initial_value
        self.current_value = initial_value
        self.final_value = final_value
        self.decay_steps = decay_steps
        self.decay_delta = (initial_value - final_value) / float(decay_steps)

    def step(self):
        self.current_value -= self.decay_delta
        if self.final_value < self.initial_value:
            self.current_value = np.clip(self.current_value, self.final_value, self.initial_value)
        if self.final_value > self.initial_value:
            self.current_value = np.clip(self.current_value, self.initial_value, self.final_value)


class OracleConstantSchedule:
    def __init__(self, v):
        self.initial_value = v
        self.current_value = v

    def step(self):
        pass


# ----------------------------------------------------------------------------------------------------------------
_OPS = {"sum": (0, 0.0), "min": (1, float("inf")), "max": (2, -float("inf"))}


class OracleSegmentTree:
    """prioritized_experience_replay.py:43-156, iterative instead of recursive, same fp operations."""

    def __init__(self, size, op, backend="python"):
        if not (size > 0 and size & (size - 1) == 0):
            raise ValueError("A segment tree size must be a positive power of 2. The given size is {}".format(size))
        self.size = size
        self.op = op
        self.opcode, init = _OPS[op]
        self.tree = np.full(2 * size - 1, init, dtype=np.float64)
        self.next_leaf_idx_to_write = 0
        self.backend = backend

    def _combine(self, a, b):
        if self.op == "sum":
            return a + b
        if self.op == "min":
            return b if b < a else a
        return b if b > a else a

    def update(self, leaf_idx, new_val):
        node = leaf_idx + self.size - 1
        if not 0 <= node < len(self.tree):
            raise ValueError("The given left index ({}) can not be found in the tree. The available leaves are: 0-{}"
                             .format(leaf_idx, self.size - 1))
        if self.backend == "c":
            clib().ost_update(_dp(self.tree), self.size, self.opcode, int(leaf_idx), float(new_val))
            return
        t = self.tree
        t[node] = new_val
        while node != 0:
            parent = (node - 1) // 2
            t[parent] = self._combine(t[2 * parent + 1], t[2 * parent + 2])
            node = parent

    def add(self, val):
        self.update(self.next_leaf_idx_to_write, val)
        self.next_leaf_idx_to_write += 1
        if self.next_leaf_idx_to_write >= self.size:
            self.next_leaf_idx_to_write = 0

    def total_value(self):
        return self.tree[0]

    def retrieve(self, val):
        """Returns (leaf_idx, leaf_value) -- get_element_by_partial_sum :131-146 without the data object."""
        if self.backend == "c":
            leaf = clib().ost_retrieve(_dp(self.tree), self.size, float(val))
            return leaf, self.tree[leaf + self.size - 1]
        t = self.tree
        n = len(t)
        node = 0
        while True:
            left = 2 * node + 1
            if left >= n:
                break
            if val <= t[left]:
                node = left
            else:
                val = val - t[left]
                node = left + 1
        return node - self.size + 1, t[node]

    def levels_str(self):
        """SegmentTree.__str__ :148-156."""
        out, start, width = "", 0, 1
        while width <= self.size:
            out += "{}\n".format(self.tree[start:start + width])
            start += width
            width *= 2
        return out


# ----------------------------------------------------------------------------------------------------------------
class OracleExperienceReplay:
    """experience_replay.py:41-150.  Transitions are stored as opaque python objects (any tuple/dict)."""

    def __init__(self, max_size, allow_duplicates_in_batch_sampling=True):
        self.max_size = max_size
        self.transitions = []
        self.allow_duplicates_in_batch_sampling = allow_duplicates_in_batch_sampling

    def num_transitions(self):
        return len(self.transitions)

    def store(self, transition):
        self.transitions.append(transition)
        while self.max_size != 0 and len(self.transitions) > self.max_size:
            del self.transitions[0]

    def sample_indices(self, size):
        """experience_replay.py:80-88 -- legacy global numpy RandomState, exactly as the reference draws."""
        if self.allow_duplicates_in_batch_sampling:
            return np.random.randint(self.num_transitions(), size=size)
        if self.num_transitions() >= size:
            return np.random.choice(self.num_transitions(), size=size, replace=False)
        raise ValueError("The replay buffer cannot be sampled since there are not enough transitions yet. "
                         "There are currently {} transitions".format(self.num_transitions()))

    def sample(self, size):
        idx = self.sample_indices(size)
        return [self.transitions[i] for i in idx]


class OraclePrioritizedExperienceReplay:
    """prioritized_experience_replay.py:159-300, including the double-append quirk of ``store`` (Q1):
    ``num_transitions()`` == min(2 * stores, size)."""

    def __init__(self, max_size, alpha=0.6, beta=None, epsilon=1e-6, backend="python"):
        self.power_of_2_size = 1
        while self.power_of_2_size < max_size:
            self.power_of_2_size *= 2
        n = self.power_of_2_size
        self.backend = backend
        self.sum_tree = OracleSegmentTree(n, "sum", backend)
        self.min_tree = OracleSegmentTree(n, "min", backend)
        self.max_tree = OracleSegmentTree(n, "max", backend)
        self.data = [None] * n
        self.alpha = alpha
        self.beta = beta if beta is not None else OracleConstantSchedule(0.4)
        self.epsilon = epsilon
        self.maximal_priority = 1.0
        self._list_len = 0          # len(self.transitions) of the reference (grows by 2 per store, capped)

    def num_transitions(self):
        return self._list_len

    def store(self, transition):
        # :271 super().store -> ExperienceReplay.store appends + enforces; :280 appends + enforces again
        self._list_len = min(self._list_len + 1, self.power_of_2_size)
        p = self.maximal_priority
        self.data[self.sum_tree.next_leaf_idx_to_write] = transition
        self.sum_tree.add(p ** self.alpha)
        self.min_tree.add(p ** self.alpha)
        self.max_tree.add(p)
        self._list_len = min(self._list_len + 1, self.power_of_2_size)

    def store_many(self, transitions):
        """n consecutive stores; C fast path (all stores in one call use the same maximal_priority, which
        ``store`` never changes)."""
        n = len(transitions)
        cur = self.sum_tree.next_leaf_idx_to_write
        for k, t in enumerate(transitions):
            self.data[(cur + k) % self.power_of_2_size] = t
        if self.backend == "c":
            new_cur = clib().oper_store(_dp(self.sum_tree.tree), _dp(self.min_tree.tree), _dp(self.max_tree.tree),
                                        self.power_of_2_size, cur, n, float(self.maximal_priority), float(self.alpha))
            for tr in (self.sum_tree, self.min_tree, self.max_tree):
                tr.next_leaf_idx_to_write = new_cur
        else:
            for _ in range(n):
                p = self.maximal_priority
                self.sum_tree.add(p ** self.alpha)
                self.min_tree.add(p ** self.alpha)
                self.max_tree.add(p)
        self._list_len = min(self._list_len + 2 * n, self.power_of_2_size)

    def update_priorities(self, indices, error_values):
        if len(indices) != len(error_values):
            raise ValueError("The number of indexes requested for update don't match the number of error values given")
        if self.backend == "c":
            idx = np.ascontiguousarray(indices, dtype=np.int64)
            err = np.ascontiguousarray(error_values, dtype=np.float64)
            mp = ctypes.c_double(self.maximal_priority)
            rc = clib().oper_update_priorities(_dp(self.sum_tree.tree), _dp(self.min_tree.tree),
                                               _dp(self.max_tree.tree), self.power_of_2_size, _ip(idx), _dp(err),
                                               len(idx), float(self.epsilon), float(self.alpha), ctypes.byref(mp))
            self.maximal_priority = mp.value
            if rc != 0:
                raise ValueError("The priorities must be non-negative values")
            return
        for leaf_idx, error in zip(indices, error_values):
            if error < 0:
                raise ValueError("The priorities must be non-negative values")
            priority = error + self.epsilon
            self.sum_tree.update(leaf_idx, priority ** self.alpha)
            self.min_tree.update(leaf_idx, priority ** self.alpha)
            self.max_tree.update(leaf_idx, priority)
            self.maximal_priority = self.max_tree.total_value()

    def sample_indices(self, size, uniforms=None):
        """Returns (leaf indices int64[size], normalised IS weights float64[size]).  ``uniforms`` = the raw
        ``random.random()`` draws; when None they are drawn here from the global ``random`` module, one per sample,
        which consumes the generator exactly like ``random.uniform`` at :244."""
        if not self.num_transitions() >= size:
            raise ValueError("The replay buffer cannot be sampled since there are not enough transitions yet. "
                             "There are currently {} transitions".format(self.num_transitions()))
        if uniforms is None:
            uniforms = [random.random() for _ in range(size)]
        beta = float(self.beta.current_value)
        nt = self.num_transitions()
        if self.backend == "c":
            u = np.ascontiguousarray(uniforms, dtype=np.float64)
            idx = np.empty(size, dtype=np.int64)
            w = np.empty(size, dtype=np.float64)
            clib().oper_sample(_dp(self.sum_tree.tree), _dp(self.min_tree.tree), self.power_of_2_size, size, _dp(u),
                               nt, beta, _ip(idx), _dp(w), None)
        else:
            idx = np.empty(size, dtype=np.int64)
            w = np.empty(size, dtype=np.float64)
            total = self.sum_tree.total_value()
            segment_size = total / size
            min_probability = self.min_tree.total_value() / total
            max_weight = (min_probability * nt) ** -beta
            for i in range(size):
                a = segment_size * i
                b = segment_size * (i + 1)
                val = a + (b - a) * uniforms[i]
                leaf, priority = self.sum_tree.retrieve(val)
                priority = priority / total
                weight = (nt * priority) ** -beta
                idx[i] = leaf
                w[i] = weight / max_weight
        self.beta.step()
        return idx, w

    def sample(self, size, uniforms=None):
        idx, w = self.sample_indices(size, uniforms)
        return [self.data[i] for i in idx], idx, w


# ----------------------------------------------------------------------------------------------------------------
def batch_columns(transitions, obs_key="observation"):
    """core_types.py:488-585 -- AoS -> SoA exactly as Batch does it (np.array of per-transition np.array).
    ``transitions`` are objects with .state/.next_state dicts, .action, .reward, .game_over."""
    states = np.array([np.array(t.state[obs_key]) for t in transitions])
    next_states = np.array([np.array(t.next_state[obs_key]) for t in transitions])
    actions = np.array([t.action for t in transitions])
    rewards = np.array([t.reward for t in transitions])
    game_overs = np.array([t.game_over for t in transitions])
    return states, next_states, actions, rewards, game_overs
