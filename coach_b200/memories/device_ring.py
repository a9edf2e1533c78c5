"""HBM-resident struct-of-arrays transition ring.

The reference keeps a Python ``list`` of ``Transition`` objects (memories/non_episodic/experience_replay.py:53,146)
and converts AoS -> SoA on every sample (core_types.py:488-623).  Here every transition field is one row-major
``[capacity, row_bytes]`` uint8 matrix in HBM ("column"); a sample is a row gather
(``cb200_gather`` / ``cb200_per_sample_gather``), an append is a row scatter (``cb200_scatter_ring``).

Layout for the Atari configuration (2^20 slots): state 28,224 B + next_state 28,224 B + action 8 B + reward 8 B +
game_over 1 B per slot = 59.2 GB, sized for the 180 GB of a B200.

Columns
  ``state:<key>`` / ``next_state:<key>``  one per entry of the transition's state dict, dtype/shape as stored
  ``action``     int64 scalar (discrete) or float vector (continuous), as given
  ``reward``     float64 (the reference's rewards are Python floats, core_types.py:523)
  ``game_over``  uint8
"""
from collections import OrderedDict

import numpy as np
import torch

from coach_b200 import _lib


def _ptr_of(p):
    return p.data_ptr() if torch.is_tensor(p) else p.ptr


class ColumnSpec(object):
    def __init__(self, name, shape, dtype):
        self.name = name
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.row_bytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize if self.shape else \
            self.dtype.itemsize

    def torch_dtype(self):
        return {np.dtype(np.uint8): torch.uint8, np.dtype(np.int8): torch.int8, np.dtype(np.int32): torch.int32,
                np.dtype(np.int64): torch.int64, np.dtype(np.float32): torch.float32,
                np.dtype(np.float64): torch.float64, np.dtype(np.bool_): torch.uint8,
                np.dtype(np.float16): torch.float16, np.dtype(np.int16): torch.int16}[self.dtype]

    def __repr__(self):
        return "ColumnSpec(%s, %s, %s)" % (self.name, self.shape, self.dtype)


def schema_from_transition(t):
    """Column layout inferred from the first stored transition."""
    specs = OrderedDict()
    for prefix, d in (("state:", t.state), ("next_state:", t.next_state)):
        for key in sorted(d.keys()):
            a = np.asarray(d[key])
            specs[prefix + key] = ColumnSpec(prefix + key, a.shape, a.dtype)
    a = np.asarray(t.action)
    if a.dtype.kind in "iub":
        a = a.astype(np.int64)
    specs["action"] = ColumnSpec("action", a.shape, a.dtype)
    specs["reward"] = ColumnSpec("reward", (), np.float64)
    specs["game_over"] = ColumnSpec("game_over", (), np.uint8)
    return specs


class DeviceTransitionRing(object):
    """Owns the HBM columns plus a pinned host staging area so that single ``store(transition)`` calls are a memcpy
    into pinned memory and reach the GPU in batches (one H2D copy + one scatter kernel per column per flush)."""

    def __init__(self, capacity, device=None, stage_rows=256):
        self.capacity = int(capacity)
        self.device = torch.device(device if device is not None else "cuda")
        self.lib = _lib.load()
        self.specs = None
        self.columns = None          # name -> uint8 [capacity, row_bytes] on device
        self.stage_rows = int(stage_rows)
        self._table_cache = {}
        self._stage_host = None      # name -> pinned uint8 [stage_rows, row_bytes]
        self._stage_dev = None
        self._pending = 0
        self._flush_event = None     # H2D of the pinned stage still in flight?
        self.cursor = 0              # next slot to write
        self.count = 0               # valid slots (<= capacity)

    # -- schema ----------------------------------------------------------------------------------------------------
    def set_schema(self, specs):
        if self.specs is not None:
            return
        self.specs = specs
        self.columns, self._stage_host, self._stage_dev = OrderedDict(), OrderedDict(), OrderedDict()
        self._stage_np = {}
        pin = self.device.type == "cuda"
        # staging area: ONE pinned record per staged transition (columns at 16-byte aligned offsets inside it), so a
        # flush is one H2D copy of the used prefix plus one scatter launch for all columns
        self._rec_off, off = {}, 0
        for name, sp in specs.items():
            self._rec_off[name] = off
            off = (off + sp.row_bytes + 15) // 16 * 16
        self._rec_bytes = off
        self._stage_host_all = torch.zeros((self.stage_rows, self._rec_bytes), dtype=torch.uint8, pin_memory=pin)
        self._stage_dev_all = torch.empty((self.stage_rows, self._rec_bytes), dtype=torch.uint8, device=self.device)
        host_np = self._stage_host_all.numpy()
        for name, sp in specs.items():
            self.columns[name] = torch.empty((self.capacity, sp.row_bytes), dtype=torch.uint8, device=self.device)
            o = self._rec_off[name]
            self._stage_np[name] = host_np[:, o:o + sp.row_bytes]      # same memory, no per-store tensor objects
        self._flush_table = None

    def declare_schema(self, columns):
        """Fix the column layout before the first store: {name: (shape, numpy dtype)} or {name: batch tensor [n, ...]}.
        ``store(Transition)`` then converts every field to the declared dtype (gym hands out float64 observations and
        actions where the networks -- and the agents' persistent batch buffers -- are float32)."""
        if self.specs is not None:
            raise RuntimeError("the replay already holds transitions; the schema is fixed")
        specs = OrderedDict()
        for name, v in columns.items():
            if torch.is_tensor(v):
                dt = np.dtype(str(v.dtype).replace("torch.", "")) if v.dtype != torch.bool else np.dtype(np.uint8)
                specs[name] = ColumnSpec(name, tuple(v.shape[1:]), dt)
            else:
                specs[name] = ColumnSpec(name, v[0], v[1])
        for need in ("action", "reward", "game_over"):
            if need not in specs:
                raise ValueError("schema lacks the %r column" % need)
        self.set_schema(specs)

    def hbm_bytes(self):
        return 0 if self.columns is None else sum(c.numel() for c in self.columns.values())

    # -- append ----------------------------------------------------------------------------------------------------
    def stage_transition(self, t):
        """Copies one transition into the pinned staging rows; returns True when the stage is full."""
        if self.specs is None:
            self.set_schema(schema_from_transition(t))
        r = self._pending
        if r == 0 and self._flush_event is not None:
            self._flush_event.synchronize()      # the previous flush's H2D must have drained the pinned rows
            self._flush_event = None
        for name, sp in self.specs.items():
            if name.startswith("state:"):
                v = t.state[name[6:]]
            elif name.startswith("next_state:"):
                v = t.next_state[name[11:]]
            elif name == "action":
                v = t.action
            elif name == "reward":
                v = t.reward
            else:
                v = t.game_over
            a = np.asarray(v, dtype=sp.dtype, order='C')
            if a.shape != sp.shape:
                raise ValueError("transition field %s has shape %s, the replay was created with %s"
                                 % (name, a.shape, sp.shape))
            self._stage_np[name][r] = a.reshape(-1).view(np.uint8)
        self._pending += 1
        return self._pending >= min(self.stage_rows, self.capacity)

    def flush(self):
        """Moves the staged rows into the ring.  Returns (first_slot, n)."""
        n = self._pending
        if n == 0:
            return self.cursor, 0
        first = self.cursor
        self._stage_dev_all[:n].copy_(self._stage_host_all[:n], non_blocking=True)
        if self._flush_table is None:
            base = self._stage_dev_all.data_ptr()
            pairs = [(self.columns[name].data_ptr(), base + self._rec_off[name], sp.row_bytes)
                     for name, sp in self.specs.items()]
            self._flush_table = [_lib.make_columns(pairs[k:k + _lib.CB200_MAX_COLUMNS])
                                 for k in range(0, len(pairs), _lib.CB200_MAX_COLUMNS)]
        for arr, cnt in self._flush_table:
            _lib.check(self.lib.cb200_scatter_ring_packed(arr, cnt, self._rec_bytes, self.cursor, self.capacity, n,
                                                          _lib.current_stream()))
        self.cursor = (self.cursor + n) % self.capacity
        self.count = min(self.count + n, self.capacity)
        self._pending = 0
        if self.device.type == "cuda":
            self._flush_event = torch.cuda.Event()
            self._flush_event.record()
        return first, n

    def append_columns(self, cols):
        """Bulk append of n transitions given as {column name: tensor/ndarray [n, ...]} (device tensors are used in
        place, host arrays go through one H2D copy).  Returns (first_slot, n)."""
        self.flush()
        n = None
        staged = {}
        for name, v in cols.items():
            tt = torch.as_tensor(v)
            n = tt.shape[0] if n is None else n
            if tt.shape[0] != n:
                raise ValueError("all columns must have the same number of rows")
            staged[name] = tt
        if self.specs is None:
            specs = OrderedDict()
            for name, tt in staged.items():
                dt = np.dtype(str(tt.dtype).replace("torch.", "")) if tt.dtype != torch.bool else np.dtype(np.uint8)
                specs[name] = ColumnSpec(name, tuple(tt.shape[1:]), dt)
            self.set_schema(specs)
        if set(staged) != set(self.specs):
            raise ValueError("append_columns needs exactly the columns %s" % list(self.specs))
        if n > self.capacity:
            raise ValueError("cannot append more rows than the ring holds in one call")
        first = self.cursor
        pairs, keep = [], []
        for name, sp in self.specs.items():
            tt = staged[name].to(self.device, non_blocking=True).contiguous()
            tt = tt.view(torch.uint8).reshape(n, -1) if tt.dtype != torch.bool else tt.to(torch.uint8).reshape(n, -1)
            if tt.shape[1] != sp.row_bytes:
                raise ValueError("column %s: %d bytes per row, expected %d" % (name, tt.shape[1], sp.row_bytes))
            keep.append(tt)
            pairs.append((self.columns[name].data_ptr(), tt.data_ptr(), sp.row_bytes))
        self._scatter(pairs, n)
        return first, n

    def _scatter(self, pairs, n):
        for k in range(0, len(pairs), _lib.CB200_MAX_COLUMNS):
            arr, cnt = _lib.make_columns(pairs[k:k + _lib.CB200_MAX_COLUMNS])
            _lib.check(self.lib.cb200_scatter_ring(arr, cnt, self.cursor, self.capacity, n, _lib.current_stream()))
        self.cursor = (self.cursor + n) % self.capacity
        self.count = min(self.count + n, self.capacity)

    def clear(self):
        self.cursor = 0
        self.count = 0
        self._pending = 0

    # -- gather ----------------------------------------------------------------------------------------------------
    def alloc_batch(self, n):
        """Output tensors for one minibatch: {column: typed tensor [n, *shape]}."""
        out = OrderedDict()
        for name, sp in self.specs.items():
            out[name] = torch.empty((n,) + sp.shape, dtype=sp.torch_dtype(), device=self.device)
        return out

    def column_table(self, out, n):
        """ctypes gather table for ``n`` rows into ``out``.  The kernels write ``n * row_bytes`` bytes per column: a
        destination of another dtype or size would be overrun (or read back reinterpreted), so every buffer is checked
        against the ring's schema -- once per buffer set, the agents sample into the same persistent buffers."""
        key = (int(n),) + tuple(out[name].data_ptr() for name in self.specs)
        hit = self._table_cache.get(key)
        if hit is None:
            for name, sp in self.specs.items():
                t = out[name]
                if not (t.is_cuda == (self.device.type == "cuda") and t.is_contiguous()):
                    raise ValueError("batch buffer %r must be a contiguous tensor on %s" % (name, self.device))
                if t.dtype != sp.torch_dtype() or t.numel() * t.element_size() != n * sp.row_bytes:
                    raise ValueError(
                        "batch buffer %r is %s%s but the replay stores %s%s per transition (%d rows): declare the "
                        "schema up front (memory.declare_schema) or store transitions in the agent's dtypes"
                        % (name, t.dtype, tuple(t.shape), sp.dtype, sp.shape, n))
            pairs = [(self.columns[name].data_ptr(), out[name].data_ptr(), sp.row_bytes)
                     for name, sp in self.specs.items()]
            hit = self._table_cache[key] = _lib.make_columns(pairs)
            if len(self._table_cache) > 64:
                self._table_cache.clear()
                self._table_cache[key] = hit
        return hit

    def s2d_tables(self, s2d, out, n):
        """ctypes tables of the fused gather + space-to-depth launch (cb200_per_sample_gather_s2d / cb200_gather_s2d).
        s2d: {"columns": {ring column name: plane tensor / PlaneBuf}, "geometry": (H, W, C, S)}; every other column
        is copied into ``out``.  Returns (image table, n_image, small table, n_small, image names)."""
        key = ("s2d", int(n)) + tuple(_ptr_of(p) for p in s2d["columns"].values()) + \
            tuple(out[name].data_ptr() for name in self.specs if name not in s2d["columns"])
        hit = self._table_cache.get(key)
        if hit is None:
            H, W, C, S = s2d["geometry"]
            img, small = [], []
            for name, sp in self.specs.items():
                if name in s2d["columns"]:
                    if sp.dtype != np.uint8 or sp.row_bytes != H * W * C:
                        raise ValueError("column %r is %s%s, the fused image path needs uint8 [%d, %d, %d] frames"
                                         % (name, sp.dtype, sp.shape, H, W, C))
                    img.append((self.columns[name].data_ptr(), _ptr_of(s2d["columns"][name]), sp.row_bytes))
                else:
                    t = out[name]
                    if t.dtype != sp.torch_dtype() or t.numel() * t.element_size() != n * sp.row_bytes or \
                            not t.is_contiguous():
                        raise ValueError("batch buffer %r does not match the replay's %s%s" % (name, sp.dtype, sp.shape))
                    small.append((self.columns[name].data_ptr(), t.data_ptr(), sp.row_bytes))
            ia, ni = _lib.make_columns(img)
            sa, ns = _lib.make_columns(small)
            hit = self._table_cache[key] = (ia, ni, sa, ns)
        return hit

    def gather_column(self, name, idx):
        """one column of the given slots as a typed tensor (lazy materialisation of un-staged batch columns)"""
        sp = self.specs[name]
        n = idx.shape[0]
        out = torch.empty((n,) + sp.shape, dtype=sp.torch_dtype(), device=self.device)
        arr, cnt = _lib.make_columns([(self.columns[name].data_ptr(), out.data_ptr(), sp.row_bytes)])
        _lib.check(self.lib.cb200_gather(arr, cnt, idx.data_ptr(), n, _lib.current_stream()))
        return out

    def gather(self, idx, out=None):
        """out[c][i] = column c of slot idx[i]; idx int64 CUDA tensor."""
        n = idx.shape[0]
        if out is None:
            out = self.alloc_batch(n)
        arr, cnt = self.column_table(out, n)
        _lib.check(self.lib.cb200_gather(arr, cnt, idx.data_ptr(), n, _lib.current_stream()))
        return out
