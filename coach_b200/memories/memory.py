"""The contract every replay memory of this package fulfils.

Coach addresses a memory through three things (``rl_coach/memories/memory.py:24-77``; call sites in
``agents/agent.py:90,449-465``): a ``MemoryParameters`` object whose ``path`` names the class to instantiate, the
``(MemoryGranularity, count)`` capacity tuple, and a handful of methods invoked by name through
``Agent.call_memory``.  The device memories (ring + segment trees in HBM) derive from ``Memory`` below, which fixes
that method set and the optional write-through to a distributed ``memory_backend``.
"""
import enum
from typing import Tuple


class MemoryGranularity(enum.Enum):
    """unit of the capacity in ``max_size``: single transitions or whole episodes"""
    Transitions = 0
    Episodes = 1


class MemoryParameters(object):
    """Attribute bag consumed by ``dynamic_import_and_instantiate_module_from_params`` (utils.py:401-404): every
    attribute whose name matches a constructor argument of the class behind ``path`` is passed to it."""

    path = property(lambda self: "coach_b200.memories.memory:Memory")

    def __init__(self):
        self.load_memory_from_file_path = None      # pre-recorded data set (batch RL), not used on the device path
        self.shared_memory = False                  # reference: Manager-proxied memory shared by workers
        self.max_size = None                        # (MemoryGranularity, count)


def _abstract(name):
    def method(self, *args, **kwargs):
        raise NotImplementedError("%s.%s" % (type(self).__name__, name))
    method.__name__ = name
    return method


class Memory(object):
    """Base of all memories.  Sub-classes provide ``get / length / sample / clean / num_transitions``; ``store`` and
    ``store_episode`` here only mirror the object to the distributed backend when one is attached (the reference's
    Redis pub/sub transport), which sub-classes call before writing their own storage."""

    def __init__(self, max_size: Tuple[MemoryGranularity, int]):
        self.memory_backend = None
        self.max_size = max_size
        self._length = 0

    def set_memory_backend(self, memory_backend):
        self.memory_backend = memory_backend

    def _mirror(self, obj):
        if self.memory_backend:
            self.memory_backend.store(obj)

    def store(self, obj):
        self._mirror(obj)

    def store_episode(self, episode):
        self._mirror(episode)

    get = _abstract("get")
    length = _abstract("length")
    sample = _abstract("sample")
    clean = _abstract("clean")
    num_transitions = _abstract("num_transitions")
