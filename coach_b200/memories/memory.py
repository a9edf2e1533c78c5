"""Memory base types, mirroring ``rl_coach/memories/memory.py:24-77``."""
from enum import Enum
from typing import Tuple


class MemoryGranularity(Enum):
    Transitions = 0
    Episodes = 1


class MemoryParameters(object):
    """Plain attribute bag with the reference's field names (memories/memory.py:29-38).  ``path`` is the
    ``'module:Class'`` string resolved by Coach's ``short_dynamic_import`` (utils.py:334-356)."""

    def __init__(self):
        self.max_size = None
        self.shared_memory = False
        self.load_memory_from_file_path = None

    @property
    def path(self):
        return 'coach_b200.memories.memory:Memory'


class Memory(object):
    def __init__(self, max_size: Tuple[MemoryGranularity, int]):
        self.max_size = max_size
        self._length = 0
        self.memory_backend = None

    def store(self, obj):
        if self.memory_backend:
            self.memory_backend.store(obj)

    def store_episode(self, episode):
        if self.memory_backend:
            self.memory_backend.store(episode)

    def get(self, index):
        raise NotImplementedError("")

    def length(self):
        raise NotImplementedError("")

    def sample(self, size):
        raise NotImplementedError("")

    def clean(self):
        raise NotImplementedError("")

    def set_memory_backend(self, memory_backend):
        self.memory_backend = memory_backend

    def num_transitions(self) -> int:
        raise NotImplementedError("")
