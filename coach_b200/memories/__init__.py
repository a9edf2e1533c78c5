from coach_b200.memories.memory import Memory, MemoryGranularity, MemoryParameters  # noqa: F401
