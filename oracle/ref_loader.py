"""Import the unmodified reference (rl_coach) in the build container.

``rl_coach/__init__.py:6`` imports tensorflow unconditionally and ``memories/backend/redis.py`` imports redis; neither
is installed, so both are replaced by ``MagicMock`` modules *before* the first import.  Only numpy-level modules
(memories, filters, core_types, schedules, agents' numpy prologues) are usable this way.

TEST INFRASTRUCTURE ONLY -- never imported by ``coach_b200``; unavailable on the GPU box (no /root/reference there).
"""
import os
import sys
from unittest import mock

REFERENCE_ROOT = os.environ.get("COACH_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "rl_coach"))


def load():
    """Returns the imported ``rl_coach`` package (raises RuntimeError when the reference tree is absent)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for name in ("tensorflow", "tensorflow.contrib", "tensorflow.python", "redis", "pygame", "pygame.locals", "skimage",
                 "skimage.transform"):
        if name not in sys.modules:
            sys.modules[name] = mock.MagicMock()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import rl_coach  # noqa: F401
    return rl_coach
