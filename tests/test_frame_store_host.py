"""Host side of the frame-deduplicated ring (coach_b200/memories/device_ring.py: _frame_slots): which frames of the
stacked observations a transition stream carries are new, which are shared.  No GPU: the staging arrays are plain numpy.
The stream is the reference fixture of tests/golden/agent_prologues.npz (fs_*: unmodified ObservationStackingFilter)."""
import os
from collections import OrderedDict

import numpy as np
import pytest

from coach_b200.filters.filter import ObservationStackingFilter
from coach_b200.memories.device_ring import DeviceTransitionRing


def _bare_ring(capacity, frame_capacity, frame_bytes, stage_rows=512):
    r = DeviceTransitionRing.__new__(DeviceTransitionRing)
    r.capacity, r.count, r._pending, r.cursor = capacity, 0, 0, 0
    r.frame_capacity, r._fc, r._pending_frames, r.frame_bytes, r.frame_slack = frame_capacity, 0, 0, frame_bytes, 0.25
    r.frame_streams = 1
    r._recent = OrderedDict()
    r._min_fc = np.zeros(capacity, dtype=np.int64)
    r._frame_stage_np = np.zeros((stage_rows, frame_bytes), dtype=np.uint8)
    return r


@pytest.fixture(scope="module")
def fx():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "agent_prologues.npz"))


def _stream(fx, lazy):
    frames = fx["fs_frames"]
    flt = ObservationStackingFilter(4)
    out, f = [], 0
    for L in fx["fs_episode_lengths"]:
        flt.reset()
        s = flt.filter(frames[f])
        f += 1
        for _ in range(int(L)):
            s2 = flt.filter(frames[f])
            f += 1
            out.append((s, s2) if lazy else (np.array(s), np.array(s2)))
            s = s2
    return out


@pytest.mark.parametrize("lazy", [True, False], ids=["lazystack_identity", "arrays_by_content"])
def test_every_frame_of_the_reference_stream_is_staged_once(fx, lazy):
    ring = _bare_ring(1024, 2048, 16 * 16)
    rows = []
    for k, (s, s2) in enumerate(_stream(fx, lazy)):
        a, b = np.zeros(4, np.int32), np.zeros(4, np.int32)
        o1 = ring._frame_slots(s, 4, a)
        o2 = ring._frame_slots(s2, 4, b)
        ring._min_fc[k] = min(o1, o2)
        ring._pending += 1
        rows.append((a, b))
    n_frames = fx["fs_frames"].shape[0]
    assert ring._fc == n_frames == ring._pending_frames
    # the staged frames are the stream's frames in arrival order ...
    np.testing.assert_array_equal(ring._frame_stage_np[:n_frames].reshape(n_frames, 16, 16), fx["fs_frames"])
    # ... and the slot tables rebuild exactly the stacks the reference filter handed out
    store = ring._frame_stage_np[:n_frames].reshape(n_frames, 16, 16)
    for k, (a, b) in enumerate(rows):
        np.testing.assert_array_equal(np.stack([store[i] for i in a], axis=-1), fx["fs_states"][k])
        np.testing.assert_array_equal(np.stack([store[i] for i in b], axis=-1), fx["fs_next_states"][k])
    # episode start: the first frame is replicated (one slot four times), then the window slides by one new frame
    assert len(set(rows[0][0])) == 1 and rows[0][1][3] == rows[0][0][0] + 1
    assert (np.diff(ring._min_fc[:len(rows)]) >= 0).all()          # older transitions reference older frames


def test_frame_store_guard_and_argument_checks():
    ring = _bare_ring(4, 6, 16)                      # 4 transitions, 6 frame slots
    flt = ObservationStackingFilter(4)
    rng = np.random.RandomState(0)
    s = flt.filter(rng.randint(0, 256, (4, 4)).astype(np.uint8))
    row = np.zeros(4, np.int32)
    with pytest.raises(RuntimeError, match="frame store exhausted"):
        for k in range(32):                          # every transition starts a new episode: two frames each
            flt.reset()
            s = flt.filter(rng.randint(0, 256, (4, 4)).astype(np.uint8))
            s2 = flt.filter(rng.randint(0, 256, (4, 4)).astype(np.uint8))
            o = min(ring._frame_slots(s, 4, row), ring._frame_slots(s2, 4, row))
            ring._min_fc[(ring.cursor + ring._pending) % 4] = o
            ring._pending += 1
            if ring._pending == 2:                   # what flush() does to the counters
                ring.cursor, ring.count = (ring.cursor + 2) % 4, min(ring.count + 2, 4)
                ring._pending = ring._pending_frames = 0
    with pytest.raises(ValueError, match="frames on the last axis"):
        ring2 = _bare_ring(4, 64, 16)
        ring2._frame_slots(np.zeros((4, 4, 3), np.uint8), 4, row)
    with pytest.raises(ValueError, match="-byte frames"):
        ring3 = _bare_ring(4, 64, 16)
        ring3._frame_slots(np.zeros((5, 5, 4), np.uint8), 4, row)


def test_interleaved_environment_streams_share_frames_within_each_stream():
    """E environments store() in turn: with the identity cache sized for E streams every frame is still staged once"""
    E, T = 6, 20
    rng = np.random.RandomState(1)
    for streams, exact in ((E, True), (1, False)):
        ring = _bare_ring(4096, 8192, 16, stage_rows=2048)
        ring.frame_streams = streams
        flts = [ObservationStackingFilter(4) for _ in range(E)]
        cur = [f.filter(rng.randint(0, 256, (4, 4)).astype(np.uint8)) for f in flts]
        row = np.zeros(4, np.int32)
        stacks = []
        for t in range(T):
            for e in range(E):
                nxt = flts[e].filter(rng.randint(0, 256, (4, 4)).astype(np.uint8))
                a, b = np.zeros(4, np.int32), np.zeros(4, np.int32)
                ring._frame_slots(cur[e], 4, a)
                ring._frame_slots(nxt, 4, b)
                stacks.append((np.array(cur[e]), a, np.array(nxt), b))
                ring._pending += 1
                cur[e] = nxt
        store = ring._frame_stage_np[:ring._fc].reshape(-1, 4, 4)
        for s, a, s2, b in stacks:               # correct either way ...
            np.testing.assert_array_equal(np.stack([store[i] for i in a], axis=-1), s)
            np.testing.assert_array_equal(np.stack([store[i] for i in b], axis=-1), s2)
        if exact:                                # ... and without a single duplicate when the cache covers the streams
            assert ring._fc == E * (T + 1)
        else:
            assert ring._fc > E * (T + 1)
