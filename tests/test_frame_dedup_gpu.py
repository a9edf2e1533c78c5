"""Frame-deduplicated ring: every frame stored once, stacks assembled by the gather kernels.

Pinned to the reference: tests/golden/agent_prologues.npz fs_* = an episode stream through the UNMODIFIED
rl_coach ObservationStackingFilter / LazyStack (observation_stacking_filter.py:27-115; written by
oracle/make_golden_agents.py golden_frame_stream).  The ring must hand back byte for byte the stacked state / next_state
every transition of that stream carried."""
import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fx():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "agent_prologues.npz"))


def _stream_transitions(fx, lazy=True):
    """the fixture's raw frames through coach_b200's own stacking filter -> Transitions carrying LazyStacks (or arrays)"""
    from coach_b200.core_types import Transition
    from coach_b200.filters.filter import ObservationStackingFilter
    frames = fx["fs_frames"]
    flt = ObservationStackingFilter(4)
    ts, f, k = [], 0, 0
    for L in fx["fs_episode_lengths"]:
        flt.reset()
        s = flt.filter(frames[f])
        f += 1
        for t in range(int(L)):
            s2 = flt.filter(frames[f])
            f += 1
            st, nx = (s, s2) if lazy else (np.array(s), np.array(s2))
            ts.append(Transition(state={"observation": st}, action=int(fx["fs_actions"][k]),
                                 reward=float(fx["fs_rewards"][k]), next_state={"observation": nx},
                                 game_over=bool(fx["fs_game_overs"][k])))
            k += 1
            s = s2
    return ts


def _memory(per, capacity, H=16, W=16):
    from coach_b200.memories.experience_replay import ExperienceReplay
    from coach_b200.memories.memory import MemoryGranularity
    from coach_b200.memories.prioritized_experience_replay import PrioritizedExperienceReplay
    cls = PrioritizedExperienceReplay if per else ExperienceReplay
    mem = cls((MemoryGranularity.Transitions, capacity), frame_dedup=True)
    mem.declare_schema({"state:observation": ((H, W, 4), np.uint8), "next_state:observation": ((H, W, 4), np.uint8),
                        "action": ((), np.int64), "reward": ((), np.float64), "game_over": ((), np.uint8)},
                       image_columns=("state:observation", "next_state:observation"))
    return mem


@pytest.mark.parametrize("lazy", [True, False], ids=["lazystack_identity", "arrays_by_content"])
@pytest.mark.parametrize("per", [True, False])
def test_gathered_stacks_equal_the_reference_filter_stream(fx, per, lazy):
    ts = _stream_transitions(fx, lazy)
    n = len(ts)
    mem = _memory(per, 64)
    # our filter's LazyStacks reproduce the reference's stacks (host side)
    np.testing.assert_array_equal(np.array(ts[7].state["observation"]), fx["fs_states"][7])
    for t in ts:
        mem.store(t)
    mem._flush()
    ring = mem.ring
    n_frames = n + len(fx["fs_episode_lengths"])
    assert ring._fc == n_frames == fx["fs_frames"].shape[0]           # every frame stored exactly once
    idx = torch.arange(n, device="cuda")
    got = ring.gather(idx)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(got["state:observation"].cpu().numpy(), fx["fs_states"])
    np.testing.assert_array_equal(got["next_state:observation"].cpu().numpy(), fx["fs_next_states"])
    np.testing.assert_array_equal(got["action"].cpu().numpy(), fx["fs_actions"])
    np.testing.assert_array_equal(got["reward"].cpu().numpy(), fx["fs_rewards"])
    np.testing.assert_array_equal(got["game_over"].cpu().numpy(), fx["fs_game_overs"])
    # host-materialised Transitions (the API-compatible slow path) carry the same stacks
    random.seed(0), np.random.seed(0)
    b = mem.sample_batch(8)
    k = b.info("idx").cpu().numpy() if per else b.columns["idx"].cpu().numpy()
    np.testing.assert_array_equal(b.columns["state:observation"].cpu().numpy(), fx["fs_states"][k])


@pytest.mark.parametrize("tma", [0, 1], ids=["bulk_copies", "tma_boxes"])
@pytest.mark.parametrize("per", [True, False])
def test_fused_s2d_planes_from_the_frame_store(fx, per, tma):
    """cb200_per_sample_gather_s2d / cb200_gather_s2d on the frame store == cb200_u8_s2d_planes of the reference stacks
    (both copy paths: one bulk copy per frame, or one 2-D TMA box per stack in consecutive slots)"""
    from coach_b200 import _lib as L
    from coach_b200.architectures import tiled as tl
    lib = L.load()
    lib.cb200_tune(b"frame_tma", tma)
    dev = torch.device("cuda")
    ts = _stream_transitions(fx, True)
    mem = _memory(per, 64)
    for t in ts:
        mem.store(t)
    B, H, W, C, S = 32, 16, 16, 4, 4
    npix, Cs = (H // S) * (W // S), S * S * C
    planes = {k: tl.PlaneBuf(npix * B, Cs, dev, npix=npix, nplanes=1) for k in ("state:observation",
                                                                              "next_state:observation")}
    out = {"action": torch.zeros(B, dtype=torch.int64, device=dev), "reward": torch.zeros(B, dtype=torch.float64, device=dev),
           "game_over": torch.zeros(B, dtype=torch.uint8, device=dev)}
    for trial in range(3):
        random.seed(trial), np.random.seed(trial)
        got = mem.sample_batch(B, out=dict(out), s2d={"columns": planes, "geometry": (H, W, C, S)})
        torch.cuda.synchronize()
        k = got.columns["idx"].cpu().numpy()
        np.testing.assert_array_equal(got.columns["action"].cpu().numpy(), fx["fs_actions"][k])
        for name, src in (("state:observation", fx["fs_states"]), ("next_state:observation", fx["fs_next_states"])):
            x = torch.from_numpy(np.ascontiguousarray(src[k])).cuda()
            want = tl.PlaneBuf(npix * B, Cs, dev, npix=npix, nplanes=1)
            L.check(lib.cb200_u8_s2d_planes(x.data_ptr(), B, H, W, C, S, want.ptr, None))
            torch.cuda.synchronize()
            assert torch.equal(planes[name].t, want.t), name
            assert torch.equal(got.column(name), x)                    # lazily materialised uint8 column
    lib.cb200_tune(b"frame_tma", 0)


def test_ring_and_frame_store_wrap_around(fx):
    """capacity 16 < 45 transitions, 28 frame slots < 52 frames: both rings wrap; the live slots hold the last 16
    transitions of the stream, frames of evicted transitions are overwritten without touching live ones"""
    ts = _stream_transitions(fx, True)
    mem = _memory(False, 16)
    assert mem.ring.specs is not None
    for i, t in enumerate(ts):
        mem.store(t)
        if i % 5 == 4:
            mem._flush()
    mem._flush()
    ring = mem.ring
    assert ring.frame_capacity < fx["fs_frames"].shape[0] and ring.count == 16
    n = len(ts)
    live = np.arange(n - 16, n)
    slots = torch.from_numpy((live % 16).astype(np.int64)).cuda()
    got = ring.gather(slots)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(got["state:observation"].cpu().numpy(), fx["fs_states"][live])
    np.testing.assert_array_equal(got["next_state:observation"].cpu().numpy(), fx["fs_next_states"][live])


def test_frame_store_exhaustion_is_reported():
    """one-step episodes need two frames per transition: a frame store without slack runs out and says so"""
    from coach_b200.core_types import Transition
    from coach_b200.memories.experience_replay import ExperienceReplay
    from coach_b200.memories.memory import MemoryGranularity
    mem = ExperienceReplay((MemoryGranularity.Transitions, 64), frame_dedup=True, frame_slack=0.0)
    mem.declare_schema({"state:observation": ((16, 16, 4), np.uint8), "next_state:observation": ((16, 16, 4), np.uint8),
                        "action": ((), np.int64), "reward": ((), np.float64), "game_over": ((), np.uint8)},
                       image_columns=("state:observation", "next_state:observation"))
    rng = np.random.RandomState(0)
    with pytest.raises(RuntimeError, match="frame store exhausted"):
        for i in range(200):
            s = rng.randint(0, 256, (16, 16, 4)).astype(np.uint8)
            s2 = rng.randint(0, 256, (16, 16, 4)).astype(np.uint8)
            mem.store(Transition(state={"observation": s}, action=0, reward=0.0, next_state={"observation": s2},
                                 game_over=True))


def test_dqn_learn_step_is_identical_on_the_frame_deduplicated_replay():
    """DQN + PER, Atari shapes, batch 128 (fused input path): the same transitions through a verbatim ring and through
    the frame store give bit-identical losses, gradients norms and parameters"""
    from coach_b200.agents.dqn_agent import DQNAgent, DQNAgentParameters
    from coach_b200.memories.memory import MemoryGranularity
    from coach_b200.memories.prioritized_experience_replay import PrioritizedExperienceReplayParameters
    res = []
    for dedup in (False, True):
        ap = DQNAgentParameters()
        ap.memory = PrioritizedExperienceReplayParameters()
        ap.memory.max_size = (MemoryGranularity.Transitions, 4096)
        ap.memory.frame_dedup = dedup
        ap.network_wrappers["main"].batch_size = 128
        agent = DQNAgent(ap, observation_shape=(84, 84, 4), num_actions=6, seed=0)
        assert bool(agent.memory.ring.stack_cols) == dedup and agent.s2d is not None
        rng = np.random.RandomState(1)
        n = 256
        agent.memory.store_columns({
            "state:observation": rng.randint(0, 256, (n, 84, 84, 4)).astype(np.uint8),
            "next_state:observation": rng.randint(0, 256, (n, 84, 84, 4)).astype(np.uint8),
            "action": rng.randint(0, 6, n).astype(np.int64), "reward": rng.randint(-1, 2, n).astype(np.float64),
            "game_over": (rng.rand(n) < 0.1).astype(np.uint8)})
        agent.memory.update_priorities(np.arange(n), np.abs(rng.randn(n)))
        losses = []
        for step in range(4):
            random.seed(step), np.random.seed(step)
            batch = agent.sample_batch()
            loss, _, gnorm = agent.learn_from_batch(batch)
            losses.append((loss, gnorm))
        torch.cuda.synchronize()
        res.append((losses, agent.net_def.store.theta.clone()))
    assert res[0][0] == res[1][0]
    assert torch.equal(res[0][1], res[1][1])
    if True:
        ring = agent.memory.ring
        assert ring.hbm_bytes() < 0.4 * 4096 * 2 * 84 * 84 * 4
