"""CPU-only tests of the host-side mirrors of the reference interfaces (no CUDA calls)."""
import os
import random

import numpy as np
import pytest


def test_linear_schedule_matches_reference_fixture(golden_dir):
    from coach_b200.schedules import LinearSchedule
    fx = np.load(os.path.join(golden_dir, "linear_schedule.npz"))
    s = LinearSchedule(0.4, 1.0, 1000)
    for v in fx["vals"]:
        assert float(s.current_value) == v
        s.step()


def test_transition_and_batch_mirror_reference_semantics():
    import copy
    from coach_b200.core_types import Batch, Transition
    t = Transition(state={'observation': np.arange(3)}, action=1, reward=0.5, game_over=False)
    assert t.next_state is t.state                      # missing next_state defaults to state (core_types.py:221-222)
    with pytest.raises(Exception, match="n_step_discounted_rewards"):
        t.n_step_discounted_rewards
    with pytest.raises(Exception, match="The state was not filled"):
        Transition().state
    t.add_info({'idx': 3})
    with pytest.raises(ValueError):
        t.add_info({'idx': 4})
    c = copy.copy(t)
    c.info['weight'] = 1.0
    assert 'weight' not in t.info and c.state is not t.state
    ts = [Transition(state={'observation': np.full(3, i)}, action=i % 2, reward=float(i),
                     next_state={'observation': np.full(3, i + 1)}, game_over=(i == 4), info={'idx': i})
          for i in range(5)]
    b = Batch(ts)
    assert b.states(['observation'])['observation'].shape == (5, 3)
    assert b.rewards().dtype == np.float64 and b.game_overs().dtype == np.bool_
    assert b.actions(expand_dims=True).shape == (5, 1)
    np.testing.assert_array_equal(b.info('idx'), np.arange(5))
    b.slice(1, 3)
    assert b.size == 2 and b.rewards().tolist() == [1.0, 2.0]
    random.seed(0)
    b.shuffle()
    assert b.size == 2


def test_stacking_filter_and_reward_filters():
    from coach_b200.filters.filter import (LazyStack, ObservationStackingFilter, RewardClippingFilter,
                                           RewardRescaleFilter, ObservationToUInt8Filter)
    f = ObservationStackingFilter(4)
    o1 = f.filter(np.full((2, 2), 1.0))
    assert isinstance(o1, LazyStack)
    a = np.array(o1)
    assert a.shape == (2, 2, 4) and np.all(a == 1.0)          # first observation replicated, stacked on the last axis
    a = np.array(f.filter(np.full((2, 2), 2.0)))
    assert a[0, 0].tolist() == [1.0, 1.0, 1.0, 2.0]
    a = np.array(f.filter(np.full((2, 2), 3.0), update_internal_state=False))
    assert a[0, 0].tolist() == [1.0, 1.0, 1.0, 2.0]            # no state update
    f.reset()
    assert np.array(f.filter(np.full((2, 2), 9.0)))[0, 0].tolist() == [9.0] * 4
    with pytest.raises(ValueError):
        ObservationStackingFilter(0)
    c = RewardClippingFilter(-1.0, 1.0)
    assert c.filter(5) == 1.0 and c.filter(-3) == -1.0 and c.filter(0.25) == 0.25
    assert RewardClippingFilter(-1.0, 0).filter(5) == 5.0      # a bound of 0 is ignored (truthiness quirk)
    assert RewardRescaleFilter(5.0).filter(2) == 10.0
    with pytest.raises(ValueError):
        RewardRescaleFilter(0)
    assert ObservationToUInt8Filter(0, 1).filter(np.array([0.0, 0.999, 1.0])).tolist() == [0, 254, 255]


def test_plugin_path_strings_resolve():
    from coach_b200.utils import short_dynamic_import
    from coach_b200.memories.prioritized_experience_replay import PrioritizedExperienceReplayParameters
    from coach_b200.memories.experience_replay import ExperienceReplayParameters
    from coach_b200.agents.dqn_agent import DQNAgentParameters, DDQNAgentParameters
    for p in (PrioritizedExperienceReplayParameters(), ExperienceReplayParameters(), DQNAgentParameters(),
              DDQNAgentParameters()):
        cls = short_dynamic_import(p.path)
        assert cls.__name__ in p.path


def test_input_filter_chain_matches_reference_session():
    """InputFilter driven like Agent.observe (one environment response per step, reset at episode ends) and like
    Agent.train (a list of Transitions): replay of tests/golden/agent_prologues.npz, recorded from the reference's
    filters/filter.py:295-350 + to-uint8 / stacking / reward clipping / rescale filters (oracle/make_golden_agents.py)"""
    import os
    from types import SimpleNamespace
    from coach_b200.core_types import Transition
    from coach_b200.filters.filter import (InputFilter, ObservationStackingFilter, ObservationToUInt8Filter,
                                           RewardClippingFilter, RewardRescaleFilter)
    fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "agent_prologues.npz"))
    f = InputFilter()
    f.add_observation_filter('observation', 'to_uint8', ObservationToUInt8Filter(0, 1))
    f.add_observation_filter('observation', 'stacking', ObservationStackingFilter(4))
    f.add_reward_filter('rescale', RewardRescaleFilter(2.0))
    f.add_reward_filter('clipping', RewardClippingFilter(-1.0, 1.0))
    ends = set(int(e) for e in fx["flt_ends"])

    class EnvResponse(SimpleNamespace):          # an environment response: next_state / reward / game_over, no state
        pass
    for t in range(len(fx["flt_frames"])):
        er = EnvResponse(next_state={'observation': fx["flt_frames"][t]}, reward=float(fx["flt_rewards"][t]),
                         game_over=t in ends)
        res = f.filter(er)[0]
        got = np.array(res.next_state['observation'])
        assert got.dtype == fx["flt_stacked"].dtype and got.shape == fx["flt_stacked"][t].shape
        np.testing.assert_array_equal(got, fx["flt_stacked"][t])
        assert res.reward == fx["flt_filtered_rewards"][t]
        assert np.array_equal(er.next_state['observation'], fx["flt_frames"][t])          # input untouched (deep copy)
        if t in ends:
            f.reset()
    peek = f.filter(EnvResponse(next_state={'observation': fx["flt_frames"][0]}, reward=0.5, game_over=False),
                    update_internal_state=False)[0]
    np.testing.assert_array_equal(np.array(peek.next_state['observation']), fx["flt_peek"])
    g = InputFilter()
    g.add_observation_filter('observation', 'to_uint8', ObservationToUInt8Filter(0, 2))
    g.add_reward_filter('clipping', RewardClippingFilter(-1.0, 0))
    ts = [Transition(state={'observation': fx["flt_t_states"][i]}, action=0, reward=float(fx["flt_t_rewards"][i]),
                     next_state={'observation': fx["flt_t_next"][i]}, game_over=False) for i in range(7)]
    res = g.filter(ts)
    np.testing.assert_array_equal(np.array([t.state['observation'] for t in res]), fx["flt_t_states_out"])
    np.testing.assert_array_equal(np.array([t.next_state['observation'] for t in res]), fx["flt_t_next_out"])
    np.testing.assert_array_equal(np.array([t.reward for t in res]), fx["flt_t_rewards_out"])
