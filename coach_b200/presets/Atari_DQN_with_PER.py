"""rl_coach/presets/Atari_DQN_with_PER.py:14-17 (BASELINE config 2, the headline metric)"""
from coach_b200.agents.dqn_agent import DQNAgentParameters
from coach_b200.memories.prioritized_experience_replay import PrioritizedExperienceReplayParameters
from coach_b200.schedules import LinearSchedule

agent_params = DQNAgentParameters()
agent_params.network_wrappers['main'].learning_rate = 0.00025
agent_params.memory = PrioritizedExperienceReplayParameters()
agent_params.memory.beta = LinearSchedule(0.4, 1, 12500000)  # 12.5M training iterations = 50M steps = 200M frames

observation_shape, num_actions = (84, 84, 4), 6
