"""rl_coach/presets/Atari_C51.py:10-11 (Categorical DQN on Atari, uniform replay)"""
from coach_b200.agents.categorical_dqn_agent import CategoricalDQNAgentParameters

agent_params = CategoricalDQNAgentParameters()
agent_params.network_wrappers['main'].learning_rate = 0.00025

observation_shape, num_actions = (84, 84, 4), 6
