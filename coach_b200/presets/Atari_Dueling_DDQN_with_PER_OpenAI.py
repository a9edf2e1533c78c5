"""rl_coach/presets/Atari_Dueling_DDQN_with_PER_OpenAI.py:14-27 (BASELINE config 5)"""
from coach_b200.agents.dqn_agent import DDQNAgentParameters
from coach_b200.base_parameters import EnvironmentSteps, MiddlewareScheme
from coach_b200.memories.prioritized_experience_replay import PrioritizedExperienceReplayParameters
from coach_b200.schedules import LinearSchedule

agent_params = DDQNAgentParameters()
agent_params.network_wrappers['main'].learning_rate = 0.0001
agent_params.network_wrappers['main'].middleware_parameters.scheme = MiddlewareScheme.Empty
agent_params.network_wrappers['main'].heads_parameters = ["DuelingQHead"]
agent_params.network_wrappers['main'].clip_gradients = 10
agent_params.algorithm.num_steps_between_copying_online_weights_to_target = EnvironmentSteps(40000)
agent_params.memory = PrioritizedExperienceReplayParameters()
agent_params.memory.beta = LinearSchedule(0.4, 1, 12500000)

observation_shape, num_actions = (84, 84, 4), 6
