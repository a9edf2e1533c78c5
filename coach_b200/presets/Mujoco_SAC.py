"""rl_coach/presets/Mujoco_SAC.py:27-42 (BASELINE config 4a).  The reward rescale (RewardRescaleFilter(5), an input
filter applied before the transition is stored) is ``reward_rescale`` below."""
from coach_b200.agents.soft_actor_critic_agent import SoftActorCriticAgentParameters
from coach_b200.filters.filter import InputFilter, RewardRescaleFilter

agent_params = SoftActorCriticAgentParameters()
for name in ('v', 'q', 'policy'):
    agent_params.network_wrappers[name].batch_size = 256
    agent_params.network_wrappers[name].learning_rate = 0.0003
    agent_params.network_wrappers[name].hidden_units = 256       # middleware [Dense(256)], Q head layers (256, 256)
agent_params.input_filter = InputFilter()
agent_params.input_filter.add_reward_filter('rescale', RewardRescaleFilter(5))

observation_dim, action_dim = 17, 6
