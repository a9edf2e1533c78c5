"""rl_coach/presets/Mujoco_TD3.py:26-35 (BASELINE config 4b): actor 17 -> 400 -> 300 -> A (tanh), critic two streams
(17 + A) -> 400 -> 300 -> 1 -- the network shapes are those of ``TD3Agent``."""
from coach_b200.agents.ddpg_agent import TD3AgentParameters

agent_params = TD3AgentParameters()

observation_dim, action_dim = 17, 6
