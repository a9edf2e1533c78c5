"""torchrun --nproc-per-node 2 tools/check_allreduce_overlap.py : runs a few DQN learn steps on every rank and prints a
hash of the parameters.  Run once with CB200_DQN_OVERLAP_ALLREDUCE=1 and once with =0: the hashes must agree (the
overlapped all-reduce of the dense-layer gradients changes the schedule, not the arithmetic)."""
import hashlib
import os
import random
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coach_b200 import parallel                                              # noqa: E402
from coach_b200.agents.dqn_agent import DQNAgent, DQNAgentParameters         # noqa: E402
from coach_b200.memories.memory import MemoryGranularity                     # noqa: E402
from coach_b200.memories.prioritized_experience_replay import PrioritizedExperienceReplayParameters   # noqa: E402

rank, world = parallel.init_from_env()
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
ap = DQNAgentParameters()
ap.memory = PrioritizedExperienceReplayParameters()
ap.memory.max_size = (MemoryGranularity.Transitions, 2048)
ap.network_wrappers["main"].batch_size = 128
agent = DQNAgent(ap, observation_shape=(84, 84, 4), num_actions=6, seed=5)       # same initial weights on all ranks
rng = np.random.RandomState(100 + rank)                                          # different replay shard per rank
n = 1024
agent.memory.store_columns({
    "state:observation": rng.randint(0, 256, (n, 84, 84, 4)).astype(np.uint8),
    "next_state:observation": rng.randint(0, 256, (n, 84, 84, 4)).astype(np.uint8),
    "action": rng.randint(0, 6, n).astype(np.int64), "reward": rng.randint(-1, 2, n).astype(np.float64),
    "game_over": (rng.rand(n) < 0.1).astype(np.uint8)})
losses = []
for step in range(6):
    random.seed(1000 * rank + step)
    batch = agent.sample_batch()
    loss, _, _ = agent.learn_from_batch(batch)
    losses.append(loss)
torch.cuda.synchronize()
theta = agent.net_def.store.theta.cpu().numpy()
print("rank %d overlap=%s graph=%s theta=%s losses=%s" % (
    rank, os.environ.get("CB200_DQN_OVERLAP_ALLREDUCE", "1"), agent._graphs is not None,
    hashlib.sha1(theta.tobytes()).hexdigest()[:16], ["%.6f" % l for l in losses]), flush=True)
torch.distributed.destroy_process_group()
