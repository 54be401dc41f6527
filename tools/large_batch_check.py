"""One-off check of a batch near the top of what one B200 holds (805 306 368 rockets, ~100 GB of
state and output buffers): every plane and output array is far beyond 4 GiB, so any 32-bit byte
offset in the kernels would show.  Compares the last 1 Mi rockets and a 1 Mi slice in the middle,
after three steps, with the same rockets stepped in small shards (sharding does not change
trajectories: Philox is keyed by the global rocket id).  GPU box only."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rocket_landing_rl_b200 import RocketLandingVecEnv   # noqa: E402

n = 768 * 1024 * 1024
m = 1 << 20
env = RocketLandingVecEnv(n, seed=3)
env.reset()
g = torch.Generator(device="cuda").manual_seed(0)
acts = torch.rand((n, 2), device="cuda", generator=g)
acts[:, 1].mul_(2).sub_(1)
for _ in range(3):
    obs, rew, term, trunc, _ = env.step(acts)
torch.cuda.synchronize()
st = env.stats()
assert st.env_steps == 3.0 * n, st
for start in (n - m, n // 2 + 12345 * 256):
    small = RocketLandingVecEnv(m, seed=3, env_id_offset=start)
    small.reset()
    a = acts[start:start + m].contiguous()
    for _ in range(3):
        o2, r2, t2, tr2, _ = small.step(a)
    assert torch.equal(o2, obs[start:start + m]) and torch.equal(r2, rew[start:start + m])
    assert torch.equal(small.get_state(), env.get_state()[:, start:start + m])
    small.close()
print("ok:", n, "rockets,", torch.cuda.max_memory_allocated() / 2**30, "GiB allocated at peak")
