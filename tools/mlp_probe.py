"""Time rl_mlp_forward alone (both heads) at a few batch sizes.  GPU box only."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rocket_landing_rl_b200 import ActorCriticMLP, FusedActorCritic   # noqa: E402

net = ActorCriticMLP().cuda()
fused = FusedActorCritic(net)
out = {}
for n in (4096, 65536, 1 << 20, 1 << 22):
    obs = torch.randn((n, 8), device="cuda")
    for _ in range(3):
        fused.action_mean(obs); fused.predict_values(obs); fused.act_and_value(obs)
    e0, e1, e2, e3, e4 = (torch.cuda.Event(enable_timing=True) for _ in range(5))
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        fused.action_mean(obs)
    e1.record()
    for _ in range(10):
        fused.predict_values(obs)
    e2.record()
    for _ in range(10):
        fused.act_and_value(obs)                       # both nets (and the action sampling) in one launch
    e3.record()
    for _ in range(10):
        fused.act(obs)                                 # the policy net with the action sampling in its epilogue
    e4.record()
    torch.cuda.synchronize()
    ms_pi, ms_vf, ms_both = e0.elapsed_time(e1) / 10, e1.elapsed_time(e2) / 10, e2.elapsed_time(e3) / 10
    ms_act = e3.elapsed_time(e4) / 10
    flop = 2.0 * n * (16 * 256 + 256 * 256)           # tensor-core MACs x 2 (layer 1 padded to K = 16)
    out[n] = {"ms_policy_net": round(ms_pi, 4), "ms_value_net": round(ms_vf, 4), "ms_policy_net_sampling": round(ms_act, 4), "ms_both_one_launch": round(ms_both, 4),
              "tensor_TFLOPs_policy_net": round(flop / (ms_pi * 1e-3) / 1e12, 1),
              "rockets_per_s_both_nets": round(n / ((ms_pi + ms_vf) * 1e-3))}
print(json.dumps(out))
