"""Output error of the fused MLP kernel against the fp32 PyTorch module, next to what tf32 and
bf16-autocast PyTorch lose on the same inputs (profiles/README.md).  GPU box only."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rocket_landing_rl_b200 import ActorCriticMLP, FusedActorCritic
torch.manual_seed(3)
net = ActorCriticMLP().cuda()
with torch.no_grad():
    for p in net.parameters():
        if p.ndim == 1: p.uniform_(-0.3, 0.3)
fused = FusedActorCritic(net)
obs = (torch.randn((262144, 8), device="cuda") * 3.0).clamp_(-10, 10)
with torch.no_grad():
    wm = net.action_net(net.policy_net(obs)); wv = net.predict_values(obs)
    torch.backends.cuda.matmul.allow_tf32 = True
    tm = net.action_net(net.policy_net(obs))
    with torch.autocast("cuda", dtype=torch.bfloat16):
        bm = net.action_net(net.policy_net(obs)).float()
gm = fused.action_mean(obs); gv = fused.predict_values(obs)
print("fused  max|err| mean %.2e value %.2e  rms mean %.2e" % ((gm-wm).abs().max(), (gv-wv).abs().max(), (gm-wm).pow(2).mean().sqrt()))
print("tf32   max|err| mean %.2e rms %.2e" % ((tm-wm).abs().max(), (tm-wm).pow(2).mean().sqrt()))
print("bf16ac max|err| mean %.2e rms %.2e" % ((bm-wm).abs().max(), (bm-wm).pow(2).mean().sqrt()))
print("output scale: mean abs %.3f" % wm.abs().mean())
