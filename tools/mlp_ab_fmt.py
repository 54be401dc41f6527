"""Condense the output of tools/mlp_ab.sh: ms of the policy / value net at 1 Mi, 4 Mi and 65 536 rockets per build."""
import json
import sys

for line in sys.stdin:
    line = line.strip()
    if line.startswith("{"):
        d = json.loads(line)
        print("   ", " | ".join(f"{n}: {d[n]['ms_policy_net']} {d[n]['ms_value_net']} (sampling {d[n].get('ms_policy_net_sampling', '-')}, one launch {d[n].get('ms_both_one_launch', '-')})" for n in ("1048576", "4194304", "65536")))
    elif line.startswith("==") or line.startswith("fused") or "passed" in line or "failed" in line:
        print(line)
