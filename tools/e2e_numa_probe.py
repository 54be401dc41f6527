"""Does the NUMA placement of the pinned host buffers bound the host-buffer (e2e) path?
Prints the box topology, then times RocketLandingSB3VecEnv.step_raw with the process (and so the
first touch of every pinned buffer) bound to each NUMA node in turn.  GPU box only."""
import glob
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def cpulist(s):
    out = []
    for part in s.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        out.extend(range(int(a), int(b or a) + 1))
    return out


def measure(n, steps, dev):
    import numpy as np
    import torch
    from rocket_landing_rl_b200 import RocketLandingSB3VecEnv
    env = RocketLandingSB3VecEnv(n, device=dev, seed=0, build_infos=False)
    env.reset()
    a = torch.rand((n, 2)).pin_memory()
    a[:, 1].mul_(2).sub_(1)
    an = a.numpy()
    env.step_raw(an)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        obs, rew, term, trunc = env.step_raw(an)
        _ = float(rew[0])
    dt = time.perf_counter() - t0
    env.close()
    return n * steps / dt


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16777216
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    if len(sys.argv) > 2:                      # child: bind, measure, print
        cpus = cpulist(sys.argv[2]) if sys.argv[2] != "all" else sorted(os.sched_getaffinity(0))
        os.sched_setaffinity(0, cpus)
        print(json.dumps({"cpus": sys.argv[2], "e2e_env_steps_per_s": measure(n, 6, dev)}), flush=True)
        sys.exit(0)
    print(subprocess.run("nvidia-smi topo -m; lscpu | grep -i -E 'numa|socket|model name|^CPU\\(s\\)'", shell=True,
                         capture_output=True, text=True).stdout)
    nodes = {}
    for p in sorted(glob.glob("/sys/devices/system/node/node*/cpulist")):
        nodes[p.split("/")[-2]] = open(p).read().strip()
    print("nodes:", nodes)
    import torch
    prop = torch.cuda.get_device_properties(dev)
    bus = "%04x:%02x:%02x.0" % (getattr(prop, "pci_domain_id", 0), getattr(prop, "pci_bus_id", 0), getattr(prop, "pci_device_id", 0))
    try:
        print("gpu", dev, bus, "numa_node", open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip(),
              "local_cpulist", open(f"/sys/bus/pci/devices/{bus}/local_cpulist").read().strip())
    except OSError as e:
        print("no sysfs entry for", bus, e)
    print("affinity now:", len(os.sched_getaffinity(0)), "cpus")
    for name, cl in [("all", "all")] + list(nodes.items()):
        r = subprocess.run([sys.executable, __file__, str(n), cl], capture_output=True, text=True)
        print(name, r.stdout.strip() or r.stderr[-500:], flush=True)
