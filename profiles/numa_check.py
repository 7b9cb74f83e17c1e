"""Is the PCIe floor of a box a question of NUMA placement? Prints the topology the process sees and the
full-size H2D + D2H floor with buffers pinned from every NUMA node's CPUs in turn."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ONE = r"""
import os, sys, torch
sys.path.insert(0, %r)
from pico_tree_b200 import hostmem
from bench import pcie_floor_ms
cpus = os.environ.get("CPUS")
if cpus:
    os.sched_setaffinity(0, hostmem._parse_cpulist(cpus) & os.sched_getaffinity(0))
elif os.environ.get("BIND"):
    print(hostmem.bind_to_gpu_node(0))
torch.cuda.init()
q = torch.empty((7200863, 3), dtype=torch.float32).pin_memory()
out = torch.empty((7200863, 1, 2), dtype=torch.int32).pin_memory()
print("affinity %%d cpus -> floor %%.3f ms" %% (len(os.sched_getaffinity(0)), pcie_floor_ms(q, out, torch.device("cuda", 0))))
""" % ROOT


def sh(cmd):
    r = subprocess.run(cmd, shell=True, capture_output=True, text=True)
    return (r.stdout + r.stderr).strip()


if __name__ == "__main__":
    print(sh("lscpu | grep -i -E 'numa|socket|model name|^CPU\\(s\\)'"))
    print("affinity:", sorted(os.sched_getaffinity(0)))
    print(sh("nvidia-smi topo -m | head -8"))
    from pico_tree_b200 import hostmem
    print("gpu numa node:", hostmem.gpu_numa_node(0))
    nodes = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))

    def one(**env):
        r = subprocess.run([sys.executable, "-c", ONE], capture_output=True, text=True, env=dict(os.environ, **env))
        return (r.stdout + r.stderr[-400:]).strip()

    for rep in range(2):
        print("unbound      :", one())
    for n in nodes:
        cpulist = open("/sys/devices/system/node/%s/cpulist" % n).read().strip()
        print("%s (%s):" % (n, cpulist), one(CPUS=cpulist))
    print("bind_to_gpu_node:", one(BIND="1"))
