"""Diagnostic: which NUMA node the GPU hangs off and what pinned H2D bandwidth each node gives."""
import glob
import os
import time

import torch


def cpulist(s):
    out = []
    for part in s.strip().split(","):
        if "-" in part:
            a, b = part.split("-")
            out += list(range(int(a), int(b) + 1))
        elif part:
            out.append(int(part))
    return out


def main():
    print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
    nodes = sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))
    print("numa nodes", [os.path.basename(n) for n in nodes])
    props = torch.cuda.get_device_properties(0)
    bus = "%04x:%02x:%02x.0" % (getattr(props, "pci_domain_id", 0), props.pci_bus_id, props.pci_device_id)
    try:
        print("gpu", bus, "numa_node", open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
    except OSError as e:
        print("gpu numa unknown", e)
    allowed = os.sched_getaffinity(0)
    dst = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for n in nodes + [None]:
        if n is not None:
            cpus = set(cpulist(open(n + "/cpulist").read())) & allowed
            if not cpus:
                print(os.path.basename(n), "no allowed cpus")
                continue
            os.sched_setaffinity(0, cpus)
        else:
            os.sched_setaffinity(0, allowed)
        src = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
        src.fill_(1)
        for _ in range(2):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(8):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        print(os.path.basename(n) if n else "all", "H2D GB/s %.1f" % (8 * 256 / 1024 / dt))
        del src


if __name__ == "__main__":
    main()
