"""Host placement for the host-pointer entry points: run this process on the CPUs of the GPU's NUMA node so that the
pinned staging buffers it allocates afterwards (first touch at cudaHostAlloc) are local to the GPU's PCIe root.

On an 8-GPU box the host side of `execute_block` is otherwise bound by the inter-socket link: every rank's pinned
memory lands on the node the launcher happened to start it on.
"""
import os


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_numa_node(device_index: int):
    """NUMA node of CUDA device `device_index` from sysfs, or None when it cannot be determined."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        addr = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % addr) as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def bind_to_gpu_numa_node(device_index: int):
    """Restrict this process to the CPUs of the GPU's NUMA node (intersected with the CPUs it may already use).
    Returns the node, or None when nothing was changed."""
    node = gpu_numa_node(device_index)
    if node is None:
        return None
    try:
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = _parse_cpulist(f.read())
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None
