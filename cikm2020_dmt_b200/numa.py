"""Host-memory placement for the host->device copies of one process per GPU.

A pinned staging buffer is only as fast as the path from its DRAM pages to the GPU: when all ranks of a node
allocate their buffers on NUMA node 0, the copies of the GPUs behind the other socket cross the inter-socket link
and all ranks share one memory controller (SCALE_r01: end-to-end efficiency 0.58 at 8 GPUs).  `numa_local(i)` pins
the calling thread to the CPUs of GPU i's NUMA node for the duration of a `with` block; pages first touched / pinned
inside the block (torch pin_memory -> cudaHostAlloc) are allocated on that node under Linux's default local policy.
Plumbing only (sysfs + sched_setaffinity); no effect on what is computed.
"""
import contextlib
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


def gpu_numa_node(index):
    """NUMA node of CUDA device `index` (sysfs), or None when unknown (single-node host, container without sysfs)."""
    try:
        import torch
        p = torch.cuda.get_device_properties(index)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as fh:
            node = int(fh.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def node_cpus(node):
    try:
        with open("/sys/devices/system/node/node%d/cpulist" % node) as fh:
            return _parse_cpulist(fh.read())
    except Exception:
        return set()


@contextlib.contextmanager
def numa_local(index):
    """Run the block on the CPUs next to GPU `index`; yields the NUMA node (None: nothing was changed)."""
    node = gpu_numa_node(index)
    old = None
    try:
        if node is not None:
            cur = os.sched_getaffinity(0)
            want = node_cpus(node) & cur
            if want and want != cur:
                os.sched_setaffinity(0, want)
                old = cur
            elif not want:
                node = None
        yield node
    finally:
        if old is not None:
            os.sched_setaffinity(0, old)
