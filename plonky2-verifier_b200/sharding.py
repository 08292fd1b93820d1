"""Multi-GPU sharding of a proof batch.

Every proof is independent, so the batch is cut into contiguous slices, one per rank (one process per GPU);
each rank verifies its slice with its own Context/Circuit and the ONLY exchange is one all-gather of the packed
accept bitmap (NCCL over NVLink on GPUs; the same code runs over gloo for the CPU tests).  Slices are multiples
of 32 proofs so that bitmap words never straddle two ranks.
"""
import numpy as np


def slice_len(n_total, world):
    """Proofs per rank: ceil(n_total / world) rounded up to a multiple of 32."""
    per = (n_total + world - 1) // world
    return (per + 31) // 32 * 32


def shard_bounds(n_total, rank, world):
    """[start, stop) of this rank's contiguous slice (the last ranks may be short or empty)."""
    per = slice_len(n_total, world)
    start = min(n_total, rank * per)
    return start, min(n_total, start + per)


def pack_bits(flags):
    """bool[n] -> uint32 words, bit i%32 of word i/32 (the layout of p2v_verify_batch's accept_bits)."""
    flags = np.asarray(flags, dtype=bool)
    n = len(flags)
    padded = np.zeros((n + 31) // 32 * 32, dtype=np.uint32)
    padded[:n] = flags
    return (padded.reshape(-1, 32) << np.arange(32, dtype=np.uint32)).sum(axis=1, dtype=np.uint64).astype(np.uint32)


def gather_accept_bitmap(local_words, n_total, dist=None, group=None):
    """All-gather the per-rank bitmap words into the bitmap of the whole batch (torch tensors, any device).

    local_words: int32/uint32 tensor with slice_len(n_total, world)/32 words (zero padded)."""
    import torch

    if dist is None or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_words[: (n_total + 31) // 32]
    world = dist.get_world_size(group)
    words_per_rank = slice_len(n_total, world) // 32
    if local_words.numel() != words_per_rank:
        raise ValueError("rank bitmap has %d words, expected %d" % (local_words.numel(), words_per_rank))
    out = torch.empty(words_per_rank * world, dtype=local_words.dtype, device=local_words.device)
    dist.all_gather_into_tensor(out, local_words.contiguous(), group=group)
    return out[: (n_total + 31) // 32]


def verify_batch_sharded(circuit, blobs_local, n_total, rank, world, dist=None, group=None):
    """Verify this rank's slice on its GPU and return (bitmap of the WHOLE batch as a device tensor, local status).

    blobs_local: this rank's slice (host array or device tensor), AoS [n_local][blob_words]."""
    import torch

    start, stop = shard_bounds(n_total, rank, world)
    n_local = stop - start
    words_per_rank = slice_len(n_total, world) // 32
    dev = torch.device("cuda", circuit.ctx.device)
    bits = torch.zeros(words_per_rank, dtype=torch.int32, device=dev)
    status = torch.zeros(max(n_local, 1), dtype=torch.int32, device=dev)
    if n_local:
        circuit.verifyProof(blobs_local, n=n_local, accept_bits=bits, status=status)
    if dist is not None and dist.is_initialized() and world > 1:
        # the gather is enqueued on the verifier's stream: no host synchronisation in between
        with torch.cuda.stream(torch.cuda.ExternalStream(circuit.ctx.stream)):
            full = gather_accept_bitmap(bits, n_total, dist, group)
    else:
        full = bits[: (n_total + 31) // 32]
    return full, status[:n_local]


def gpu_numa_cpus(device):
    """CPUs of the NUMA node the GPU hangs off (None if the topology cannot be read).  Pinned staging buffers
    that are first-touched from these CPUs sit on the GPU's own node; on a two-socket host the other placement
    halves the H2D bandwidth of the end-to-end path."""
    try:
        import torch

        pr = torch.cuda.get_device_properties(device)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        return cpus or None
    except Exception:
        return None


def bind_to_gpu_numa_node(device):
    """Restrict this process to the CPUs next to `device`; returns the previous affinity (for os.sched_setaffinity)."""
    import os

    prev = os.sched_getaffinity(0)
    cpus = gpu_numa_cpus(device)
    if cpus and (cpus & prev):
        os.sched_setaffinity(0, cpus & prev)
    return prev
