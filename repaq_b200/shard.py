"""Multi-GPU sharding of the codec: chunks are independent given the header (reference src/rfqcodec.cpp:163-170 reads
only mHeader and its own reads), so each rank takes a contiguous range of chunks and the only exchange is the gather of
per-chunk serialised lengths that turns local chunk offsets into file offsets.

torch.distributed is only the plumbing (NCCL on GPUs, gloo in the CPU tests)."""
import numpy as np


def chunk_ranges(n_chunks, world):
    """contiguous chunk range [lo, hi) of every rank, sizes differing by at most one"""
    base, extra = divmod(n_chunks, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def unit_range(rank, world, n_units, units_per_chunk):
    """records (or pairs) of this rank when every chunk holds `units_per_chunk` units (uniform read length, Q19)"""
    n_chunks = (n_units + units_per_chunk - 1) // units_per_chunk
    lo, hi = chunk_ranges(n_chunks, world)[rank]
    return min(lo * units_per_chunk, n_units), min(hi * units_per_chunk, n_units)


def record_slice(buf, first_record, last_record):
    """byte range of records [first, last) in a '\\n'-terminated FASTQ image"""
    nl = np.flatnonzero(np.frombuffer(buf, dtype=np.uint8) == 10) if not isinstance(buf, np.ndarray) else np.flatnonzero(buf == 10)
    a = 0 if first_record == 0 else int(nl[4 * first_record - 1]) + 1
    b = int(nl[4 * last_record - 1]) + 1 if last_record > 0 else 0
    return a, b


def exchange_lengths(chunk_bytes, device=None, group=None):
    """all_gather of per-chunk serialised lengths -> (file offset of this rank's first chunk, total body bytes,
    lengths of every rank).  `chunk_bytes`: list/array of this rank's chunk lengths in order."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        total = int(np.sum(chunk_bytes)) if len(chunk_bytes) else 0
        return 0, total, [list(map(int, chunk_bytes))]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = device if device is not None else "cpu"
    lens = torch.tensor(list(map(int, chunk_bytes)), dtype=torch.int64, device=dev)
    n = torch.tensor([lens.numel()], dtype=torch.int64, device=dev)
    ns = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(ns, n, group=group)
    mx = max(int(x) for x in ns)
    pad = torch.zeros(max(mx, 1), dtype=torch.int64, device=dev)
    pad[: lens.numel()] = lens
    allp = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(allp, pad, group=group)
    per_rank = [[int(v) for v in allp[r][: int(ns[r])]] for r in range(world)]
    offset = sum(sum(per_rank[r]) for r in range(rank))
    return offset, sum(sum(x) for x in per_rank), per_rank


def broadcast_header(header_bytes, src=0, device=None, group=None):
    """the <= 145-byte file header is built from chunk 0 by rank `src` and broadcast"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return header_bytes
    dev = device if device is not None else "cpu"
    t = torch.zeros(160, dtype=torch.uint8, device=dev)
    if dist.get_rank(group) == src:
        t[0] = len(header_bytes)
        t[1:1 + len(header_bytes)] = torch.tensor(list(header_bytes), dtype=torch.uint8)
    dist.broadcast(t, src=src, group=group)
    n = int(t[0])
    return bytes(t[1:1 + n].cpu().numpy())
