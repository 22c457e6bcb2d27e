"""N>1 path on CPU: world_size-2 gloo run of the chunk sharding + length exchange, with the emulation build of the kernels
standing in for the GPUs.  The concatenation of the ranks' outputs at the gathered offsets must equal the single-process
(and the reference's) file."""
import os
import socket

import numpy as np
import pytest

from tests.conftest import ROOT

EMU = os.path.join(ROOT, "tests", "emu", "librepaq_emu.so")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmpdir):
    import torch.distributed as dist
    from repaq_b200 import codec as K
    from repaq_b200 import shard
    from tools import fqgen
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["RPQ_EMU_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    r1, r2 = fqgen.generate(2400, seed=61, paired=True, threads=1)
    k = 100                                         # 100 kb chunks -> 334 pairs per chunk, 8 chunks
    upc = (100000 + 299) // 300
    n_pairs = 2400
    lo, hi = shard.unit_range(rank, world, n_pairs, upc)
    a1, b1 = shard.record_slice(r1, lo, hi)
    a2, b2 = shard.record_slice(r2, lo, hi)
    # header from chunk 0 on rank 0, broadcast
    hb = K.header_bytes(K.make_header(r1, r2, chunk_bases=k * 1000, lib_path=EMU), EMU) if rank == 0 else b""
    hb = shard.broadcast_header(hb)
    h, _ = K.parse_header(hb + bytes(8), EMU)
    h.support_interleaved = 1                        # not serialised; PE headers from paired Illumina names support it
    cd = K.Codec(lib_path=EMU)
    cd.set_header(h)
    data, infos, _ = cd.encode(r1[a1:b1], r2[a2:b2], chunk_bases=k * 1000, final=True)
    off, total, per_rank = shard.exchange_lengths([ci["bytes"] for ci in infos])
    assert sum(per_rank[rank]) == len(data)
    np.save(os.path.join(tmpdir, f"part{rank}.npy"), np.frombuffer(data, dtype=np.uint8))
    open(os.path.join(tmpdir, f"meta{rank}.txt"), "w").write(f"{off} {total} {len(hb)}")
    if rank == 0:
        open(os.path.join(tmpdir, "header.bin"), "wb").write(hb)
    cd.close()
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_file(tmp_path):
    import torch.multiprocessing as mp
    from oracle import oracle as O
    from tools import fqgen
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    hb = open(tmp_path / "header.bin", "rb").read()
    metas = [list(map(int, open(tmp_path / f"meta{r}.txt").read().split())) for r in range(2)]
    total = metas[0][1]
    body = bytearray(total)
    for r in range(2):
        part = np.load(tmp_path / f"part{r}.npy").tobytes()
        body[metas[r][0]:metas[r][0] + len(part)] = part
    r1, r2 = fqgen.generate(2400, seed=61, paired=True, threads=1)
    assert hb + bytes(body) == O.compress(r1, r2, chunk_bases=100000)


def test_chunk_ranges_cover_everything():
    from repaq_b200 import shard
    for n in (0, 1, 7, 8, 1430):
        for w in (1, 2, 4, 8):
            rs = shard.chunk_ranges(n, w)
            assert rs[0][0] == 0 and rs[-1][1] == n and all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in rs) - min(b - a for a, b in rs) <= 1
