#!/usr/bin/env python
"""bench.py - FASTQ <-> .rfq throughput of the B200 path on BASELINE.json's headline workload.

One step = encode a paired-end NovaSeq-shape FASTQ batch to .rfq AND decode it back.  With N GPUs the N batches are ONE file:
contiguous chunk ranges per rank, the header of chunk 0 broadcast, per-chunk lengths gathered into file offsets (NCCL).
Untimed gates: every rank's first and last chunks bit-exact against the oracle, the decode byte-exact against the input, the
chunks at every seam decoded out of the assembled file, and (N=1) every shard of the cpu_baseline sample byte-identical to
what the unmodified reference binary wrote for it (`parity_checked_bytes`).
`configs` adds BASELINE's other shapes: configs[3] (one 64 GB file, strong scaling) and configs[4] (BGI single end).
  value : FASTQ GB/s of the round trip with inputs resident in HBM (FASTQ bytes / (t_encode + t_decode)), all ranks
  e2e   : the same through the C ABI with pinned HOST buffers (H2D + kernels + D2H inside the timed region)
  roofline / cpu_baseline : see DESIGN.md section "Measurement"
`--impl reference` times the reference's own CPU implementation (oracle/_ref/repaq, else the C port) instead.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "FASTQ GB/s encode+decode (bit-exact .rfq)"
UNIT = "GB/s"


def env_int(k, d):
    return int(os.environ.get(k, d))


# ---------------------------------------------------------------------------------------------------------------------
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def shard_records(buf, n_shards):
    """record-aligned shards of a '\\n'-terminated FASTQ image with 4-line records of equal count per shard"""
    nl = np.flatnonzero(buf == 10)
    n_rec = nl.size // 4
    per = n_rec // n_shards
    cuts = [0] + [int(nl[4 * per * (i + 1) - 1]) + 1 for i in range(n_shards)]
    return [buf[cuts[i]:cuts[i + 1]] for i in range(n_shards)]


def cpu_reference_roundtrip(r1, r2, cores, budget_s=None, check=None):
    """Times the reference CPU implementation (compress + decompress) on `cores` disjoint record-aligned shards run
    concurrently, files in /dev/shm.  Returns dict(value GB/s of the round trip, kind, cores, sample, seconds).
    check(i, shard1, shard2, rfq_bytes, dec1_bytes, dec2_bytes) is called (untimed) with what the reference made of shard i."""
    from oracle import oracle as O
    kind = "reference" if O.have_ref() else "port"
    shards1 = shard_records(r1, cores)
    shards2 = shard_records(r2, cores) if r2 is not None else [None] * cores
    total = sum(s.size for s in shards1) + (sum(s.size for s in shards2) if r2 is not None else 0)
    tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    t_enc = t_dec = 0.0
    try:
        if kind == "reference":
            for i in range(cores):
                shards1[i].tofile(os.path.join(tmp, f"a{i}.fq"))
                if r2 is not None:
                    shards2[i].tofile(os.path.join(tmp, f"b{i}.fq"))

            def run(cmds):
                t = time.perf_counter()
                ps = [subprocess.Popen(c, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for c in cmds]
                rc = [p.wait() for p in ps]
                assert all(r == 0 for r in rc), "reference binary failed"
                return time.perf_counter() - t
            enc = [[O.REF_BIN, "-c", "-i", f"{tmp}/a{i}.fq"] + (["-I", f"{tmp}/b{i}.fq"] if r2 is not None else []) + ["-o", f"{tmp}/o{i}.rfq"] for i in range(cores)]
            dec = [[O.REF_BIN, "-d", "-i", f"{tmp}/o{i}.rfq", "-o", f"{tmp}/d{i}.fq"] + (["-O", f"{tmp}/e{i}.fq"] if r2 is not None else []) for i in range(cores)]
            t_enc = run(enc)
            t_dec = run(dec)
            if check is not None:
                for i in range(cores):
                    rd = lambda name: np.fromfile(os.path.join(tmp, name), dtype=np.uint8)       # noqa: E731
                    check(i, shards1[i], shards2[i], rd(f"o{i}.rfq"), rd(f"d{i}.fq"), rd(f"e{i}.fq") if r2 is not None else None)
        else:
            res = [None] * cores
            dec_out = [None] * cores

            def work(i, phase):
                if phase == 0:
                    res[i] = O.compress(shards1[i], shards2[i])
                else:
                    dec_out[i] = O.decompress(res[i], pe_out=r2 is not None)
            for phase in (0, 1):
                t = time.perf_counter()
                th = [threading.Thread(target=work, args=(i, phase)) for i in range(cores)]
                [x.start() for x in th]
                [x.join() for x in th]
                if phase == 0:
                    t_enc = time.perf_counter() - t
                else:
                    t_dec = time.perf_counter() - t
            if check is not None:
                for i in range(cores):
                    d = dec_out[i]
                    d1, d2 = (d if r2 is not None else (d, None))
                    check(i, shards1[i], shards2[i], np.frombuffer(res[i], dtype=np.uint8), np.frombuffer(d1, dtype=np.uint8), None if d2 is None else np.frombuffer(d2, dtype=np.uint8))
    finally:
        subprocess.call(["rm", "-rf", tmp])
    return dict(value=total / 1e9 / (t_enc + t_dec), unit=UNIT, cores=cores, kind=kind,
                sample=f"{total / 1e6:.0f} MB of the same workload in {cores} record-aligned shards, one process each, files in /dev/shm",
                encode_gbs=total / 1e9 / t_enc, decode_gbs=total / 1e9 / t_dec, seconds=t_enc + t_dec)


def bind_near_gpu(gpu):
    """Runs this rank on the CPUs that are local to its GPU (NVML's affinity mask), so that the pinned buffers it allocates and the
    threads that feed the copies sit on the GPU's side of the host: with several ranks per host the e2e path is bound by host memory
    and the PCIe fabric, and a rank on the wrong socket pays for every byte twice.  RPQ_BENCH_NUMA=0 leaves the affinity alone.
    Returns the previous affinity (restored before the CPU baseline is timed), or None."""
    if os.environ.get("RPQ_BENCH_NUMA", "1") == "0":
        return None
    try:
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        idx = int(vis.split(",")[gpu]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else gpu
        h = nv.nvmlDeviceGetHandleByIndex(idx)
        words = nv.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        near = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        old = os.sched_getaffinity(0)
        want = near & old
        if len(want) >= 2 and want != old:
            os.sched_setaffinity(0, want)
            return old
    except Exception:
        pass
    return None


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region, sampled in-process through NVML (pynvml) every ~2 ms by a
    thread: the default timed region lasts ~80 ms, too short for `nvidia-smi -lms` to land a sample in it reliably.  Only
    samples taken between begin() and end() are reported; if NVML is unavailable, nvidia-smi is the fallback."""
    Q = "timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu, uuid=None):
        self.gpu, self.uuid = gpu, uuid
        self.nv = self.h = self.th = None
        self.on = False
        self.rows = []
        self.p = self.path = None
        self.t0 = self.t1 = None

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = None
            if self.uuid:
                for cand in (self.uuid, "GPU-" + self.uuid):
                    try:
                        h = nv.nvmlDeviceGetHandleByUUID(cand.encode() if isinstance(cand, str) else cand)
                        break
                    except Exception:
                        h = None
            if h is None:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
                idx = self.gpu
                if vis and all(x.strip().isdigit() for x in vis.split(",")):
                    idx = int(vis.split(",")[self.gpu])
                h = nv.nvmlDeviceGetHandleByIndex(idx)
            nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            self.nv, self.h = nv, h
            self.smax = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.run = True
            self.th = threading.Thread(target=self._loop, daemon=True)
            self.th.start()
            return
        except Exception:
            self.nv = None
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.gpu)],
                                      stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def _loop(self):
        nv, h = self.nv, self.h
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while self.run:
            if self.on:
                try:
                    self.rows.append((time.time(), float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), int(get_reasons(h))))
                except Exception:
                    pass
                time.sleep(0.002)
            else:
                time.sleep(0.0005)

    def begin(self):
        self.t0 = time.time()
        self.on = True

    def end(self):
        self.on = False
        self.t1 = time.time()

    def stop(self):
        if self.nv is not None:
            self.run = False
            self.th.join()
            nv = self.nv
            names = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4), ("hw_power_brake_slowdown", 0x80))
            reasons = sorted({n for _, _, r in self.rows for n, bit in names if r & bit})
            sm = [c for _, c, _ in self.rows if c > 0]
            try:
                nv.nvmlShutdown()
            except Exception:
                pass
            return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_min_mhz=min(sm) if sm else None, sm_max_mhz=self.smax, reasons=reasons, samples=len(sm),
                        sampled="NVML, every ~2 ms between the first and the last timed step")
        if not self.p:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.05)
        self.p.terminate()
        self.p.wait()
        import datetime
        rows = []
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[2]), float(f[3]), f[6:10]))
            except ValueError:
                continue
        os.unlink(self.path)
        inside = [r for r in rows if self.t0 is not None and self.t0 <= r[0] <= self.t1]
        where = "nvidia-smi, timed region"
        if not inside:                                  # a region shorter than the sampling period: the nearest samples around it
            inside = [r for r in rows if self.t0 is not None and self.t0 - 0.25 <= r[0] <= self.t1 + 0.05]
            where = "nvidia-smi, timed region +-0.25 s (region shorter than the sampling period)"
        sm, mx, reasons = [], 0, set()
        for _, clk, cmax, flags in inside:
            sm.append(clk)
            mx = max(mx, cmax)
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), flags):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [x for x in sm if x > 0]
        return dict(sm_mhz=float(np.median(busy)) if busy else None, sm_max_mhz=mx or None, reasons=sorted(reasons), samples=len(sm), sampled=where)


# ---------------------------------------------------------------------------------------------------------------------
def run_reference_arm(args, rank, world):
    """`--impl reference`: the reference's CPU path on this box's host cores, same metric/config.  Rank 0 only."""
    if rank != 0:
        return
    from tools import fqgen
    cores = host_cores()
    # bounded sample: ~24 MB of FASTQ per core and step (~0.5 s per core per step at ~0.1 GB/s round trip)
    pairs_per_core = env_int("RPQ_REF_PAIRS_PER_CORE", 33340)
    r1, r2 = fqgen.generate(pairs_per_core * cores, seed=2, paired=True)
    vals = []
    for it in range(args.warmup + args.steps):
        r = cpu_reference_roundtrip(r1, r2, cores)
        if it >= args.warmup:
            vals.append(r)
    secs = sum(v["seconds"] for v in vals)
    total = (r1.size + r2.size) * len(vals)
    value = total / 1e9 / secs
    last = vals[-1]
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * secs / len(vals),
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="u8", data="synthetic", impl="reference",
                config=dict(workload="paired-end NovaSeq-shape 150bp (configs[1]/[2] shape), bounded sample per step", sample_bytes_per_step=int(r1.size + r2.size), chunk_kb=1000),
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind=last["kind"], sample=last["sample"], encode_gbs=last["encode_gbs"], decode_gbs=last["decode_gbs"]),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))


UPC_150 = 3334                 # pairs per chunk at 150 bp and -k 1000 (Q19: the pair that reaches 1 000 000 bases closes the chunk)


def chunk_lengths(eo):
    """per-chunk serialised lengths of an rpq_encode_out as a numpy array - no Python loop over the chunk table"""
    from repaq_b200 import _lib
    if eo.n_chunks == 0:
        return np.zeros(0, dtype=np.int64)
    arr = np.ctypeslib.as_array((_lib.ChunkInfo * eo.n_chunks).from_address(C.addressof(eo.chunks.contents)))
    return arr["bytes"].astype(np.int64)


def chunk_reads(eo):
    from repaq_b200 import _lib
    arr = np.ctypeslib.as_array((_lib.ChunkInfo * eo.n_chunks).from_address(C.addressof(eo.chunks.contents)))
    return arr["reads"].astype(np.int64)


def generate_pairs(fqgen, lo, hi, seed, threads):
    """records [lo, hi) of the paired NovaSeq-shape generator stream `seed` (rows of 300 pairs)"""
    row0, row1 = lo // fqgen.ROW_READS, (hi + fqgen.ROW_READS - 1) // fqgen.ROW_READS
    r1, r2 = fqgen.generate((row1 - row0) * fqgen.ROW_READS, seed=seed, paired=True, first_row=row0, threads=threads)
    if lo % fqgen.ROW_READS == 0 and hi == row1 * fqgen.ROW_READS:
        return r1, r2
    out = []
    for r in (r1, r2):
        nl = np.flatnonzero(r == 10)
        a = lo - row0 * fqgen.ROW_READS
        z = hi - row0 * fqgen.ROW_READS
        start = 0 if a == 0 else int(nl[4 * a - 1]) + 1
        out.append(r[start:int(nl[4 * z - 1]) + 1])
    return out[0], out[1]


class LengthExchange:
    """The one exchange of the sharded encode: every rank's per-chunk serialised lengths -> every rank's
    offset in the file.  One fixed-size all_gather per step, issued by a helper thread on a side stream while the main thread is
    inside the next C-ABI call; nothing of it is waited for before the timed region ends."""

    def __init__(self, torch, dist, world, rank, cap, device):
        import queue
        self.torch, self.dist, self.world, self.rank, self.cap = torch, dist, world, rank, cap
        self.stream = torch.cuda.Stream(device=device)
        self.host = torch.zeros(cap + 1, dtype=torch.int64).pin_memory()
        self.dev = torch.zeros(cap + 1, dtype=torch.int64, device=device)
        self.all = torch.zeros((world, cap + 1), dtype=torch.int64, device=device)
        self.q = queue.Queue()
        self.done = threading.Semaphore(0)
        self.device = device
        self.th = threading.Thread(target=self._loop, daemon=True)
        self.th.start()
        self.steps = 0

    def _loop(self):
        torch = self.torch
        torch.cuda.set_device(self.device)
        while True:
            lens = self.q.get()
            if lens is None:
                return
            self.stream.synchronize()                      # the previous copy out of the pinned buffer is done
            n = int(lens.size)
            self.host[0] = n
            self.host[1:1 + n] = torch.from_numpy(lens)
            with torch.cuda.stream(self.stream):
                self.dev.copy_(self.host, non_blocking=True)
                if self.world > 1:
                    self.dist.all_gather_into_tensor(self.all.view(-1), self.dev)
                else:
                    self.all[0].copy_(self.dev)
            self.done.release()

    def submit(self, lens):
        self.steps += 1
        self.q.put(lens)

    def drain(self):
        """all submitted exchanges issued and complete on the device"""
        while self.steps:
            self.done.acquire()
            self.steps -= 1
        self.stream.synchronize()

    def offsets(self):
        """-> (offset of this rank's first chunk in the body, total body bytes, chunks per rank) from the last exchange"""
        a = self.all.cpu().numpy()
        sums = [int(a[r, 1:1 + int(a[r, 0])].sum()) for r in range(self.world)]
        return sum(sums[:self.rank]), sum(sums), [int(a[r, 0]) for r in range(self.world)], sums

    def close(self):
        self.q.put(None)
        self.th.join()


def link_probe(torch, dist, world, local, nbytes=1 << 30, reps=3):
    """pinned H2D and D2H copies of `nbytes` on two streams AT THE SAME TIME, on all ranks at once -> (h2d GB/s, d2h GB/s) of this
    rank under that load: what the e2e path could reach if it did nothing but its copies"""
    h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    d_b = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    best = None
    for it in range(reps + 1):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        with torch.cuda.stream(s1):
            ev[0].record()
            d_a.copy_(h_in, non_blocking=True)
            ev[1].record()
        with torch.cuda.stream(s2):
            ev[2].record()
            h_out.copy_(d_b, non_blocking=True)
            ev[3].record()
        torch.cuda.synchronize()
        if it == 0:
            continue
        r = (nbytes / 1e9 / (ev[0].elapsed_time(ev[1]) / 1e3), nbytes / 1e9 / (ev[2].elapsed_time(ev[3]) / 1e3))
        best = r if best is None or r[0] + r[1] > best[0] + best[1] else best
    del h_in, h_out, d_a, d_b
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=env_int("RPQ_BENCH_PAIRS", 4760000), help="read pairs per GPU (4.76 M = the 3.4 GB nova pair)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-roofline", action="store_true", help="skip the per-kernel profiling pass (experiments)")
    ap.add_argument("--no-extra", action="store_true", help="skip the configs[3] (64 GB strong scaling) and configs[4] (BGI shape) blocks")
    ap.add_argument("--strong-gb", type=float, default=float(os.environ.get("RPQ_BENCH_STRONG_GB", 64)), help="total FASTQ GB of the strong-scaling job (configs[3])")
    ap.add_argument("--bgi-gb", type=float, default=float(os.environ.get("RPQ_BENCH_BGI_GB", 4.2)), help="FASTQ GB of the BGI-shape encode (configs[4] shape)")
    args = ap.parse_args()

    rank = env_int("RANK", 0)
    world = env_int("WORLD_SIZE", 1)
    local = env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from oracle import oracle as O            # checker only: never inside a timed region
    from repaq_b200 import codec as K
    from repaq_b200 import shard
    from tools import fqgen

    old_affinity = bind_near_gpu(local)
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    W = max(3, args.warmup)
    gen_threads = max(1, host_cores() // max(1, world))
    chunk_bases = 1000000

    # ---- the job: ONE paired-end file of world x (pairs per GPU) pairs + a partial last chunk, cut into contiguous chunk ranges
    # (shard.unit_range); rank r generates exactly its records of the generator's stream (weak scaling: per-GPU work is fixed)
    chunks_per_gpu = max(1, args.pairs // UPC_150)
    tail_pairs = 1000                                           # the file's last chunk is a partial one (the post-loop flush, src/repaq.cpp:710-754)
    n_pairs_total = chunks_per_gpu * UPC_150 * world + tail_pairs
    lo, hi = shard.unit_range(rank, world, n_pairs_total, UPC_150)
    t0 = time.perf_counter()
    r1, r2 = generate_pairs(fqgen, lo, hi, 2, gen_threads)
    gen_s = time.perf_counter() - t0
    fastq_bytes = int(r1.size + r2.size)
    # header from chunk 0 (rank 0 holds it), broadcast
    hb = K.header_bytes(K.make_header(r1, r2, chunk_bases=chunk_bases)) if rank == 0 else b""
    hb = shard.broadcast_header(hb, device=device)
    header, used = K.parse_header(hb + bytes(8))
    assert used == len(hb)
    enc, dec = K.Codec(device=local), K.Codec(device=local)
    enc.set_header(header)
    dec.set_header(header)

    d1 = torch.from_numpy(r1).cuda()
    d2 = torch.from_numpy(r2).cuda()
    torch.cuda.synchronize()
    es = torch.cuda.ExternalStream(enc.L.rpq_stream(enc.ctx))
    ds = torch.cuda.ExternalStream(dec.L.rpq_stream(dec.ctx))
    xch = LengthExchange(torch, dist, world, rank, chunks_per_gpu + 8, device)

    state = {}

    def step_device(exchange=True):
        eo = enc.encode_raw(d1.data_ptr(), d1.numel(), d2.data_ptr(), d2.numel(), 1, False, chunk_bases, True, (K.NEVER, K.NEVER), 0, 1)
        se = enc.stats()
        if exchange:
            xch.submit(chunk_lengths(eo))                      # gathered on a side stream while the decoder runs
        do = dec.decode_raw(eo.data, eo.bytes, 1, True, 1)
        sd = dec.stats()
        state.update(rfq_bytes=int(eo.bytes), n_chunks=int(eo.n_chunks), eo=eo, do=do, out=(int(do.out1_bytes), int(do.out2_bytes)), dec_walk=int(sd.dec_walk))
        return se.ms_total, sd.ms_total, se.launches + sd.launches, eo

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    cudart = C.cdll.LoadLibrary("libcudart.so.12")

    def from_device(ptr, n):
        t = torch.empty(n, dtype=torch.uint8, device="cuda")
        cudart.cudaMemcpy(C.c_void_p(t.data_ptr()), C.c_void_p(ptr), C.c_size_t(n), 3)
        return t

    # ---- correctness gate (not timed)
    _, _, _, eo = step_device()
    xch.drain()
    my_off, body_total, chunks_per_rank, bytes_per_rank = xch.offsets()
    rfq_dev = from_device(eo.data, state["rfq_bytes"])
    lens = chunk_lengths(eo)
    reads = chunk_reads(eo)
    assert int(lens.sum()) == state["rfq_bytes"] and bytes_per_rank[rank] == state["rfq_bytes"] and int(reads.sum()) == 2 * (hi - lo)
    # (1) every rank: its first and its last six chunks are the oracle's, encoded with the broadcast header
    nl1, nl2 = np.flatnonzero(r1 == 10), np.flatnonzero(r2 == 10)

    def records(first_pair, n_pairs):
        a1 = 0 if first_pair == 0 else int(nl1[4 * first_pair - 1]) + 1
        a2 = 0 if first_pair == 0 else int(nl2[4 * first_pair - 1]) + 1
        return bytes(r1[a1:int(nl1[4 * (first_pair + n_pairs) - 1]) + 1]), bytes(r2[a2:int(nl2[4 * (first_pair + n_pairs) - 1]) + 1])
    n_chk = min(6, state["n_chunks"])
    oracle_checked = 0
    for first_chunk in sorted({0, state["n_chunks"] - n_chk}):
        pair0 = int(reads[:first_chunk].sum()) // 2
        npairs = int(reads[first_chunk:first_chunk + n_chk].sum()) // 2
        p1, p2 = records(pair0, npairs)
        ref = O.compress_with_header(hb + bytes(8), p1, p2, chunk_bases=chunk_bases)
        a = int(lens[:first_chunk].sum())
        got = bytes(rfq_dev[a:a + int(lens[first_chunk:first_chunk + n_chk].sum())].cpu().numpy())
        assert got == ref, "encode is not bit-exact against the oracle (rank %d, chunks %d..)" % (rank, first_chunk)
        oracle_checked += len(p1) + len(p2)
    # (2) decode restores this rank's input byte for byte
    o1, o2 = from_device(state["do"].out1, state["out"][0]), from_device(state["do"].out2, state["out"][1])
    assert torch.equal(o1, d1) and torch.equal(o2, d2), "decode does not restore the input"
    del o1, o2
    # (3) the ranks' parts form ONE file: every rank writes its chunks at its gathered offset; rank 0 checks that the total is the
    # sum, that the offsets are contiguous, and decodes the two chunks at every seam out of the assembled file
    seam_note = None
    if world > 1:
        path = "/dev/shm/rpq_bench_%s.rfq" % os.environ.get("MASTER_PORT", "0")
        if rank == 0:
            with open(path, "wb") as f:
                f.write(hb)
                f.truncate(len(hb) + body_total)
        barrier()
        part = rfq_dev.cpu().numpy()
        with open(path, "r+b") as f:
            f.seek(len(hb) + my_off)
            f.write(part.tobytes())
        del part
        # the chunk on either side of my upper seam: lengths and the pairs they hold, for rank 0
        seam = torch.tensor([int(lens[0]), int(reads[0]), int(lens[-1]), int(reads[-1]), lo, hi], dtype=torch.int64, device="cuda")
        seams = [torch.zeros_like(seam) for _ in range(world)]
        dist.all_gather(seams, seam)
        barrier()
        if rank == 0:
            assert os.path.getsize(path) == len(hb) + sum(bytes_per_rank)
            offs = np.concatenate([[0], np.cumsum(bytes_per_rank)])
            chk = K.Codec(device=local)
            chk.set_header(header)
            with open(path, "rb") as f:
                for r in range(world - 1):
                    last_len, last_reads = int(seams[r][2]), int(seams[r][3])
                    first_len, first_reads = int(seams[r + 1][0]), int(seams[r + 1][1])
                    f.seek(len(hb) + int(offs[r + 1]) - last_len)
                    two = f.read(last_len + first_len)
                    a1, a2, infos, _ = chk.decode(two, split_pairs=True)
                    assert len(infos) == 2 and infos[0]["reads"] == last_reads and infos[1]["reads"] == first_reads
                    seam_pair = int(seams[r][5])                 # first pair of rank r + 1
                    e1, e2 = generate_pairs(fqgen, seam_pair - last_reads // 2, seam_pair + first_reads // 2, 2, gen_threads)
                    assert a1 == bytes(e1) and a2 == bytes(e2), "the chunks at the seam of ranks %d|%d do not decode to the file's records" % (r, r + 1)
            chk.close()
            os.unlink(path)
            seam_note = "%d seams decoded from the assembled file" % (world - 1)
        barrier()
    del rfq_dev

    # ---- timed: inputs resident in HBM (inputs are ~3.4 GB per step: far larger than the 126 MB L2, no flush needed)
    try:
        uuid = str(torch.cuda.get_device_properties(local).uuid)
    except Exception:
        uuid = None
    clocks = ClockSampler(local, uuid)
    if rank == 0:
        clocks.start()
    for _ in range(W):
        step_device()
    xch.drain()
    barrier()
    clocks.begin()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall = time.perf_counter()
    ev0.record(es)
    enc_ms = dec_ms = 0.0
    launches = 0
    for _ in range(args.steps):
        a, b_, l, eo = step_device()
        enc_ms += a
        dec_ms += b_
        launches += l
    ev1.record(ds)
    xch.drain()                                      # every step's length exchange has completed inside the timed region
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t_wall)
    clocks.end()
    dev_ms = ev0.elapsed_time(ev1)                  # CUDA events: first encode op .. last decode op, host gaps included
    clk = clocks.stop() if rank == 0 else None
    t = torch.tensor([dev_ms, wall_ms, enc_ms, dec_ms], dtype=torch.float64, device="cuda")
    tot_bytes = torch.tensor([fastq_bytes, state["rfq_bytes"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot_bytes, op=dist.ReduceOp.SUM)
    dev_ms, wall_ms, enc_ms, dec_ms = [float(x) for x in t]
    job_bytes, job_rfq = float(tot_bytes[0]), float(tot_bytes[1])
    value = job_bytes * args.steps / 1e9 / (max(dev_ms, 1e-9) / 1e3)
    my_off2, body_total2, _, _ = xch.offsets()
    assert (my_off2, body_total2) == (my_off, body_total)        # the exchange of the last timed step gives the same file layout

    # ---- per-kernel device time (separate pass, event pair around every launch) -> roofline of the dominant kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

    def profile_pass(fn):
        enc.set_profiling(True)
        dec.set_profiling(True)
        fn()
        prof = {}
        for k, (n, ms) in list(enc.profile().items()) + list(dec.profile().items()):
            prof[k] = (n, ms)
        enc.set_profiling(False)
        dec.set_profiling(False)
        return prof

    def alg_bytes(fq_b, rfq_b, n_reads, bases):
        """algorithmic bytes per launch of each kernel: what it must read + write once"""
        qual_b = seq_b = bases
        head_b = fq_b - qual_b - n_reads                                        # name + sequence + strand lines
        stream_b = max(0.0, rfq_b - seq_b / 4 - 8 * n_reads)                    # position streams ~ the .rfq minus 2-bit bases and per-read columns
        return {
            "k_index_lines": fq_b + 4 * 4 * n_reads,                             # text in, line index out
            "k_unit_lengths": (16 + 16 + 4 + 4) * n_reads,                       # line index in, record index + lengths out
            "k_meta3": head_b + (16 + 44) * n_reads,                             # record heads in, ReadMeta + packed read out
            "k_streams3": qual_b + stream_b, "k_streams4": qual_b + stream_b, "k_streams7": qual_b + stream_b,   # qualities in, tokens out
            "k_emit2": 44 * n_reads + seq_b / 4,                                 # packed reads in, 2-bit stream out
            "k_emit_names": 2 * (fq_b - 2 * bases), "k_gather": 2 * stream_b,
            "k_chunk_finish": 44 * n_reads, "k_coords": 9 * n_reads,
            "k_dec_format4": rfq_b + qual_b + fq_b,                              # columns + quality tiles in, text out
            "k_dec_planes": stream_b + qual_b,                                   # position streams in, quality tiles out
            "k_dec_qindex": stream_b * (1 + 8 / 128),                            # position streams in, a checkpoint per 128 bytes out
            "k_dec_coords3": (1 + 8) * n_reads, "k_dec_reads": 48 * n_reads,
        }

    def roofline_of(prof, alg, fq_b, rfq_b, traffic=None, both_ways=True):
        kern_total = sum(ms for _, ms in prof.values())
        name, (n, ms) = max(prof.items(), key=lambda kv: kv[1][1])
        ab = float(alg.get(name, fq_b)) / n                                     # the table holds bytes per step; per launch like the time
        ach = ab / 1e9 / (ms / n / 1e3)
        pipe_b = (fq_b + rfq_b) * (2 if both_ways else 1)
        return dict(bound="hbm", kernel=name, achieved=ach, peak=peak, unit="GB/s", frac=ach / peak, traffic=(traffic or {}).get(name), peak_source=peak_src,
                    traffic_source="profiles/r02_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu capture of this workload" if traffic and name in traffic else None,
                    share_of_kernel_time=ms / kern_total, algorithmic_bytes_per_launch=ab,
                    pipeline=dict(achieved=pipe_b / 1e9 / (kern_total / 1e3), frac=pipe_b / 1e9 / (kern_total / 1e3) / peak,
                                  note="(F+R) algorithmic bytes per direction over the sum of all kernel times"),
                    kernels_ms={k: round(v[1], 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])})

    rfq_b = state["rfq_bytes"]
    roofline = None
    if not args.no_roofline:
        prof = profile_pass(lambda: step_device(exchange=False))
        traffic = {}
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
            if abs(int(tj.get("pairs_per_gpu", 0)) - (hi - lo)) <= 2 * UPC_150:
                traffic = {k: v["traffic_bytes_per_launch"] for k, v in tj["kernels"].items()}
        except Exception:
            pass
        roofline = roofline_of(prof, alg_bytes(fastq_bytes, rfq_b, 2 * (hi - lo), 2 * (hi - lo) * 150), fastq_bytes, rfq_b, traffic)

    # ---- configs[3]: ONE 64 GB-shape paired-end file, chunk-sharded over the ranks (strong scaling: the total is fixed, a rank takes
    # 1/N of the chunks, in batches of at most 2.7 GB of text per file pair that are generated, encoded with the header of chunk 0,
    # decoded and compared one after the other).  Device time only (CUDA events of the library); generation and H2D are not timed.
    extra = {}
    if not args.no_extra and args.strong_gb > 0:
        pairs_total = int(args.strong_gb * 1e9 / 716) // UPC_150 * UPC_150 + tail_pairs       # ~716 bytes of text per pair
        slo, shi = shard.unit_range(rank, world, pairs_total, UPC_150)
        batch_pairs = 1130 * UPC_150                                                      # 3.77 M pairs ~ 2.7 GB per batch
        hb64 = K.header_bytes(K.make_header(*generate_pairs(fqgen, 0, UPC_150, 3, gen_threads), chunk_bases=chunk_bases)) if rank == 0 else b""
        hb64 = shard.broadcast_header(hb64, device=device)
        h64, _ = K.parse_header(hb64 + bytes(8))
        enc.set_header(h64)
        dec.set_header(h64)
        t_enc = t_dec = 0.0
        fq64 = rfq64 = 0
        nb = 0
        lens64 = []
        for a in range(slo, shi, batch_pairs):
            z = min(shi, a + batch_pairs)
            b1, b2 = generate_pairs(fqgen, a, z, 3, gen_threads)
            g1, g2 = torch.from_numpy(b1).cuda(), torch.from_numpy(b2).cuda()
            torch.cuda.synchronize()
            eo = enc.encode_raw(g1.data_ptr(), g1.numel(), g2.data_ptr(), g2.numel(), 1, False, chunk_bases, True, (K.NEVER, K.NEVER), 0, 1)
            t_enc += enc.stats().ms_total
            assert int(chunk_reads(eo).sum()) == 2 * (z - a)
            lens64.append(chunk_lengths(eo))
            do = dec.decode_raw(eo.data, eo.bytes, 1, True, 1)
            t_dec += dec.stats().ms_total
            q1, q2 = from_device(do.out1, int(do.out1_bytes)), from_device(do.out2, int(do.out2_bytes))
            assert torch.equal(q1, g1) and torch.equal(q2, g2), "64 GB job: decode does not restore the input"
            fq64 += int(b1.size + b2.size)
            rfq64 += int(eo.bytes)
            nb += 1
            del g1, g2, q1, q2, b1, b2
        xch64 = LengthExchange(torch, dist, world, rank, pairs_total // UPC_150 // world + 16, device)
        xch64.submit(np.concatenate(lens64) if lens64 else np.zeros(0, dtype=np.int64))
        xch64.drain()
        off64, total64, chunks64, _ = xch64.offsets()
        xch64.close()
        tt = torch.tensor([t_enc, t_dec], dtype=torch.float64, device="cuda")
        bb = torch.tensor([fq64, rfq64], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(bb, op=dist.ReduceOp.SUM)
        assert int(bb[1]) == total64 and sum(chunks64) == (pairs_total + UPC_150 - 1) // UPC_150
        extra["strong_64gb"] = dict(workload="configs[3]: ONE paired-end NovaSeq-shape 150bp file of %.1f GB, chunk-sharded over %d GPU(s), header of chunk 0 broadcast, "
                                             "per-chunk lengths gathered into file offsets" % (float(bb[0]) / 1e9, world),
                                    scaling="strong", fastq_bytes=int(bb[0]), rfq_bytes=int(bb[1]), chunks=sum(chunks64), batches_per_gpu=nb,
                                    encode_gbs=float(bb[0]) / 1e9 / (float(tt[0]) / 1e3), decode_gbs=float(bb[0]) / 1e9 / (float(tt[1]) / 1e3),
                                    roundtrip_gbs=float(bb[0]) / 1e9 / ((float(tt[0]) + float(tt[1])) / 1e3),
                                    encode_ms=float(tt[0]), decode_ms=float(tt[1]), timing="device time of the library calls (CUDA events), max over ranks; inputs resident in HBM",
                                    verified="every batch decoded back to its input byte for byte; chunk count and file size from the gathered lengths")
        enc.set_header(header)
        dec.set_header(header)

    # ---- configs[4] shape: BGI-SEQ single end 100 bp (names without lane/tile/x/y, ~40 quality values), 1 x B200 encode, in batches
    # of at most 2.1 GB of text; rank 0 only
    if not args.no_extra and args.bgi_gb > 0 and rank == 0:
        reads_total = int(args.bgi_gb * 1e9 / 235) // 10000 * 10000 + 4321                 # ~235 bytes of text per read; 10 000 reads per chunk
        batch_reads = 900 * 10000
        hdr_b = None
        t_enc = t_dec = 0.0
        fqb = rfb = 0
        bprof = {}
        nbatch = 0
        bgi_oracle = 0
        for a in range(0, reads_total, batch_reads):
            z = min(reads_total, a + batch_reads)
            row0, row1 = a // fqgen.ROW_READS, (z + fqgen.ROW_READS - 1) // fqgen.ROW_READS
            t1, _ = fqgen.generate((row1 - row0) * fqgen.ROW_READS, seed=5, shape=fqgen.BGI, first_row=row0, threads=host_cores())
            if z != row1 * fqgen.ROW_READS:
                t1 = fqgen.truncate_reads(t1, z - row0 * fqgen.ROW_READS)
            if hdr_b is None:
                hdr_b = K.make_header(t1, chunk_bases=chunk_bases)
                enc.set_header(hdr_b)
                dec.set_header(hdr_b)
            g1 = torch.from_numpy(t1).cuda()
            torch.cuda.synchronize()
            for rep in range(2 if a == 0 else 1):                                          # the first batch once more: warm buffers
                eo = enc.encode_raw(g1.data_ptr(), g1.numel(), None, 0, 1, False, chunk_bases, True, (K.NEVER, K.NEVER), 0, 1)
            t_enc += enc.stats().ms_total
            if a == 0:
                # the first chunks against the oracle
                n6 = int(chunk_reads(eo)[:6].sum())
                ref = O.compress(bytes(fqgen.truncate_reads(t1, n6)), chunk_bases=chunk_bases)
                got = bytes(from_device(eo.data, int(chunk_lengths(eo)[:6].sum())).cpu().numpy())
                assert K.header_bytes(hdr_b) + got == ref, "BGI shape: encode is not bit-exact against the oracle"
                bgi_oracle = int(fqgen.truncate_reads(t1, n6).size)
                if not args.no_roofline:
                    enc.set_profiling(True)
                    enc.encode_raw(g1.data_ptr(), g1.numel(), None, 0, 1, False, chunk_bases, True, (K.NEVER, K.NEVER), 0, 1)
                    bprof = dict(enc.profile())
                    enc.set_profiling(False)
                    eo = enc.encode_raw(g1.data_ptr(), g1.numel(), None, 0, 1, False, chunk_bases, True, (K.NEVER, K.NEVER), 0, 1)
            for rep in range(2 if a == 0 else 1):                                          # the first batch once more: the decoder's buffers exist
                do = dec.decode_raw(eo.data, eo.bytes, 1, False, 1)
            t_dec += dec.stats().ms_total
            consumed = int(eo.r1_consumed)
            q1 = from_device(do.out1, int(do.out1_bytes))
            assert torch.equal(q1, g1[:consumed]), "BGI shape: decode does not restore the input"
            assert consumed == g1.numel(), "BGI shape: batches are cut at chunk ends"       # 900 chunks of 10 000 reads per batch
            fqb += consumed
            rfb += int(eo.bytes)
            nbatch += 1
            bgi_first = (int(t1.size), int(eo.bytes), z - a) if a == 0 else bgi_first
            del g1, q1, t1
        blk = dict(workload="configs[4] shape: BGI-SEQ single end 100 bp, %.2f GB FASTQ in %d batches of <= 2.1 GB, 1 x B200 encode (and decode back)" % (fqb / 1e9, nbatch),
                   fastq_bytes=fqb, rfq_bytes=rfb, rfq_ratio=rfb / fqb, encode_gbs=fqb / 1e9 / (t_enc / 1e3), decode_gbs=fqb / 1e9 / (t_dec / 1e3),
                   roundtrip_gbs=fqb / 1e9 / ((t_enc + t_dec) / 1e3), encode_ms=t_enc, decode_ms=t_dec, oracle_checked_bytes=bgi_oracle,
                   verified="every batch decoded back to its input byte for byte; the first six chunks bit-exact against the oracle")
        if bprof:
            f1, rq1, n1 = bgi_first
            blk["roofline"] = roofline_of(bprof, alg_bytes(f1, rq1, n1, n1 * 100), f1, rq1, both_ways=False)
            blk["roofline"]["note"] = "encode of the first batch (%.2f GB)" % (f1 / 1e9)
        extra["bgi_se100"] = blk
        enc.set_header(header)
        dec.set_header(header)

    # ---- e2e: pinned host buffers through the C ABI, H2D and D2H inside the timed region.
    # Two measurements of the same K steps (encode host->host, then decode host->host):
    #   serial    : step i+1 starts when step i is done; one PCIe direction is busy at a time
    #   pipelined : the encode of step i+1 (host->device heavy) runs while step i is decoded (device->host heavy): two host
    #               threads, one context each (the ABI's rule is one caller thread per context); encode results alternate
    #               between two contexts so that the decoder reads a buffer nobody is overwriting.  Every step's H2D and
    #               D2H is inside the timed region; all K steps have completed when the clock stops.
    e2e = None
    if not args.no_e2e:
        import queue
        del d1, d2
        torch.cuda.empty_cache()
        link = link_probe(torch, dist, world, local)
        h1 = torch.from_numpy(r1).pin_memory()
        h2 = torch.from_numpy(r2).pin_memory()
        enc_b = K.Codec(device=local)
        enc_b.set_header(header)
        encs = [enc, enc_b]

        def encode_host(cd):
            return cd.encode_raw(h1.data_ptr(), h1.numel(), h2.data_ptr(), h2.numel(), 0, False, chunk_bases, True, (K.NEVER, K.NEVER), 0, 0)

        def check_host(do):
            a1 = np.ctypeslib.as_array(C.cast(do.out1, C.POINTER(C.c_uint8)), shape=(do.out1_bytes,))
            a2 = np.ctypeslib.as_array(C.cast(do.out2, C.POINTER(C.c_uint8)), shape=(do.out2_bytes,))
            assert do.out1_bytes == r1.size and np.array_equal(a1[:1 << 20], r1[:1 << 20]) and np.array_equal(a1[-(1 << 20):], r1[-(1 << 20):])
            assert do.out2_bytes == r2.size and np.array_equal(a2[:1 << 20], r2[:1 << 20]) and np.array_equal(a2[-(1 << 20):], r2[-(1 << 20):])

        def run_serial(n):
            acc = dict(h2d=0, d2h=0, dec_h2d_ms=0.0, dec_kernels_ms=0.0, dec_d2h_ms=0.0)
            do = None
            for _ in range(n):
                eo = encode_host(enc)
                se = enc.stats()
                do = dec.decode_raw(eo.data, eo.bytes, 0, True, 0)
                sd = dec.stats()
                acc["h2d"] += se.h2d_bytes + sd.h2d_bytes
                acc["d2h"] += se.d2h_bytes + sd.d2h_bytes
                acc["dec_h2d_ms"] += sd.ms_h2d / n
                acc["dec_kernels_ms"] += sd.ms_kernels / n
                acc["dec_d2h_ms"] += sd.ms_d2h / n
            return acc, do

        def run_pipelined(n):
            acc = dict(h2d=0, d2h=0)
            ready = queue.Queue()
            free = [threading.Semaphore(1), threading.Semaphore(1)]
            err = []

            def producer():
                try:
                    torch.cuda.set_device(local)
                    for i in range(n):
                        free[i & 1].acquire()                    # the decoder is done with this context's previous result
                        t_c = time.perf_counter()
                        eo = encode_host(encs[i & 1])
                        acc["enc_call_ms"] = acc.get("enc_call_ms", 0.0) + 1e3 * (time.perf_counter() - t_c) / n
                        se = encs[i & 1].stats()
                        ready.put((i, eo, se.h2d_bytes, se.d2h_bytes))
                except Exception as ex:                          # noqa: BLE001
                    err.append(ex)
                    ready.put(None)
            th = threading.Thread(target=producer)
            th.start()
            do = None
            for _ in range(n):
                item = ready.get()
                if item is None:
                    break
                i, eo, hb_, db = item
                t_c = time.perf_counter()
                do = dec.decode_raw(eo.data, eo.bytes, 0, True, 0)
                acc["dec_call_ms"] = acc.get("dec_call_ms", 0.0) + 1e3 * (time.perf_counter() - t_c) / n
                sd = dec.stats()
                free[i & 1].release()
                acc["h2d"] += hb_ + sd.h2d_bytes
                acc["d2h"] += db + sd.d2h_bytes
            th.join()
            if err:
                raise err[0]
            return acc, do

        def timed(fn):
            fn(W)
            barrier()
            t_e = time.perf_counter()
            acc, do = fn(args.steps)
            barrier()
            ms = 1e3 * (time.perf_counter() - t_e)
            check_host(do)
            te = torch.tensor([ms], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            return float(te[0]), acc
        s_ms, s_acc = timed(run_serial)
        p_ms, p_acc = timed(run_pipelined)
        gbs = lambda ms: job_bytes * args.steps / 1e9 / (ms / 1e3)      # noqa: E731
        h2d_step, d2h_step = int(p_acc["h2d"] // args.steps), int(p_acc["d2h"] // args.steps)
        # what the link allows: this rank's copies at the duplex rates measured with all ranks copying at once
        floor_ms = 1e3 * max(h2d_step / 1e9 / link[0], d2h_step / 1e9 / link[1])
        fl = torch.tensor([floor_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(fl, op=dist.ReduceOp.MAX)
        ceiling = job_bytes / 1e9 / (float(fl[0]) / 1e3)
        e2e = dict(value=gbs(p_ms), unit=UNIT, h2d_bytes_per_step=h2d_step, d2h_bytes_per_step=d2h_step,
                   ms_per_step=p_ms / args.steps, encode_call_ms=round(p_acc.get("enc_call_ms", 0.0), 2), decode_call_ms=round(p_acc.get("dec_call_ms", 0.0), 2), mode="pipelined: encode of step i+1 overlaps decode of step i (two host threads, one context each; H2D and D2H share the link full duplex)",
                   link_ceiling_gbs=ceiling, frac_of_ceiling=gbs(p_ms) / ceiling,
                   link=dict(h2d_gbs=link[0], d2h_gbs=link[1], how="1 GiB pinned copies in both directions at once on every rank at the same time (rank 0's rates; the ceiling is the slowest rank's)"),
                   serial=dict(value=gbs(s_ms), ms_per_step=s_ms / args.steps, h2d_bytes_per_step=int(s_acc["h2d"] // args.steps), d2h_bytes_per_step=int(s_acc["d2h"] // args.steps),
                               decode_breakdown_ms_per_step={k: round(v, 3) for k, v in s_acc.items() if k.endswith("_ms")}),
                   timing="host wall clock around the C-ABI calls (they return after their D2H completed), K steps, max over ranks")
        enc_b.close()
        del h1, h2

    # ---- CPU baseline beside it (rank 0, N=1 only): bounded sample of the same workload, one reference process per host core;
    # what the reference binary wrote for every shard is compared with what the GPU path makes of the same shard, both directions
    cpu = None
    parity = None
    if old_affinity is not None:
        os.sched_setaffinity(0, old_affinity)              # the CPU baseline gets every core of the box
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = host_cores()
        n_pairs = min(hi - lo, 33340 * cores)
        s1, s2 = fqgen.truncate_reads(r1, n_pairs), fqgen.truncate_reads(r2, n_pairs)
        acc = dict(enc=0, dec=0, shards=0)

        def check(i, a1, a2, ref_rfq, ref_d1, ref_d2):
            got = K.compress(a1, a2, k=1000, codec=enc)
            assert np.array_equal(np.frombuffer(got, dtype=np.uint8), ref_rfq), "shard %d: the GPU .rfq differs from the reference binary's" % i
            g1, g2 = K.decompress(ref_rfq.tobytes(), pe_out=True, codec=dec)
            assert np.array_equal(np.frombuffer(g1, dtype=np.uint8), ref_d1) and np.array_equal(np.frombuffer(g2, dtype=np.uint8), ref_d2), "shard %d: the GPU decode differs from the reference binary's" % i
            acc["enc"] += int(a1.size + a2.size)
            acc["dec"] += int(ref_d1.size + ref_d2.size)
            acc["shards"] += 1
        cpu = cpu_reference_roundtrip(s1, s2, cores, check=check)
        parity = dict(encode_bytes=acc["enc"], decode_bytes=acc["dec"], shards=acc["shards"], against="oracle/_ref/repaq (the unmodified reference binary)" if cpu["kind"] == "reference" else "oracle/ (C port)",
                      how="every shard of the cpu_baseline sample: GPU .rfq == the reference's .rfq file byte for byte, GPU decode of the reference's .rfq == the reference's decoded FASTQ")

    # ---- the command-line drivers against each other (rank 0, N=1): repaq_b200_cli on the whole 3.4 GB pair, the unmodified reference
    # binary on a bounded sample of it, files in /dev/shm, wall clock of the whole process (start-up, file I/O, copies included)
    cli = None
    cli_bin = os.path.join(ROOT, "repaq_b200", "repaq_b200_cli")
    if rank == 0 and world == 1 and not args.no_cpu and os.path.exists(cli_bin) and os.path.isdir("/dev/shm"):
        tmp = tempfile.mkdtemp(dir="/dev/shm")
        try:
            r1.tofile(os.path.join(tmp, "a.fq"))
            r2.tofile(os.path.join(tmp, "b.fq"))

            def wall(cmd):
                t = time.perf_counter()
                subprocess.check_call(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                return time.perf_counter() - t
            t_c = wall([cli_bin, "-c", "-i", f"{tmp}/a.fq", "-I", f"{tmp}/b.fq", "-o", f"{tmp}/o.rfq", "--device=%d" % local])
            t_d = wall([cli_bin, "-d", "-i", f"{tmp}/o.rfq", "-o", f"{tmp}/d1.fq", "-O", f"{tmp}/d2.fq", "--device=%d" % local])
            same = all(np.array_equal(np.fromfile(os.path.join(tmp, a), dtype=np.uint8), b) for a, b in (("d1.fq", r1), ("d2.fq", r2)))
            assert same, "repaq_b200_cli: decoded files differ from the inputs"
            rfq_file = np.fromfile(os.path.join(tmp, "o.rfq"), dtype=np.uint8)
            assert rfq_file.size == len(hb) + rfq_b and bytes(rfq_file[:len(hb)]) == hb
            cli = dict(files="%.2f GB pair in /dev/shm" % (fastq_bytes / 1e9), gpu_cli=dict(encode_s=t_c, decode_s=t_d, encode_gbs=fastq_bytes / 1e9 / t_c, decode_gbs=fastq_bytes / 1e9 / t_d,
                                                                                          roundtrip_gbs=fastq_bytes / 1e9 / (t_c + t_d)),
                       verified="decoded files byte-identical to the inputs; .rfq size and header equal to the library's")
            for n in ("o.rfq", "d1.fq", "d2.fq"):
                os.unlink(os.path.join(tmp, n))
            if O.have_ref():
                n_pairs = min(hi - lo, 600000)                # ~0.43 GB: ~4 s per direction for the single-threaded reference
                s1, s2 = fqgen.truncate_reads(r1, n_pairs), fqgen.truncate_reads(r2, n_pairs)
                s1.tofile(os.path.join(tmp, "sa.fq"))
                s2.tofile(os.path.join(tmp, "sb.fq"))
                sb = int(s1.size + s2.size)
                r_c = wall([O.REF_BIN, "-c", "-i", f"{tmp}/sa.fq", "-I", f"{tmp}/sb.fq", "-o", f"{tmp}/s.rfq"])
                r_d = wall([O.REF_BIN, "-d", "-i", f"{tmp}/s.rfq", "-o", f"{tmp}/s1.fq", "-O", f"{tmp}/s2.fq"])
                g_c = wall([cli_bin, "-c", "-i", f"{tmp}/sa.fq", "-I", f"{tmp}/sb.fq", "-o", f"{tmp}/g.rfq", "--device=%d" % local])
                assert open(f"{tmp}/s.rfq", "rb").read() == open(f"{tmp}/g.rfq", "rb").read(), "repaq_b200_cli and the reference binary wrote different .rfq files"
                cli["reference_cli"] = dict(sample="%.2f GB pair (the first %d pairs of the same files), one process, one thread" % (sb / 1e9, n_pairs), encode_s=r_c, decode_s=r_d,
                                            encode_gbs=sb / 1e9 / r_c, decode_gbs=sb / 1e9 / r_d, roundtrip_gbs=sb / 1e9 / (r_c + r_d))
                cli["same_sample_gpu_cli_encode_s"] = g_c
                cli["verified"] += "; on the sample both drivers wrote the same .rfq file byte for byte"
                cli["roundtrip_ratio"] = cli["gpu_cli"]["roundtrip_gbs"] / cli["reference_cli"]["roundtrip_gbs"]
        finally:
            subprocess.call(["rm", "-rf", tmp])

    if rank == 0:
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=W, ms_per_step=dev_ms / args.steps,
                    higher_is_better=True, scaling="weak", vs_baseline=None, dtype="u8", data="synthetic",
                    config=dict(workload="configs[1]+[2]: paired-end NovaSeq-shape 150bp, %.2f GB FASTQ per GPU (R1+R2), encode to .rfq then decode back" % (fastq_bytes / 1e9),
                                pairs_per_gpu=hi - lo, chunk_kb=1000, rfq_bytes_per_gpu=rfq_b, rfq_ratio=rfq_b / fastq_bytes,
                                l2="inputs (GBs) far larger than the 126 MB L2; no flush needed", generator="tools/fqgen.c seed 2", gen_seconds=round(gen_s, 1),
                                parallelism="ONE file of %d pairs, contiguous chunk ranges per GPU (one process each), header of chunk 0 broadcast, one NCCL all_gather of per-chunk lengths per step (side stream)" % n_pairs_total,
                                host_affinity="CPUs local to the rank's GPU (NVML)" if old_affinity is not None else "unchanged"),
                    encode_gbs=job_bytes * args.steps / 1e9 / (enc_ms / 1e3), decode_gbs=job_bytes * args.steps / 1e9 / (dec_ms / 1e3),
                    wall_ms_per_step=wall_ms / args.steps, gpu_launches=int(launches),
                    sharding=dict(file_pairs=n_pairs_total, file_chunks=sum(chunks_per_rank), body_bytes=body_total, chunks_per_rank=chunks_per_rank,
                                  verified="every rank: first and last 6 chunks bit-exact against the oracle run with the broadcast header; decode == input; offsets from the gathered lengths" + ("; " + seam_note if seam_note else "")),
                    parity_checked_bytes=dict(oracle_per_rank=oracle_checked, reference_binary=parity),
                    decode_chunk_walk={1: "one warp on the mSize chain", 2: "16 warps on the mSize chain (k_dec_walk_par)", 3: "exact sequential walk"}.get(state.get("dec_walk"), "host"), clocks=clk, e2e=e2e, roofline=roofline, cpu_baseline=cpu,
                    configs=extra or None, cli_e2e=cli)
        print(json.dumps(line))
    xch.close()
    enc.close()
    dec.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
