#!/usr/bin/env python
"""bench.py - FASTQ <-> .rfq throughput of the B200 path on BASELINE.json's headline workload.

One step = encode a paired-end NovaSeq-shape FASTQ batch to .rfq AND decode it back (the .rfq is checked
bit-exact against the oracle on a sample, the decode byte-exact against the input).
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


def cpu_reference_roundtrip(r1, r2, cores, budget_s=None):
    """Times the reference CPU implementation (compress + decompress) on `cores` disjoint record-aligned shards run
    concurrently, files in /dev/shm.  Returns dict(value GB/s of the round trip, kind, cores, sample, seconds)."""
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
        else:
            res = [None] * cores

            def work(i, phase):
                if phase == 0:
                    res[i] = O.compress(shards1[i], shards2[i])
                else:
                    O.decompress(res[i], pe_out=r2 is not None)
            for phase in (0, 1):
                t = time.perf_counter()
                th = [threading.Thread(target=work, args=(i, phase)) for i in range(cores)]
                [x.start() for x in th]
                [x.join() for x in th]
                if phase == 0:
                    t_enc = time.perf_counter() - t
                else:
                    t_dec = time.perf_counter() - t
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
    from tools import fqgen

    old_affinity = bind_near_gpu(local)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W = max(3, args.warmup)

    # ---- synthetic workload: every rank its own rows of the same generator (weak scaling)
    rows = (args.pairs + fqgen.ROW_READS - 1) // fqgen.ROW_READS
    t0 = time.perf_counter()
    r1, r2 = fqgen.generate(rows * fqgen.ROW_READS, seed=2, paired=True, first_row=rank * rows,
                            threads=max(1, host_cores() // max(1, world)))
    gen_s = time.perf_counter() - t0
    fastq_bytes = int(r1.size + r2.size)
    header = K.make_header(r1, r2)
    enc, dec = K.Codec(device=local), K.Codec(device=local)
    enc.set_header(header)
    dec.set_header(header)
    chunk_bases = 1000000

    d1 = torch.from_numpy(r1).cuda()
    d2 = torch.from_numpy(r2).cuda()
    torch.cuda.synchronize()
    es = torch.cuda.ExternalStream(enc.L.rpq_stream(enc.ctx))
    ds = torch.cuda.ExternalStream(dec.L.rpq_stream(dec.ctx))

    state = {}

    def step_device():
        eo = enc.encode_raw(d1.data_ptr(), d1.numel(), d2.data_ptr(), d2.numel(), 1, False, chunk_bases, True, (K.NEVER, K.NEVER), 0, 1)
        se = enc.stats()
        do = dec.decode_raw(eo.data, eo.bytes, 1, True, 1)
        sd = dec.stats()
        state.update(rfq_bytes=int(eo.bytes), n_chunks=int(eo.n_chunks), eo=eo, do=do, out=(int(do.out1_bytes), int(do.out2_bytes)), dec_walk=int(sd.dec_walk))
        return se.ms_total, sd.ms_total, se.launches + sd.launches, eo

    def gather_lengths(eo):
        """the one real exchange of the multi-GPU path: per-chunk serialised lengths -> file offsets of every rank"""
        if world == 1:
            return
        lens = torch.tensor([eo.chunks[i].bytes for i in range(eo.n_chunks)], dtype=torch.int64, device="cuda")
        n = torch.tensor([lens.numel()], dtype=torch.int64, device="cuda")
        ns = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(ns, n)
        mx = int(max(int(x) for x in ns))
        pad = torch.zeros(mx, dtype=torch.int64, device="cuda")
        pad[: lens.numel()] = lens
        allp = [torch.zeros_like(pad) for _ in range(world)]
        dist.all_gather(allp, pad)
        state["file_offset"] = int(sum(int(allp[r][: int(ns[r])].sum()) for r in range(rank)))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- correctness gate (not timed): bit-exact .rfq on the first chunks, byte-exact round trip on everything
    _, _, _, eo = step_device()
    rfq_dev = torch.empty(state["rfq_bytes"], dtype=torch.uint8, device="cuda")
    cudart = C.cdll.LoadLibrary("libcudart.so.12")
    cudart.cudaMemcpy(C.c_void_p(rfq_dev.data_ptr()), C.c_void_p(eo.data), C.c_size_t(eo.bytes), 3)
    n_chk = min(12, state["n_chunks"])
    p1, p2 = fqgen.truncate_reads(r1, 3334 * n_chk), fqgen.truncate_reads(r2, 3334 * n_chk)
    ref = O.compress(bytes(p1), bytes(p2), chunk_bases=chunk_bases)
    hb = K.header_bytes(header)
    got = bytes(rfq_dev[: len(ref) - len(hb)].cpu().numpy())
    assert hb + got == ref, "encode is not bit-exact against the oracle"
    o1 = torch.empty(state["out"][0], dtype=torch.uint8, device="cuda")
    o2 = torch.empty(state["out"][1], dtype=torch.uint8, device="cuda")
    cudart.cudaMemcpy(C.c_void_p(o1.data_ptr()), C.c_void_p(state["do"].out1), C.c_size_t(state["out"][0]), 3)
    cudart.cudaMemcpy(C.c_void_p(o2.data_ptr()), C.c_void_p(state["do"].out2), C.c_size_t(state["out"][1]), 3)
    assert torch.equal(o1, d1) and torch.equal(o2, d2), "decode does not restore the input"
    del o1, o2, rfq_dev

    # ---- timed: inputs resident in HBM (inputs are ~3.4 GB per step: far larger than the 126 MB L2, no flush needed)
    try:
        uuid = str(torch.cuda.get_device_properties(local).uuid)
    except Exception:
        uuid = None
    clocks = ClockSampler(local, uuid)
    if rank == 0:
        clocks.start()
    for _ in range(W):
        _, _, _, eo = step_device()
        gather_lengths(eo)
    barrier()
    clocks.begin()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall = time.perf_counter()
    ev0.record(es)
    enc_ms = dec_ms = 0.0
    launches = 0
    for _ in range(args.steps):
        a, b_, l, eo = step_device()
        gather_lengths(eo)
        enc_ms += a
        dec_ms += b_
        launches += l
    ev1.record(ds)
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t_wall)
    clocks.end()
    dev_ms = ev0.elapsed_time(ev1)                  # CUDA events: first encode op .. last decode op, host gaps included
    clk = clocks.stop() if rank == 0 else None
    t = torch.tensor([dev_ms, wall_ms, enc_ms, dec_ms], dtype=torch.float64, device="cuda")
    tot_bytes = torch.tensor([fastq_bytes], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot_bytes, op=dist.ReduceOp.SUM)
    dev_ms, wall_ms, enc_ms, dec_ms = [float(x) for x in t]
    job_bytes = float(tot_bytes[0])
    value = job_bytes * args.steps / 1e9 / (dev_ms / 1e3)

    # ---- per-kernel device time (separate pass, event pair around every launch) -> roofline of the dominant kernel
    prof = {}
    if not args.no_roofline:
        enc.set_profiling(True)
        dec.set_profiling(True)
        step_device()
        for k, (n, ms) in list(enc.profile().items()) + list(dec.profile().items()):
            prof[k] = (n, ms)
        enc.set_profiling(False)
        dec.set_profiling(False)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    n_reads = 2 * rows * fqgen.ROW_READS
    rfq_b = state["rfq_bytes"]
    qual_b = seq_b = n_reads * 150
    # algorithmic bytes per launch of each kernel (DESIGN.md "Kernels"): what it must read + write once
    head_b = fastq_bytes - qual_b - n_reads                                     # name + sequence + strand lines
    alg = {
        "k_index_lines": fastq_bytes + 4 * 4 * n_reads,                          # text in, line index out
        "k_unit_lengths": (16 + 16 + 4 + 4) * n_reads,                           # line index in, record index + lengths out
        "k_meta": head_b + 16 * n_reads, "k_meta2": head_b + (16 + 16 + 3 * 44 / 2) * n_reads,   # heads in, metadata + packed reads out
        "k_streams": qual_b + 0.6 * rfq_b, "k_streams2": qual_b + 0.6 * rfq_b,   # qualities in, tokens out
        "k_emit": seq_b + seq_b / 4, "k_emit2": 44 * n_reads + seq_b / 4,        # packed reads in, 2-bit stream out
        "k_gather": 1.2 * rfq_b,
        "k_dec_format": rfq_b + qual_b + fastq_bytes, "k_dec_format2": rfq_b + qual_b + fastq_bytes,   # columns + plane in, text out
        "k_dec_streams": 0.6 * rfq_b + 0.1 * qual_b,
        # current generation
        "k_meta3": head_b + (16 + 44) * n_reads,                                 # record heads in, ReadMeta + packed read out
        "k_streams3": qual_b + 0.6 * rfq_b, "k_streams4": qual_b + 0.6 * rfq_b,  # qualities in, tokens out
        "k_dec_format3": rfq_b + qual_b + fastq_bytes,                           # columns + quality plane in, text out
        "k_dec_format4": rfq_b + fastq_bytes,                                    # columns (streams included) in, text out
        "k_dec_qindex": 0.3 * rfq_b,                                             # position streams in, checkpoints out
        "k_dec_coords3": (1 + 8) * n_reads, "k_dec_reads": 48 * n_reads, "k_chunk_finish": 44 * n_reads,
        "k_coords": 9 * n_reads,
    }
    # DRAM bytes per launch from the committed ncu launch list of this workload (profiles/r01_traffic.json, tools/ncu_traffic.py)
    traffic = {}
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        if int(tj.get("pairs_per_gpu", 0)) == rows * fqgen.ROW_READS:
            traffic = {k: v["traffic_bytes_per_launch"] for k, v in tj["kernels"].items()}
    except Exception:
        pass
    kern_total = sum(ms for _, ms in prof.values())
    top = max(prof.items(), key=lambda kv: kv[1][1]) if prof else None
    roofline = None
    if top:
        name, (n, ms) = top
        ab = float(alg.get(name, fastq_bytes))
        ach = ab / 1e9 / (ms / n / 1e3)
        roofline = dict(bound="hbm", kernel=name, achieved=ach, peak=peak, unit="GB/s", frac=ach / peak, traffic=traffic.get(name), peak_source=peak_src,
                        traffic_source="profiles/r01_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu capture of this workload" if name in traffic else None,
                        share_of_kernel_time=ms / kern_total, algorithmic_bytes_per_launch=ab,
                        pipeline=dict(achieved=(fastq_bytes + rfq_b) * 2 / 1e9 / (kern_total / 1e3), frac=(fastq_bytes + rfq_b) * 2 / 1e9 / (kern_total / 1e3) / peak,
                                      note="whole encode+decode: (F+R)+(R+F) algorithmic bytes over the sum of all kernel times"),
                        kernels_ms={k: round(v[1], 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])})

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
                i, eo, hb, db = item
                t_c = time.perf_counter()
                do = dec.decode_raw(eo.data, eo.bytes, 0, True, 0)
                acc["dec_call_ms"] = acc.get("dec_call_ms", 0.0) + 1e3 * (time.perf_counter() - t_c) / n
                sd = dec.stats()
                free[i & 1].release()
                acc["h2d"] += hb + sd.h2d_bytes
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
        e2e = dict(value=gbs(p_ms), unit=UNIT, h2d_bytes_per_step=int(p_acc["h2d"] // args.steps), d2h_bytes_per_step=int(p_acc["d2h"] // args.steps),
                   ms_per_step=p_ms / args.steps, encode_call_ms=round(p_acc.get("enc_call_ms", 0.0), 2), decode_call_ms=round(p_acc.get("dec_call_ms", 0.0), 2), mode="pipelined: encode of step i+1 overlaps decode of step i (two host threads, one context each; H2D and D2H share the link full duplex)",
                   serial=dict(value=gbs(s_ms), ms_per_step=s_ms / args.steps, h2d_bytes_per_step=int(s_acc["h2d"] // args.steps), d2h_bytes_per_step=int(s_acc["d2h"] // args.steps),
                               decode_breakdown_ms_per_step={k: round(v, 3) for k, v in s_acc.items() if k.endswith("_ms")}),
                   timing="host wall clock around the C-ABI calls (they return after their D2H completed), K steps, max over ranks")
        enc_b.close()
        del h1, h2

    # ---- CPU baseline beside it (rank 0, N=1 only): bounded sample of the same workload
    cpu = None
    if old_affinity is not None:
        os.sched_setaffinity(0, old_affinity)              # the CPU baseline gets every core of the box
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = host_cores()
        n_pairs = min(args.pairs, 33340 * cores)
        s1, s2 = fqgen.truncate_reads(r1, n_pairs), fqgen.truncate_reads(r2, n_pairs)
        cpu = cpu_reference_roundtrip(s1, s2, cores)

    if rank == 0:
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=W, ms_per_step=dev_ms / args.steps,
                    higher_is_better=True, scaling="weak", vs_baseline=None, dtype="u8", data="synthetic",
                    config=dict(workload="configs[1]+[2]: paired-end NovaSeq-shape 150bp, %.2f GB FASTQ per GPU (R1+R2), encode to .rfq then decode back" % (fastq_bytes / 1e9),
                                pairs_per_gpu=rows * fqgen.ROW_READS, chunk_kb=1000, rfq_bytes_per_gpu=rfq_b, rfq_ratio=rfq_b / fastq_bytes,
                                l2="inputs (GBs) far larger than the 126 MB L2; no flush needed", generator="tools/fqgen.c seed 2", gen_seconds=round(gen_s, 1),
                                parallelism="chunk-sharded, one process per GPU; NCCL all_gather of per-chunk lengths only",
                                host_affinity="CPUs local to the rank's GPU (NVML)" if old_affinity is not None else "unchanged"),
                    encode_gbs=job_bytes * args.steps / 1e9 / (enc_ms / 1e3), decode_gbs=job_bytes * args.steps / 1e9 / (dec_ms / 1e3),
                    wall_ms_per_step=wall_ms / args.steps, gpu_launches=int(launches),
                    decode_chunk_walk={1: "one warp on the mSize chain", 2: "16 warps on the mSize chain (k_dec_walk_par)", 3: "exact sequential walk"}.get(state.get("dec_walk"), "host"), clocks=clk, e2e=e2e, roofline=roofline, cpu_baseline=cpu)
        print(json.dumps(line))
    enc.close()
    dec.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
