#!/usr/bin/env python
"""Benchmark of the lossycomp hot path on B200 (contract: see the task's bench.py section).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision P]

N = 1: a step = one compress() of one synthetic ERA5-shaped field (BASELINE.json configs[1]: 24 h x 1 level x
720 x 1440 float32 + land-sea channel, abs_error 0.1).
N > 1 (torchrun): a step = compress() of ONE field of 24*N hours, time-sharded over the N ranks in whole
16-step chunks (lossycomp.dist.compress_sharded): NCCL all-reduce of the per-block statistics and
of strict mode's verdict, NCCL send/recv of the blob parts from the Huffman output buffers to one rank and one
device->host copy there -- all inside the timed region.  After the timed region the sharded result is checked
bit for bit (mask, codes, reconstruction) against a single-GPU call on the same field.
value = bytes of channel-0 input all ranks compressed / max over ranks of the device time.

Keys beyond the base contract: `roofline` (dominant kernel = the 3x3x3 tensor-core convolution, against the
measured bf16 peak; also the conv stack, decompress and per-kernel numbers), `cpu_baseline` (the oracle timed on
the host cores over a bounded sample, N = 1 only), `e2e` (public API with host buffers; also carries compression
factor, violations, parity verdicts and the decompress numbers so that they survive the driver's key filter).
"""
import argparse
import json
import os
import pickle
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "lossy-ml_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "compress_throughput_fp32_input"
UNIT = "GB/s"
ABS_ERR = 0.1
FIELD = (24, 720, 1440)          # configs[1] of BASELINE.json
HOURS_PER_RANK = 24              # N > 1: one field of 24*N hours
CPU_SAMPLE = (24, 336, 720)      # bounded sample of the same field for the CPU leg inside the GPU arm (~4 s)
MODEL_NAME = "final_models/model_2 (F=23,k=4,res=1)"
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "ncu_traffic.json")   # ncu --set full DRAM bytes per launch, by kernel


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("LOSSYCOMP_PRECISION", "fp16"))
    ap.add_argument("--codec", default="huff")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--field", default=None, help="T,H,W override of the per-rank field (tests)")
    ap.add_argument("--hours-per-rank", type=int, default=HOURS_PER_RANK,
                    help="N > 1: the field has N times this many hours (32 = two whole time chunks per rank)")
    ap.add_argument("--dst", type=int, default=None, help="N > 1: rank that receives the blob (default: the last)")
    ap.add_argument("--skip-parity", action="store_true", help="N > 1: skip the single-GPU comparison")
    ap.add_argument("--lanes", type=int, default=None,
                    help="fields in flight per GPU (lossycomp.lanes; default LOSSYCOMP_LANES or 2; 1 = one after the other)")
    return ap.parse_args()


def workload_config(world, T, H, W, n_chunks, codec, precision, in_bytes, dst=None):
    """The `config` object; the reference arm prints the same one (its CPU sample is described in cpu_baseline)."""
    if world == 1:
        par = "1 rank"
        wl = (f"compress() of one synthetic ERA5 slice {T}x{H}x{W} fp32 + lsm channel, abs_err {ABS_ERR}, "
              f"strict bound, {n_chunks} chunks of 16x48x48, codec {codec}")
    else:
        par = (f"one field time-sharded over {world} ranks in whole 16-step chunks (dist.compress_sharded): NCCL "
               f"all-reduce of per-block statistics and of the strict-mode verdict, NCCL send/recv of the blob parts "
               f"from the Huffman output buffers to rank {dst}, one device->host copy there (copy stream + host thread, "
               f"overlapping the next step's kernels); all K blobs are complete before the timed region ends")
        wl = (f"compress() of ONE synthetic ERA5 field {T}x{H}x{W} fp32 + lsm channel ({T // world} h per rank), "
              f"abs_err {ABS_ERR}, strict bound, {n_chunks} chunks of 16x48x48, codec {codec}")
    return {"workload": wl, "model": MODEL_NAME, "precision": precision,
            "l2": f"inputs {in_bytes * 2 / 1e6:.0f} MB per rank and step exceed the 126 MB L2",
            "parallelism": par}


# ------------------------------------------------------------------------------------------
def cpu_leg(steps, warmup, sample, threads=None):
    """The reference's CPU path (oracle port: NumPy + torch-CPU fp32 convs, batch 10, bz2 -9)."""
    import torch
    if threads:
        torch.set_num_threads(threads)           # torchrun exports OMP_NUM_THREADS=1: do not inherit it
    from lossycomp.models import load_model
    from oracle import lossycomp_oracle as O
    model = load_model()
    T, H, W = sample
    x = O.synthetic_field(T, H, W, seed=1234)
    nbytes = x[..., 0].nbytes
    for _ in range(max(0, min(warmup, 1))):
        O.compress(x[:16, :96, :96], ABS_ERR, model)
    times, dtimes = [], []
    blob = parts = None
    for _ in range(steps):
        t0 = time.perf_counter()
        blob, parts = O.compress(x, ABS_ERR, model, return_parts=True)
        times.append(time.perf_counter() - t0)
    t0 = time.perf_counter()
    xhat = O.decompress(blob, model)
    dtimes.append(time.perf_counter() - t0)
    viol = int((np.abs(x[..., 0].astype(np.float64) - xhat[..., 0]) > ABS_ERR).sum())
    return {
        "value": nbytes / np.mean(times) / 1e9, "unit": UNIT, "cores": torch.get_num_threads(),
        "kind": "port",
        "sample": f"oracle compress() of a {T}x{H}x{W}x2 window of the same synthetic field "
                  f"({nbytes / 1e6:.1f} MB), abs_err {ABS_ERR}, torch-CPU fp32 convs standing in for TF-CPU",
        "seconds_per_step": float(np.mean(times)), "decompress_GBps": nbytes / dtimes[0] / 1e9,
        "compression_factor": nbytes / len(blob), "violations_reference_arithmetic": viol,
        "host_cpus": os.cpu_count(),
    }, x, parts


def run_reference(args):
    """The reference's own CPU implementation of the path (the oracle port: TensorFlow cannot be installed here),
    all host threads, one full 24x720x1440 slice per step.  N > 1: rank 0 alone runs it, on the same per-step
    slice (the field of 24*N hours is N such slices; GB/s is size-normalised), the other ranks exit."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = args.gpus
    steps = max(1, min(args.steps, 2))
    t0 = time.perf_counter()
    T, H, W = FIELD if not args.field else tuple(int(v) for v in args.field.split(","))
    cb, _, _ = cpu_leg(steps, args.warmup, (T, H, W), threads=os.cpu_count())
    Tw = T if world == 1 else args.hours_per_rank * world
    n_chunks = -(-Tw // 16) * -(-H // 48) * -(-W // 48)
    if world > 1:
        cb["sample"] += (f"; the GPU arm's field at {world} ranks is {Tw}x{H}x{W} = {world} such slices, the CPU arm "
                         f"times one of them per step on all {cb['cores']} host threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": world,
        "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": cb["seconds_per_step"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(world, Tw, H, W, n_chunks, args.codec, args.precision, T * H * W * 4,
                                  dst=(world - 1 if args.dst is None else args.dst)),
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "compression_factor": cb["compression_factor"], "decompress_GBps": cb["decompress_GBps"]},
        "compression_factor": cb["compression_factor"], "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in ln.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 and len(r) >= 9] or [r for (_, r) in self.rows if len(r) >= 9]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in rows)
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "reasons": sorted(reasons),
                "samples": len(rows), "power_w_max": max(float(r[3]) for r in rows)}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return p, "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, \
            "fallback (B200_PROFILING.md)"


def measured_traffic(kernel, n_chunks):
    """DRAM bytes per launch of `kernel` from the committed `ncu --set full` capture of THIS workload
    (profiles/ncu_traffic.json, written by tools/ncu_traffic.py from the .ncu-rep); None when the capture is of
    another chunk count."""
    try:
        with open(TRAFFIC_FILE) as fh:
            t = json.load(fh)
        e = t["kernels"][kernel]
        if int(t["n_chunks"]) != int(n_chunks):
            return None, None
        return float(e["dram_bytes_per_launch"]), t.get("source")
    except Exception:
        return None, None


def run_ours(args):
    import torch
    import torch.distributed as dist
    from lossycomp import _native as N
    from lossycomp import dist as LD
    from lossycomp._engine import Engine
    from lossycomp.compress import compress, compress_device, compress_stream
    from lossycomp.decompress import decompress, decompress_device, decompress_stream

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N.lib()
    eng = Engine.get(None, args.precision, local)
    Tr, H, W = FIELD if not args.field else tuple(int(v) for v in args.field.split(","))
    if world > 1:
        Tr = args.hours_per_rank if not args.field else Tr
    T = Tr * world                                   # rows of the whole field
    dst = (world - 1) if args.dst is None else args.dst
    shards = LD.shard_time(T, world) if world > 1 else [(0, T)]
    t0r, t1r = shards[rank]
    # this rank's rows of the synthetic field (window [t0r, t1r) of the global time axis)
    x = torch.empty((t1r - t0r, H, W, 2), dtype=torch.float32, device=eng.dev)
    if t1r > t0r:
        N.check(eng.lib.lc_synth_field(x.data_ptr(), t1r - t0r, H, W, t0r, 0, 1234, None), "lc_synth_field")
    torch.cuda.synchronize()
    in_bytes = T * H * W * 4                          # whole job, channel 0
    my_bytes = (t1r - t0r) * H * W * 4
    n_chunks_all = eng.n_chunks(T, H, W)
    n_chunks = eng.n_chunks(t1r - t0r, H, W) if t1r - t0r >= 16 else 0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        tms = torch.tensor([v], dtype=torch.float64, device=eng.dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        return float(tms.item())

    # Fields in flight.  N = 1: consecutive steps run in alternating lanes (engine + stream + host thread each), so the
    # kernels of one field fill the host round trips of the other.  N > 1: the same overlap as a software pipeline in
    # THIS thread -- field k+1 is begun (statistics all-reduce, encoder, decoder, quantiser enqueued on the other
    # engine) before field k is finished (verdict, Huffman tables, NCCL gather) -- so that every collective is issued
    # by one thread in the same order on every rank.  Lane 0 is `eng`.
    from lossycomp.lanes import LaneSet, n_lanes
    lanes = LaneSet.get(None, args.precision, local, n_lanes(args.lanes))

    def job(lane):
        return compress_device(lane.eng, x, ABS_ERR, mode="strict", codec=args.codec)

    def finish(fut):
        out, info = fut.result()
        return (out.result() if hasattr(out, "result") else out), info

    def sharded_begin(lane):
        with torch.cuda.stream(lane.stream):
            return LD.compress_sharded_begin(lane.eng, x, ABS_ERR, t0=t0r, T=T, mode="strict", codec=args.codec)

    def sharded_finish(lane, st):
        # rank dst gets a PendingBlob: the gathered blob's device->host copy (copy stream) and its assembly (host
        # thread) overlap the following kernels; the timed region ends when every step's blob is complete
        with torch.cuda.stream(lane.stream):
            out, info = LD.compress_sharded_finish(st, dst=dst, wait=False)
            torch.cuda.current_stream(lane.eng.dev).synchronize()      # this rank's share of the field is done
        return out, info

    def run_steps(k, one_lane=False):
        """k steps, at most one per lane in flight; -> (slots, info) of the last one.  Results are not hoarded:
        step j's blob is collected (and dropped) while the steps after it run."""
        L = 1 if one_lane else len(lanes)
        last = (None, None)
        if world == 1:
            inflight = []
            for i in range(k):
                inflight.append(lanes.lanes[i % L].submit(job))
                while len(inflight) > L:
                    last = finish(inflight.pop(0))
            while inflight:
                last = finish(inflight.pop(0))
            return last
        open_fields, pending = [], []
        for i in range(k):
            lane = lanes.lanes[i % L]
            open_fields.append((lane, sharded_begin(lane)))
            while len(open_fields) > L - 1:
                pending.append(sharded_finish(*open_fields.pop(0)))
            while len(pending) > 1:
                out, info = pending.pop(0)
                last = (out.result() if hasattr(out, "result") else out, info)
        while open_fields:
            pending.append(sharded_finish(*open_fields.pop(0)))
        while pending:
            out, info = pending.pop(0)
            last = (out.result() if hasattr(out, "result") else out, info)
        return last

    for i in range(len(lanes)):          # every lane once on its own (allocations), then together
        if world == 1:
            slots, info = finish(lanes.lanes[i].submit(job))
        else:
            out, info = sharded_finish(lanes.lanes[i], sharded_begin(lanes.lanes[i]))
            slots = out.result() if hasattr(out, "result") else out
    slots, info = run_steps(max(args.warmup, 1))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    slots, info = run_steps(len(lanes))  # untimed again: the GPUs idled while the clock sampler started
    barrier()
    launches0 = lanes.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    e0.record()
    slots, info = run_steps(args.steps)  # all K blobs are in host memory before the clock stops
    e1.record()
    barrier()
    w1 = time.perf_counter()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = lanes.launches() - launches0
    clocks = sampler.stop(w0, w1) if rank == 0 else None
    # Per-layer kernel times (CUDA events around every layer's launch, on the launching stream): taken over the same
    # K steps run again one after the other on lane 0 -- with two fields in flight an event pair would also span
    # the other lane's kernels.  The same pass gives the one-field-at-a-time step time.
    barrier()
    eng.profile(True)
    eng.profile_read()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    run_steps(args.steps, one_lane=True)
    s1.record()
    barrier()
    seq_ms = max_over_ranks(s0.elapsed_time(s1)) / args.steps
    layer_ms, layer_cnt = eng.profile_read()
    eng.profile(False)
    value = in_bytes * args.steps / (ms * 1e-3) / 1e9

    # ---- N > 1: the sharded blob against a single-GPU call on the same field (outside the timed region)
    sharded = None
    if world > 1:
        verdict = [None]
        if rank == dst and not args.skip_parity:
            full = torch.empty((T, H, W, 2), dtype=torch.float32, device=eng.dev)
            N.check(eng.lib.lc_synth_field(full.data_ptr(), T, H, W, 0, 0, 1234, None), "lc_synth_field")
            ref, rinfo = compress_device(eng, full, ABS_ERR, mode="strict", codec=args.codec, return_parts=True)
            same_mask = bool(torch.equal(eng.unpack_mask(slots[7]), rinfo["mask"]))
            same_q = bool(torch.equal(eng.unpack_codes(slots[6]), rinfo["q"]))
            whole = decompress_device(eng, slots)
            same_x = bool(torch.equal(whole, decompress_device(eng, ref)))
            v, mx = eng.verify(full, whole, ABS_ERR)
            verdict[0] = {"mask_equal": same_mask, "codes_equal": same_q, "xhat_equal": same_x,
                          "stats_equal": bool(slots[4] == ref[4] and slots[5] == ref[5]),
                          "violations": v, "max_abs_err": mx,
                          "cf_sharded": in_bytes / len(pickle.dumps(slots)),
                          "cf_single_call": in_bytes / len(pickle.dumps(ref))}
            del full, ref, rinfo, whole
            torch.cuda.empty_cache()
        dist.broadcast_object_list(verdict, src=dst)
        sharded = verdict[0]

    blob_len = len(pickle.dumps(slots)) if slots is not None else 0
    if world > 1:
        bl = [blob_len]
        dist.broadcast_object_list(bl, src=dst)
        blob_len = bl[0]
    cf = in_bytes / blob_len

    # ---- decompress (device resident) + exhaustive bound check of every rank's rows
    def dstep():
        if world == 1:
            return decompress_device(eng, slots)
        return LD.decompress_sharded(eng, slots, src=dst)[2]

    barrier()
    d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(max(min(args.warmup, 2), 1)):
        xhat = dstep()
    barrier()
    d0.record()
    for _ in range(args.steps):
        xhat = dstep()
    d1.record()
    barrier()
    dms_seq = max_over_ranks(d0.elapsed_time(d1))
    dms = dms_seq
    if world == 1 and len(lanes) > 1:
        # the same K decompressions with one blob per lane in flight (host work of one behind the kernels of the other)
        def djob(lane):
            out = decompress_device(lane.eng, slots)
            torch.cuda.current_stream(lane.eng.dev).synchronize()
            return out

        def drun(k):
            inflight, last = [], None
            for _ in range(k):
                inflight.append(lanes.submit(djob))
                while len(inflight) > len(lanes):
                    last = inflight.pop(0).result()
            while inflight:
                last = inflight.pop(0).result()
            return last
        drun(2 * len(lanes))
        torch.cuda.synchronize()
        d0.record()
        xhat = drun(args.steps)
        d1.record()
        torch.cuda.synchronize()
        dms = d0.elapsed_time(d1)
    dvalue = in_bytes * args.steps / (dms * 1e-3) / 1e9
    viol, max_err = eng.verify(x, xhat, ABS_ERR) if t1r > t0r else (0, 0.0)
    if world > 1:
        vv = torch.tensor([float(viol), max_err], dtype=torch.float64, device=eng.dev)
        dist.all_reduce(vv[:1], op=dist.ReduceOp.SUM)
        dist.all_reduce(vv[1:], op=dist.ReduceOp.MAX)
        viol, max_err = int(vv[0].item()), float(vv[1].item())

    # ---- end to end through the public API with host buffers (pinned input rows, blob bytes out)
    host = torch.empty(tuple(x.shape), dtype=torch.float32, pin_memory=True)
    host.copy_(x)
    xh = host.numpy()

    def e2e_step():
        if world == 1:
            return compress(xh, ABS_ERR, precision=args.precision, codec=args.codec)
        return LD.compress_sharded_host(xh, ABS_ERR, t0=t0r, T=T, precision=args.precision, codec=args.codec, dst=dst)

    def e2e_stream(k, want=None):
        """K fields through the bulk-job API: every step still copies its input rows host->device and its blob
        device->host, but the copies of neighbouring steps overlap the kernels.  The blobs are consumed as they
        arrive (compared with `want` and dropped -- a consumer that hoards them pays first-touch page faults for
        every new blob instead of reusing warm pages)."""
        if world == 1:
            it = compress_stream([xh] * k, ABS_ERR, precision=args.precision, codec=args.codec, lanes=len(lanes))
        else:
            it = LD.compress_sharded_stream([xh] * k, ABS_ERR, t0=t0r, T=T, precision=args.precision,
                                            codec=args.codec, dst=dst, lanes=len(lanes))
        same, last = True, None
        for blob in it:          # lengths on the fly; the last blob is compared byte for byte after the clock stops
            same = same and (want is None or blob is None or len(blob) == len(want))
            last = blob
        return same, last

    b = e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        b = e2e_step()
    torch.cuda.synchronize()
    e2e_single = max_over_ranks((time.perf_counter() - t0) / args.steps)
    e2e_stream(2)
    barrier()
    t0 = time.perf_counter()
    same_len, last_blob = e2e_stream(args.steps, want=b)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / args.steps)
    stream_equal = bool(same_len and (b is None or last_blob == b))
    e2e = {"value": in_bytes / e2e_s / 1e9, "unit": UNIT, "h2d_bytes_per_step": in_bytes * 2,
           "d2h_bytes_per_step": blob_len, "ms_per_step": e2e_s * 1e3,
           "api": ("lossycomp.compress.compress_stream" if world == 1 else "lossycomp.dist.compress_sharded_stream")
                  + ": K fields from pinned host memory -> K blobs (bytes); every step copies its 2-channel input "
                    "host->device and its blob device->host inside the timed region, overlapped with the kernels "
                    "of the neighbouring steps (%d fields in flight per GPU)" % len(lanes),
           "single_call_GBps": in_bytes / e2e_single / 1e9, "single_call_ms": e2e_single * 1e3,
           "single_call_api": "lossycomp.compress.compress" if world == 1 else "lossycomp.dist.compress_sharded_host",
           "stream_blobs_equal_single_call": stream_equal}
    if world == 1:
        xp = np.array(xh)                       # the same field in ordinary (pageable) memory
        compress(xp, ABS_ERR, precision=args.precision, codec=args.codec)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            compress(xp, ABS_ERR, precision=args.precision, codec=args.codec)
        e2e["pageable_input_GBps"] = in_bytes / ((time.perf_counter() - t0) / args.steps) / 1e9
        del xp
        xr = decompress(b)                      # warm-up (pinned staging, allocator)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            xr = decompress(b)
        e2e_d = (time.perf_counter() - t0) / args.steps
        # bulk-job form: K blobs -> K arrays (a ring of result arrays, as a consumer that writes them out would keep)
        ring = [np.empty(xr.shape, dtype=np.float32) for _ in range(len(lanes) + 2)]
        outs = [ring[i % len(ring)] for i in range(args.steps)]
        for _ in decompress_stream([b] * (2 * len(lanes)), lanes=len(lanes), outs=[ring[i % len(ring)] for i in range(2 * len(lanes))]):
            pass
        t0 = time.perf_counter()
        for xs_ in decompress_stream([b] * args.steps, lanes=len(lanes), outs=outs):
            pass
        e2e["decompress_stream_GBps"] = in_bytes / ((time.perf_counter() - t0) / args.steps) / 1e9
        e2e["decompress_stream_equal_single_call"] = bool(np.array_equal(xs_, xr))
        del ring, outs
        e2e_viol = int((np.abs(xh[..., 0].astype(np.float64) - xr[..., 0]) > ABS_ERR).sum())
    else:
        a_, b_, xr = LD.decompress_sharded_host(b, precision=args.precision, src=dst)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            a_, b_, xr = LD.decompress_sharded_host(b, precision=args.precision, src=dst)
        e2e_d = max_over_ranks((time.perf_counter() - t0) / args.steps)
        e2e_viol = int((np.abs(xh[..., 0].astype(np.float64) - xr[..., 0]) > ABS_ERR).sum()) if t1r > t0r else 0
        vv = torch.tensor([float(e2e_viol)], dtype=torch.float64, device=eng.dev)
        dist.all_reduce(vv, op=dist.ReduceOp.SUM)
        e2e_viol = int(vv.item())
    e2e.update(decompress_GBps=in_bytes / e2e_d / 1e9, violations=e2e_viol)

    # per-layer times of every rank -> rank 0 reports the slowest rank's conv stack
    conv_ms = sum(layer_ms)
    if world > 1:
        lm = torch.tensor(layer_ms + [float(n_chunks)], dtype=torch.float64, device=eng.dev)
        alls = [torch.zeros_like(lm) for _ in range(world)]
        dist.all_gather(alls, lm)
        heavy = max(range(world), key=lambda r: float(alls[r][:-1].sum()))
        layer_ms = alls[heavy][:-1].cpu().tolist()
        n_chunks = int(alls[heavy][-1].item())
        conv_ms = sum(layer_ms)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel: the tcgen05 plane-sweep kernel of the 3x3x3 layers
    pk, pk_src = peaks()
    macs = eng.model.arch.macs_per_chunk()
    layers = eng.model.layers
    backends = eng.layer_backends()
    shape = (16, 48, 48)
    per_layer, groups = [], {}
    for i, L in enumerate(layers):
        fl = 2.0 * L.macs(shape) * n_chunks * args.steps       # algorithmic (direct conv) flops
        shape = L.out_shape(shape)
        kname = eng.BACKENDS[backends[i]]
        per_layer.append({"layer": f"{L.part}/{L.name}", "kernel": kname, "ms": layer_ms[i], "launches": layer_cnt[i],
                          "tflops": (fl / (layer_ms[i] * 1e-3) / 1e12) if layer_ms[i] > 0 else None})
        g = groups.setdefault(kname, {"ms": 0.0, "flops": 0.0, "launches": 0})
        g["ms"] += layer_ms[i]; g["flops"] += fl; g["launches"] += layer_cnt[i]
    conv_flops = 2.0 * (macs["Encoder"] + macs["Decoder"]) * n_chunks * args.steps
    if args.precision == "fp32":
        peak, peak_note = 148 * 128 * 2 * 1.965e9 / 1e12, "nominal fp32 FMA peak (148 SM x 128 lanes x 1.965 GHz), CUDA-core path"
        burst = peak
    else:
        peak, peak_note = pk["bf16_tflops_sustained"], f"dense 16-bit tensor, sustained, {pk_src}"
        burst = pk.get("bf16_tflops", peak)
    dom_name = max(groups, key=lambda k: groups[k]["ms"])
    dom = groups[dom_name]
    achieved = dom["flops"] / (dom["ms"] * 1e-3) / 1e12 if dom["ms"] > 0 else None
    chunks_per_launch = n_chunks * args.steps * sum(1 for bk in backends if eng.BACKENDS[bk] == dom_name) / max(dom["launches"], 1)
    traffic, traffic_src = measured_traffic(dom_name, chunks_per_launch)
    roofline = {
        "bound": "tensor", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
        "kernel_note": "layer groups are named after their convolution mode; the K3 group = the four 3x3x3 layers, which "
                       "run on sweep_acc_kernel<3> (time taps accumulated in TMEM, tcgen05.mma.cta_group::2 on SM pairs) "
                       "unless LOSSYCOMP_NO_ACC / LOSSYCOMP_ACC_CTA2=0 say otherwise; K4 = first layer, same kernel",
        "frac": (achieved / peak) if achieved else None,
        "frac_of_burst_peak": (achieved / burst) if achieved else None, "burst_peak": burst,
        "traffic": traffic,
        "traffic_note": (f"dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over this kernel's layers, from "
                         f"{traffic_src}") if traffic else "no ncu --set full capture of this chunk count is committed",
        "timing_note": "per-layer CUDA events over the same K steps run one field at a time right after the timed "
                       "region (with several fields in flight an event pair would span the other lane's kernels)",
        "peak_source": peak_note, "launch_ms": dom["ms"] / max(dom["launches"], 1), "launches": dom["launches"],
        "algorithmic_flops_per_launch": dom["flops"] / max(dom["launches"], 1),
        "share_of_step": dom["ms"] / ms,
        "conv_stack": {"achieved": conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms else None, "ms_per_step":
                       conv_ms / args.steps, "share_of_step": conv_ms / ms, "flops_per_chunk": 2.0 * (macs["Encoder"] + macs["Decoder"]),
                       "frac": (conv_flops / (conv_ms * 1e-3) / 1e12 / peak) if conv_ms else None},
        "decompress": {"value": dvalue, "unit": UNIT, "ms_per_step": dms / args.steps,
                       "frac_of_decoder_ceiling": dvalue / (world * peak * 1e12 / (2.0 * macs["Decoder"] / (16 * 48 * 48 * 4)) / 1e9)},
        "per_kernel": {k: {"ms_per_step": v["ms"] / args.steps, "tflops": v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] else None,
                           "launches": v["launches"]} for k, v in groups.items()},
        "per_layer": per_layer,
    }
    e2e.update(compression_factor=cf, device_violations=viol, max_abs_err=max_err,
               outlier_fraction=(info["n_out"] / max((t1r - t0r) * H * W, 1)))
    if sharded is not None:
        e2e["sharded_parity"] = bool(sharded["mask_equal"] and sharded["codes_equal"] and sharded["xhat_equal"]
                                     and sharded["stats_equal"])
        e2e["sharded"] = sharded
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32", "bf16": "bf16", "fp16": "f16"}[args.precision], "data": "synthetic",
        "config": workload_config(world, T, H, W, n_chunks_all, args.codec, args.precision, my_bytes, dst),
        "lanes": {"fields_in_flight_per_gpu": len(lanes), "one_field_at_a_time_ms_per_step": seq_ms,
                  "note": "lossycomp.lanes: consecutive steps alternate between engines (own stream, workspace, staging "
                          "buffers) so that the kernels of one field fill the host round trips of the other; N = 1: one "
                          "host thread per lane; N > 1: a software pipeline in one thread (begin field k+1, then "
                          "finish field k) so that all collectives keep one order on every rank; every field is "
                          "compressed by the single-lane code and the blobs are identical"},
        "clocks": clocks, "gpu_launches": launches,
        "e2e": e2e,
        "decompress": {"value": dvalue, "unit": UNIT, "ms_per_step": dms / args.steps,
                       "one_blob_at_a_time_ms_per_step": dms_seq / args.steps},
        "compression_factor": cf, "violations": viol, "max_abs_err": max_err,
        "roofline": roofline,
    }
    if world == 1:
        line["blob_bytes"] = {"latent": info["latent_bytes"], "codes": info["error_bytes"], "mask": info["mask_bytes"]}
    else:
        line["sharded_parity"] = e2e.get("sharded_parity")
    if not args.no_cpu_baseline and world == 1:
        cb, xs, parts = cpu_leg(1, 1, CPU_SAMPLE)
        line["cpu_baseline"] = cb
        # the oracle's own input: (1) integer stages bit for bit given the oracle's reconstruction (2.9 M voxels,
        # 2160 scan tiles), (2) compression factor within 1 % of the reference's coder (north_star)
        nb = xs[..., 0].nbytes
        xd = torch.from_numpy(xs).to(eng.dev)
        mask, q = eng.quantize(xd, torch.from_numpy(np.ascontiguousarray(parts["recon_std"][..., 0])).to(eng.dev),
                               parts["mean"], parts["std"], ABS_ERR)
        e2e["oracle_parity_mask_codes"] = bool(np.array_equal(mask.cpu().numpy().reshape(parts["mask"].shape), parts["mask"])
                                               and np.array_equal(q.cpu().numpy(), parts["q"]))
        e2e["cf_same_input"] = {
            "oracle_bz2": cb["compression_factor"],
            "gpu_parity_mode_bz2_codec": nb / len(compress(xs, ABS_ERR, precision=args.precision, codec="bz2", mode="parity")),
            "gpu_strict_mode_bz2_codec": nb / len(compress(xs, ABS_ERR, precision=args.precision, codec="bz2")),
            "gpu_strict_mode_huff_codec": nb / len(compress(xs, ABS_ERR, precision=args.precision, codec="huff")),
            "input": "oracle.synthetic_field%s, abs_err %g" % (str(CPU_SAMPLE), ABS_ERR)}
        line["cf_same_input"] = e2e["cf_same_input"]
        # BASELINE configs[3] in small: three variables (temperature; a humidity-like field of small positive values;
        # a geopotential-like field of large values) x the abs_err sweep, relative to each variable's spread; every
        # round trip is checked element by element on the host
        sweep = {}
        base = xs[..., 0].astype(np.float64)
        sd = float(base.std())
        variables = {"temperature": (1.0, 0.0), "humidity_like": (2.5e-3 / sd, -float(base.min()) * 2.5e-3 / sd + 1e-4),
                     "geopotential_like": (3.0e3 / sd, 5.0e4)}
        for vname, (scale, offset) in variables.items():
            xv = np.array(xs)
            xv[..., 0] = (base * scale + offset).astype(np.float32)
            for rel in (1e-3, 1e-2, 0.1, 0.3, 1.0):
                thr = float(rel * scale)                       # the sweep's abs_err in the variable's own units
                bl = compress(xv, thr, precision=args.precision, codec=args.codec)
                xr = decompress(bl)
                err = np.abs(xv[..., 0].astype(np.float64) - xr[..., 0])
                sweep[f"{vname}@{rel:g}"] = {"abs_err": thr, "compression_factor": nb / len(bl),
                                             "violations": int((err > thr).sum()), "max_abs_err": float(err.max())}
        e2e["variable_sweep"] = sweep
        e2e["variable_sweep_violations"] = int(sum(v["violations"] for v in sweep.values()))
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
