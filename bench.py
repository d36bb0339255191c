#!/usr/bin/env python
"""Benchmark of the lossycomp hot path on B200 (contract: see the task's bench.py section).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision P]

A step = one compress() of one synthetic ERA5-shaped field (BASELINE.json configs[1]:
24 h x 1 level x 720 x 1440 float32 + land-sea channel, abs_error 0.1) per GPU.  Under torchrun
every rank compresses its own 24-hour time shard of the month (weak scaling, no data-path
collective: chunks are independent); value = bytes of channel-0 input all ranks compressed / max
over ranks of the device time.

Keys beyond the base contract: `roofline` (dominant kernel = the convolution stack, tensor bound
against the measured bf16 peak, or CUDA-core fp32 when --precision fp32), `cpu_baseline` (the
oracle timed on the host cores over a bounded sample), `decompress`, `compression_factor`,
`violations`.
"""
import argparse
import json
import os
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
CPU_SAMPLE = (24, 336, 720)      # bounded sample of the same field for the CPU legs (~10-20 s)
# DRAM bytes per chunk per launch of the dominant kernel, from the ncu --set full capture under profiles/
NCU_DRAM_BYTES_PER_CHUNK = {"sweep_tc_kernel<K3>": 4.16e6}   # mean of its 4 layers: 3.08 MB (plain), 5.24 MB (with skip)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("LOSSYCOMP_PRECISION", "fp16"))
    ap.add_argument("--codec", default="huff")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", action="store_true",
                    help="N>1: also gather every rank's blob into rank 0's HBM inside the timed region")
    ap.add_argument("--field", default=None, help="T,H,W override (tests)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------
def cpu_leg(steps, warmup, sample=CPU_SAMPLE):
    """The reference's CPU path (oracle port: NumPy + torch-CPU fp32 convs, batch 10, bz2 -9)."""
    import torch
    from lossycomp.models import load_model
    from oracle import lossycomp_oracle as O
    model = load_model()
    T, H, W = sample
    x = O.synthetic_field(T, H, W, seed=1234)
    nbytes = x[..., 0].nbytes
    for _ in range(max(0, min(warmup, 1))):
        O.compress(x[:16, :96, :96], ABS_ERR, model)
    times, dtimes = [], []
    blob = None
    for _ in range(steps):
        t0 = time.perf_counter()
        blob = O.compress(x, ABS_ERR, model)
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
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    cb = cpu_leg(steps, args.warmup)
    T, H, W = CPU_SAMPLE
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": cb["seconds_per_step"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"compress() of one synthetic ERA5 slice 24x720x1440 fp32 + lsm, abs_err {ABS_ERR}; "
                               f"each step a bounded {T}x{H}x{W} sample", "model": "final_models/model_2"},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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


def run_ours(args):
    import torch
    import torch.distributed as dist
    from lossycomp import _native as N
    from lossycomp._engine import Engine
    from lossycomp.compress import compress, compress_device
    from lossycomp.decompress import decompress, decompress_device

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N.lib()
    eng = Engine.get(None, args.precision, local)
    T, H, W = FIELD if not args.field else tuple(int(v) for v in args.field.split(","))
    # synthetic month: rank r owns hours [24 r, 24 r + 24) (time sharding in whole 16-step chunks is
    # not needed here because every rank's shard is its own compress() call)
    x = torch.empty((T, H, W, 2), dtype=torch.float32, device=eng.dev)
    N.check(eng.lib.lc_synth_field(x.data_ptr(), T, H, W, 24 * rank, 0, 1234, None), "lc_synth_field")
    torch.cuda.synchronize()
    in_bytes = T * H * W * 4
    n_chunks = eng.n_chunks(T, H, W)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    import pickle
    from lossycomp import dist as LD

    def step():
        # one compress() per 24 h slice (the reference's unit of work).  Slices are independent objects,
        # so ranks exchange nothing (each keeps / writes its own blob); --gather adds an NCCL gather of
        # the blobs into rank 0's HBM to the timed region (what compress_sharded() pays per field)
        slots, info = compress_device(eng, x, ABS_ERR, mode="strict", codec=args.codec)
        if world > 1 and args.gather:
            info["gathered"] = LD.gather_bytes(LD.frame_slots(slots), 0, None, eng.dev, to_host=False)
        return slots, info

    for _ in range(args.warmup):
        slots, info = step()
    eng.profile(True)
    eng.profile_read()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    launches0 = eng.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        slots, info = step()
    e1.record()
    barrier()
    w1 = time.perf_counter()
    ms = e0.elapsed_time(e1)
    launches = eng.launches - launches0
    layer_ms, layer_cnt = eng.profile_read()
    eng.profile(False)
    clocks = sampler.stop(w0, w1) if rank == 0 else None
    if world > 1:
        tms = torch.tensor([ms], dtype=torch.float64, device=eng.dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    value = world * in_bytes * args.steps / (ms * 1e-3) / 1e9

    blob = pickle.dumps(slots)
    cf = in_bytes / len(blob)

    # ---- decompress (device resident) + exhaustive bound check
    barrier()
    d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(max(args.warmup, 1)):
        xhat = decompress_device(eng, slots)
    barrier()
    d0.record()
    for _ in range(args.steps):
        xhat = decompress_device(eng, slots)
    d1.record()
    barrier()
    dms = d0.elapsed_time(d1)
    if world > 1:
        tms = torch.tensor([dms], dtype=torch.float64, device=eng.dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        dms = float(tms.item())
    dvalue = world * in_bytes * args.steps / (dms * 1e-3) / 1e9
    viol, max_err = eng.verify(x, xhat, ABS_ERR)

    # ---- end to end through the public API with host buffers (pinned input, blob bytes out)
    host = torch.empty((T, H, W, 2), dtype=torch.float32, pin_memory=True)
    host.copy_(x)
    xh = host.numpy()
    b = compress(xh, ABS_ERR, precision=args.precision, codec=args.codec)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        b = compress(xh, ABS_ERR, precision=args.precision, codec=args.codec)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / args.steps
    if world > 1:
        tms = torch.tensor([e2e_s], dtype=torch.float64, device=eng.dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        e2e_s = float(tms.item())
    # bulk-job variant: the same K calls through compress_stream(), which overlaps the H2D copy of field k+1 with
    # the kernels of field k (every step still copies its 199 MB in and its blob out)
    from lossycomp.compress import compress_stream
    list(compress_stream([xh, xh], ABS_ERR, precision=args.precision, codec=args.codec))
    barrier()
    t0 = time.perf_counter()
    n_stream = sum(1 for _ in compress_stream([xh] * args.steps, ABS_ERR, precision=args.precision, codec=args.codec))
    torch.cuda.synchronize()
    e2e_stream = (time.perf_counter() - t0) / max(n_stream, 1)
    xp = np.array(xh)                       # the same field in ordinary (pageable) memory
    compress(xp, ABS_ERR, precision=args.precision, codec=args.codec)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        compress(xp, ABS_ERR, precision=args.precision, codec=args.codec)
    e2e_pageable = (time.perf_counter() - t0) / args.steps
    del xp
    xr = decompress(b)                      # warm-up (pinned staging, allocator)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        xr = decompress(b)
    e2e_d = (time.perf_counter() - t0) / args.steps
    e2e_viol = int((np.abs(xh[..., 0].astype(np.float64) - xr[..., 0]) > ABS_ERR).sum())

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
    conv_ms = sum(layer_ms)
    conv_flops = 2.0 * (macs["Encoder"] + macs["Decoder"]) * n_chunks * args.steps
    if args.precision == "fp32":
        peak, peak_note = 148 * 128 * 2 * 1.965e9 / 1e12, "nominal fp32 FMA peak (148 SM x 128 lanes x 1.965 GHz), CUDA-core path"
    else:
        peak, peak_note = pk["bf16_tflops_sustained"], f"dense 16-bit tensor, sustained, {pk_src}"
    dom_name = max(groups, key=lambda k: groups[k]["ms"])
    dom = groups[dom_name]
    achieved = dom["flops"] / (dom["ms"] * 1e-3) / 1e12 if dom["ms"] > 0 else None
    # DRAM traffic of that kernel from the committed ncu --set full capture (profiles/), per chunk
    traffic_per_chunk = NCU_DRAM_BYTES_PER_CHUNK.get(dom_name)
    chunks_per_launch = n_chunks * args.steps * sum(1 for b in backends if eng.BACKENDS[b] == dom_name) / max(dom["launches"], 1)
    roofline = {
        "bound": "tensor", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
        "frac": (achieved / peak) if achieved else None,
        "traffic": (traffic_per_chunk * chunks_per_launch) if traffic_per_chunk else None,
        "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture "
                        "(profiles/r01_ncu_sweep_final.txt), scaled to the chunks one launch processes",
        "peak_source": peak_note, "launch_ms": dom["ms"] / max(dom["launches"], 1), "launches": dom["launches"],
        "algorithmic_flops_per_launch": dom["flops"] / max(dom["launches"], 1),
        "share_of_step": dom["ms"] / ms,
        "conv_stack": {"achieved": conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms else None, "ms_per_step":
                       conv_ms / args.steps, "share_of_step": conv_ms / ms, "flops_per_chunk": 2.0 * (macs["Encoder"] + macs["Decoder"])},
        "per_kernel": {k: {"ms_per_step": v["ms"] / args.steps, "tflops": v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] else None,
                           "launches": v["launches"]} for k, v in groups.items()},
        "per_layer": per_layer,
    }
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32", "bf16": "bf16", "fp16": "f16"}[args.precision], "data": "synthetic",
        "config": {"workload": f"compress() of one synthetic ERA5 slice {T}x{H}x{W} fp32 + lsm channel per GPU, "
                               f"abs_err {ABS_ERR}, strict bound, {n_chunks} chunks of 16x48x48, codec {args.codec}",
                   "model": "final_models/model_2 (F=23,k=4,res=1)", "precision": args.precision,
                   "l2": f"inputs {in_bytes * 2 / 1e6:.0f} MB per step exceed the 126 MB L2",
                   "parallelism": f"{world} rank(s), one independent 24 h slice per rank per step, "
                                  + ("blobs gathered to rank 0 over NCCL" if (world > 1 and args.gather)
                                     else "no data-path collective")},
        "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": world * in_bytes / e2e_s / 1e9, "unit": UNIT, "h2d_bytes_per_step": in_bytes * 2,
                "d2h_bytes_per_step": len(b), "ms_per_step": e2e_s * 1e3,
                "pageable_input_GBps": in_bytes / e2e_pageable / 1e9,
                "compress_stream_GBps": in_bytes / e2e_stream / 1e9,
                "decompress_GBps": in_bytes / e2e_d / 1e9, "violations": e2e_viol},
        "decompress": {"value": dvalue, "unit": UNIT, "ms_per_step": dms / args.steps},
        "compression_factor": cf, "violations": viol, "max_abs_err": max_err,
        "outlier_fraction": info["n_out"] / (T * H * W),
        "blob_bytes": {"latent": info["latent_bytes"], "codes": info["error_bytes"], "mask": info["mask_bytes"]},
        "roofline": roofline,
    }
    if not args.no_cpu_baseline and world == 1:
        cb = cpu_leg(1, 1)
        line["cpu_baseline"] = cb
        # compression factor on the SAME input as the CPU oracle (north_star: within 1 % of the reference)
        from oracle import lossycomp_oracle as O
        xs = O.synthetic_field(*CPU_SAMPLE, seed=1234)
        nb = xs[..., 0].nbytes
        line["cf_same_input"] = {
            "oracle_bz2": cb["compression_factor"],
            "gpu_parity_mode_bz2_codec": nb / len(compress(xs, ABS_ERR, precision=args.precision, codec="bz2", mode="parity")),
            "gpu_strict_mode_bz2_codec": nb / len(compress(xs, ABS_ERR, precision=args.precision, codec="bz2")),
            "gpu_strict_mode_huff_codec": nb / len(compress(xs, ABS_ERR, precision=args.precision, codec="huff")),
            "input": "oracle.synthetic_field%s, abs_err %g" % (str(CPU_SAMPLE), ABS_ERR)}
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
