#!/usr/bin/env python
"""Headline benchmark: event clips/sec of the SlowFast-R50 8x8 feature-extraction forward
(BASELINE.json configs[1]: batch 64 synthetic event clips, bf16, per B200).

    python bench.py --gpus N --steps K --warmup W            # ours (tcgen05 kernels through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...   # the reference path on the host CPU cores

A step = one pass of the hot path over one batch of 64 clips per GPU: frame pack (uint8 ->
bf16 NTHWC, both pathways) + 105 conv launches + pools + GAP + projection head (+ for N>1 the
all-gather of the [64, 2304] features).  `value` times it with the uint8 frames already in HBM;
`e2e` times the same through vidsitu_b200.HostPipeline with the frames in pinned host memory
(H2D of every batch and D2H of its features inside the timed region).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

MODEL = "slow_fast_nl_r50_8x8"
WORKLOAD = "SlowFast-R50 8x8 feature extraction, batch 64 synthetic event clips (32x224x224 fast / 8 slow) per GPU"
GFLOP_PER_CLIP = 100.615  # SURVEY.md section 8d / appendix A.1 (asserted in tests/test_host.py)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1383.6), d.get("bf16_tflops", 1636.3), d.get("hbm_gbs", 6529.7), "measured"
    return 1400.0, 1590.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_clips_per_s(steps: int, warmup: int, clips: int = 5):
    """The reference algorithm on the host cores: the oracle port (oracle/sf_oracle.py, validated
    against the real reference in tests/test_oracle.py), fp32, eval, all threads; one step = one
    video = 5 event clips (BASELINE.json configs[0])."""
    import torch
    from common import build_model, synthetic_frames
    from oracle import sf_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model, cfg, _ = build_model(MODEL, seed=0, crop=224)
    sd = model.state_dict()
    frames = synthetic_frames(clips, 32, 224, seed=1234)
    xs = O.clips_from_frames(frames, cfg.sf_mdl)
    for _ in range(warmup):
        O.sfbase_forward(sd, cfg.sf_mdl, xs)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.sfbase_forward(sd, cfg.sf_mdl, xs)
    dt = time.perf_counter() - t0
    return clips * steps / dt, dt / steps * 1e3, cores


def gpu_library_baseline(model, cfg, host_frames, dev, iters: int = 3):
    """Informational: the reference's own torch modules (oracle restatement of the same forward: F.conv3d ->
    cuDNN 9, F.batch_norm, ...) on THIS GPU at the bench batch, (a) fp32 with TF32 off = "reference PyTorch fp32 on
    B200", (b) bf16 channels_last_3d = the cuDNN Blackwell bar (SURVEY.md 2.1 / BASELINE.md 3).  Inputs already on
    the device and normalised (the frame pack is not part of this leg).  Returns clips/s of both."""
    import torch
    from oracle import sf_oracle as O

    out = {}
    n = host_frames.shape[0]
    xs32 = [x.to(dev) for x in O.clips_from_frames(host_frames, cfg.sf_mdl)]
    sd32 = {k: v.detach().to(dev, torch.float32) for k, v in model.state_dict().items()}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.benchmark = True
    try:
        for name, tf32, dt in (("torch_fp32_tf32off", False, torch.float32), ("torch_bf16_channels_last_3d", True, torch.bfloat16)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            if dt is torch.float32:
                sd, xs = sd32, xs32
            else:
                sd = {k: (v.to(dt).contiguous(memory_format=torch.channels_last_3d) if v.dim() == 5 else v.to(dt))
                      for k, v in sd32.items()}
                xs = [x.to(dt).contiguous(memory_format=torch.channels_last_3d) for x in xs32]
            try:
                for _ in range(2):
                    O.sfbase_forward(sd, cfg.sf_mdl, xs)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(iters):
                    O.sfbase_forward(sd, cfg.sf_mdl, xs)
                e1.record()
                e1.synchronize()
                ms = e0.elapsed_time(e1) / iters
                out[name] = {"clips_per_s": round(n / (ms / 1e3), 1), "ms_per_step": round(ms, 2)}
            except Exception as e:  # noqa: BLE001 -- informational leg: never fails the bench
                out[name] = {"error": f"{type(e).__name__}: {str(e)[:160]}"}
            del sd, xs
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old
    out["note"] = ("oracle restatement of the reference modules run through torch/cuDNN on this GPU, batch "
                   f"{n}, device-resident normalised inputs, forward + head; informational, not a fallback")
    return out


def run_reference(args, rank: int):
    if rank != 0:
        return
    steps = max(1, min(args.steps, 40))
    val, ms, cores = cpu_reference_clips_per_s(steps, max(1, min(args.warmup, 3)))
    sample = f"{steps} steps x 5 clips (1 synthetic video) of the same SlowFast-R50 8x8 224x224 forward, fp32"
    print(json.dumps({
        "impl": "reference", "metric": "event clips/sec SlowFast-R50 8x8", "value": round(val, 4), "unit": "clips/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": round(ms, 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "SlowFast-R50 8x8 forward, 1 synthetic video x 5 event clips (32x224x224 fast / 8 slow) per "
                               "step, fp32 on the host CPU cores (BASELINE.json configs[0]); clips/s is batch-size "
                               "independent on this arm", "clips_per_step": 5, "headline_workload": WORKLOAD,
                   "note": "reference path = CPU fp32 forward (the reference's own test-free "
                   "PyTorch modules restated in oracle/sf_oracle.py and pinned to the real reference's outputs)"},
        "cpu_baseline": {"value": round(val, 4), "unit": "clips/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(val, 4), "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--model", default=MODEL, help="sf_mdl_name of another backbone (BASELINE.json configs 3/4); "
                    "the default is the headline SlowFast-R50 8x8")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true",
                    help="skip the informational torch/cuDNN leg (reference modules on this GPU, fp32 and bf16)")
    ap.add_argument("--overlap-pack", action="store_true",
                    help="two input slots: the pack of batch k+1 runs on its own stream while the trunk reads batch k "
                         "(measured on B200: no gain, 5628 vs 5665 clips/s - the persistent conv CTAs leave it no room)")
    ap.add_argument("--per-op", default="", help="write per-launch timings (json) to this path")
    ap.add_argument("--profile-range", action="store_true",
                    help="cudaProfilerStart/Stop around the timed steps (ncu --profile-from-start off)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from common import build_model, synthetic_frames
    from vidsitu_b200 import lib
    from vidsitu_b200.dist import gather_rows
    from vidsitu_b200.pipeline import HostPipeline

    args.warmup = max(args.warmup, 3)
    # one process per GPU, pinned to the CPUs local to its GPU; the two pinned frame batches are first-touched on
    # different NUMA nodes (when there are two) so that N ranks' H2D copies do not all read one memory controller
    from vidsitu_b200 import numa
    cpus = numa.bind_to_gpu(local_rank) if os.environ.get("VSB_NUMA_BIND", "1") == "1" else None
    nodes = sorted(numa.numa_nodes()) if os.environ.get("VSB_NUMA_BIND", "1") == "1" else []
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    model, cfg, _ = build_model(args.model, seed=0, crop=224, micro_batch=B)
    model.input_slots = 2 if args.overlap_pack else 1
    model = model.to(dev)
    t_frames = cfg.sf_mdl.DATA.NUM_FRAMES
    # two distinct synthetic batches per rank; 2 x 308 MB of uint8 frames (> the 126 MB L2) alternate between steps
    host_frames = []
    for i in range(2):
        f = synthetic_frames(B, t_frames, 224, seed=1234 + 17 * rank + i)
        with numa.on_node(nodes[(rank + i) % len(nodes)] if len(nodes) > 1 else None):
            host_frames.append(f.pin_memory())
    dev_frames = [f.to(dev) for f in host_frames]
    eng = model._engine(B, dev)
    eng.capture()
    launches_per_step = eng.num_launches + (2 if model.spec.num_pathways == 2 else 1)

    nslots = len(eng.input_sets)
    pack_stream = torch.cuda.Stream(dev) if nslots > 1 else None
    packed = [torch.cuda.Event() for _ in range(nslots)]
    in_free = [torch.cuda.Event() for _ in range(nslots)]
    for e in in_free:
        e.record()

    def step(i):
        """One batch: frame pack (uint8 -> bf16 NTHWC, both pathways) + the graph-replayed trunk and head.  With two
        input slots the pack of this batch runs on its own stream and may overlap the previous batch's trunk."""
        if nslots == 1:
            eng.load_frames(dev_frames[i % 2])
            eng.replay()
        else:
            slot = i % nslots
            main = torch.cuda.current_stream()
            with torch.cuda.stream(pack_stream):
                pack_stream.wait_event(in_free[slot])       # the trunk that last read this slot has finished
                eng.load_frames(dev_frames[i % 2], slot)
                packed[slot].record(pack_stream)
            main.wait_event(packed[slot])
            eng.replay(slot)
            in_free[slot].record(main)
        if world > 1:
            gather_rows(eng.feats, world * B, rows_per_item=B)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if args.profile_range:
        torch.cuda.profiler.start()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    barrier()
    if args.profile_range:
        torch.cuda.profiler.stop()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * B * args.steps / (ms / 1e3)

    # ---- end to end through the public host API (pinned host frames -> pinned host features)
    e2e = None
    if not args.no_e2e:
        pipe = HostPipeline(model, B, dev)
        for i in range(args.warmup):
            pipe.submit(host_frames[i % 2])
        pipe.flush()
        barrier()
        t0 = time.perf_counter()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(pipe.copy_stream)
        checksum = 0.0
        pending = []
        for i in range(args.steps):
            pending.append(pipe.submit(host_frames[i % 2]))
            if len(pending) > 1:
                checksum += float(pipe.result(pending.pop(0))[0, 0])   # the host really reads each result
        while pending:
            checksum += float(pipe.result(pending.pop(0))[0, 0])
        s1.record(pipe.compute_stream)
        pipe.flush()
        wall_ms = (time.perf_counter() - t0) * 1e3
        dev_ms = s0.elapsed_time(s1)
        t = torch.tensor([max(dev_ms, wall_ms)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
        e2e = {"value": round(world * B * args.steps / (e2e_ms / 1e3), 2), "unit": "clips/s",
               "h2d_bytes_per_step": pipe.h2d_bytes, "d2h_bytes_per_step": pipe.d2h_bytes,
               "ms_per_step": round(e2e_ms / args.steps, 3), "checksum": round(checksum, 4)}

    # ---- roofline of the conv launches.  The step is a CUDA graph, so the conv time INSIDE the timed region is the
    # graph-replayed step time minus the (eagerly timed) frame pack, times the conv launches' share of the eagerly
    # timed trunk + head (CUDA events around every launch; the ncu launch list under profiles/ gives the same share).
    peak_sus, peak_burst, hbm, peak_src = measured_peaks()
    per_op = eng.time_ops(iters=3) if rank == 0 else []
    roofline = None
    if rank == 0:
        ms_step = ms / args.steps
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pack_ms = float("inf")
        for i in range(3):
            p0.record()
            eng.load_frames(dev_frames[i % 2])
            p1.record()
            p1.synchronize()
            pack_ms = min(pack_ms, p0.elapsed_time(p1))
        conv = [(n, ms_, f) for n, ms_, f in per_op if f > 0 and not n.startswith("proj_head")]
        conv_eager = sum(ms_ for _, ms_, _ in conv)
        all_eager = sum(ms_ for _, ms_, _ in per_op)
        share = conv_eager / all_eager
        conv_ms = max(ms_step - pack_ms, 0.0) * share          # conv time inside the graph-replayed step
        assert conv_ms <= ms_step
        flops = (GFLOP_PER_CLIP * 1e9 if args.model == MODEL else eng.conv_flops / B) * B
        achieved = flops / (conv_ms / 1e3) / 1e12
        step_tf = flops / (ms_step / 1e3) / 1e12
        traffic = traffic_all = None   # DRAM bytes of one step, from the committed ncu capture of this round
        tsrc = None
        for cand in ("r02d_step_dram_traffic.json", "r02c_step_dram_traffic.json", "r02_step_dram_traffic.json", "r01_step_dram_traffic.json"):
            tp = os.path.join(ROOT, "profiles", cand)
            if os.path.exists(tp) and B == 64 and args.model == MODEL:
                ks = json.load(open(tp))["kernels"]
                traffic = int(sum(k["dram_read_bytes"] + k["dram_write_bytes"] for k in ks if k["kernel"].startswith(("conv_", "bottleneck_"))))
                traffic_all = int(sum(k["dram_read_bytes"] + k["dram_write_bytes"] for k in ks))
                tsrc = cand
                break
        roofline = {"bound": "tensor", "kernel": "all conv launches of a step (conv_igemm_kernel, conv_igemm2_kernel, "
                    "conv_win_kernel, stem_pool_kernel, bottleneck_thin_kernel)",
                    "achieved": round(achieved, 2), "peak": peak_sus, "unit": "TFLOP/s",
                    "frac": round(achieved / peak_sus, 4), "frac_sustained": round(achieved / peak_sus, 4),
                    "frac_burst": round(achieved / peak_burst, 4), "peak_burst": peak_burst,
                    "peak_source": f"{peak_src}: bf16_tflops_sustained (kernels timed inside a long step) and bf16_tflops (burst)",
                    "traffic": traffic,
                    "traffic_note": f"dram__bytes_read+write of the conv launches of one step (ncu, profiles/{tsrc}); not "
                                    "re-measured by this run",
                    "launches_per_step": len(conv), "conv_ms_per_step": round(conv_ms, 3), "pack_ms_per_step": round(pack_ms, 3),
                    "conv_share_of_trunk_and_head": round(share, 4), "eager_conv_ms_sum": round(conv_eager, 3),
                    "step_tflops": round(step_tf, 2), "step_frac": round(step_tf / peak_sus, 4),
                    "step_frac_burst": round(step_tf / peak_burst, 4),
                    "hbm": None if traffic_all is None else {
                        "dram_bytes_per_step": traffic_all, "gbs": round(traffic_all / (ms_step / 1e3) / 1e9, 1),
                        "peak_gbs": hbm, "frac": round(traffic_all / (ms_step / 1e3) / 1e9 / hbm, 4)}}
        if args.per_op:
            json.dump([{"op": n, "ms": round(m, 4), "gflop": round(f / 1e9, 3),
                        "mbytes": round(eng.op_bytes.get(n, 0) / 1e6, 2)} for n, m, f in per_op],
                      open(args.per_op, "w"), indent=0)

    gpu_baseline = None
    if rank == 0 and world == 1 and not args.no_gpu_baseline:
        try:
            gpu_baseline = gpu_library_baseline(model, cfg, host_frames[0], dev)
            for k, v in gpu_baseline.items():
                if isinstance(v, dict) and "clips_per_s" in v:
                    v["ours_speedup"] = round(value / v["clips_per_s"], 2)
        except Exception as e:  # noqa: BLE001
            gpu_baseline = {"error": f"{type(e).__name__}: {str(e)[:200]}"}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        val, cms, cores = cpu_reference_clips_per_s(steps=12, warmup=1)   # about 10 s of CPU work
        cpu_baseline = {"value": round(val, 4), "unit": "clips/s", "cores": cores, "kind": "port",
                        "sample": "12 steps x 5 clips (1 synthetic video, BASELINE.json configs[0]) of the same forward, "
                                  "fp32 oracle port on the host cores"}

    if rank == 0:
        print(json.dumps({
            "metric": "event clips/sec SlowFast-R50 8x8" if args.model == MODEL else f"event clips/sec {args.model}",
            "value": round(value, 2), "unit": "clips/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD if args.model == MODEL else f"{args.model} feature extraction, batch {B} "
                       f"synthetic event clips ({t_frames}x224x224) per GPU", "clips_per_gpu_per_step": B, "model": args.model,
                       "weights": "random-init (seed 0) + seeded BatchNorm statistics",
                       "l2": "inputs larger than L2: 2 alternating 308 MB uint8 frame batches per GPU, "
                             "activations >> 126 MB", "cuda_graph": True, "replay": ("clip program: one vsb_program_run per step (C ABI v8, graph captured inside the library)" if eng._use_program() else "torch CUDA graph of the Python launch loop"), "pack_overlap": nslots > 1,
                       "parallelism": f"clip-sharded x{world}, features all-gathered each step" if world > 1 else "single GPU",
                       "host_placement": {"rank0_cpus": (f"{cpus[0]}-{cpus[-1]} ({len(cpus)})" if cpus else None),
                                          "numa_nodes": nodes}},
            "e2e": e2e, "gpu_launches": launches_per_step * args.steps, "launches_per_step": launches_per_step,
            "roofline": roofline, "cpu_baseline": cpu_baseline,
            "cpu_baseline_note": None if cpu_baseline is not None else "measured on rank 0 at N=1 only (see --impl reference)",
            "gpu_baseline": gpu_baseline, "fused_blocks": len(getattr(eng, "fused_blocks", [])),
            "fused_stems": list(getattr(eng, "fused_stems", [])), "clocks": clocks,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
