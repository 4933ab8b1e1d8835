#!/usr/bin/env python
"""Headline benchmark: clips/s of the VideoPrism FactorizedEncoder forward (16x288x288x3 clips).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--model base|large] [--global-batch 32]
    python bench.py --impl reference ...        # the reference's CPU path (oracle port) on the host cores

Workload (BASELINE.json configs[1]): videoprism_public_v1_base, batch 32 synthetic clips, bf16 tensor-core
math with fp32 accumulation, the batch sharded contiguously over the N GPUs of one node (32/N clips per
rank, no inter-GPU traffic inside the encoder: "scaling": "strong", the config as BASELINE.json states it).
`--scaling weak` keeps 32 clips on EVERY rank instead (global batch 32*N).  One "step" = one forward of the
whole batch.

Prints ONE JSON line (rank 0).  `value` is device-resident throughput (inputs already in HBM), `e2e` is the
same metric through the public API with host buffers (H2D of the clips and D2H of the features inside the
timed region), `roofline` is the dominant kernel (FFN1 GEMM) against the measured bf16 peak, and
`cpu_baseline` is the oracle on the host cores (N=1 only).
"""
from __future__ import annotations

import argparse
import datetime
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

try:
    _ORIG_AFFINITY = set(os.sched_getaffinity(0))   # bind_to_gpu_numa_node narrows it; the CPU-baseline legs get it back
except AttributeError:
    _ORIG_AFFINITY = None


def _restore_affinity():
    if _ORIG_AFFINITY:
        try:
            os.sched_setaffinity(0, _ORIG_AFFINITY)
        except OSError:
            pass


GF_PER_CLIP = {"base": 973.29, "large": 2998.52}          # BASELINE.md §3 (2*M*N*K per GEMM + 4*S^2*dh*H per sequence)
MODEL_NAME = {"base": "videoprism_public_v1_base", "large": "videoprism_public_v1_large"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons for one GPU while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None, t_load0=None):
        """sm_mhz / power: samples inside the timed region [t0, t1].  reasons: every throttle reason seen from the start of
        the load (t_load0: first warm-up step, the same kernels back to back) to just after the timed region: nvidia-smi
        refreshes its throttle flags and its (averaged) power reading more slowly than a 5-step timed region lasts."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [ln for ts, ln in self.lines if t0 is None or (t0 <= ts <= t1 + 0.15)]
        load = [ln for ts, ln in self.lines if t0 is None or ((t_load0 if t_load0 is not None else t0) <= ts <= t1 + 0.3)]
        for ln in (load or [ln for _, ln in self.lines]):
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                mx.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        power_load = power
        power = []
        for ln in (inside or load or [ln for _, ln in self.lines]):
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); power.append(float(parts[2]))
            except ValueError:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s, p in zip(sm, power) if p > 300.0] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(power_load or power), "reasons_window": "first warm-up step .. end of the timed region"}


def _parse_cpulist(text: str):
    cpus = set()
    for part in text.strip().split(","):
        part = part.strip()
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def _smi_cpu_affinity(bus_id: str):
    """CPU affinity of the GPU with this PCI bus id from `nvidia-smi topo -m`, or None.  The row of GPU<i> holds link types
    (X, NV18, SYS, PIX, NODE, PHB ...) and then the 'CPU Affinity' cpulist ("0-55,112-167"), the 'NUMA Affinity' node list
    ("0") and the GPU NUMA id: the CPU affinity is the first cpulist-shaped cell that names at least 8 CPUs."""
    import re
    q = subprocess.run(["nvidia-smi", "--query-gpu=index,pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True, timeout=20)
    index = None
    for ln in q.stdout.splitlines():
        idx, _, bid = ln.partition(",")
        if bid.strip().lower().endswith(bus_id.lower()):   # nvidia-smi prints an 8-digit domain, sysfs a 4-digit one
            index = int(idx)
    if index is None:
        return None
    topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout.splitlines()
    shape = re.compile(r"^\d+(-\d+)?(,\d+(-\d+)?)*$")
    for ln in topo:
        cells = [c for c in re.split(r"[\t ]+", ln.strip()) if c]
        if not cells or cells[0] != f"GPU{index}":
            continue
        for c in cells[1:]:
            if shape.match(c):
                cpus = _parse_cpulist(c)
                if len(cpus) >= 8:
                    return cpus
    return None


def bind_to_gpu_numa_node(local: int):
    """Pins this rank (and the pinned host buffers it allocates afterwards: first touch) to the CPUs next to its GPU (sysfs
    numa_node of the PCI device, else the 'CPU Affinity' column of `nvidia-smi topo -m`), so that 8 ranks' H2D / D2H
    streams do not cross the socket interconnect.  Returns a description of what was done, or None (nothing changed)."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        cpus, how = None, None
        try:
            with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
                node = int(f.read().strip())
            if node >= 0:
                with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
                    cpus, how = _parse_cpulist(f.read()), f"numa node {node} (sysfs)"
        except Exception:
            cpus = None
        if not cpus:
            cpus = _smi_cpu_affinity(bus)
            how = "nvidia-smi topo CPU affinity"
        if not cpus:
            return None
        current = set(os.sched_getaffinity(0))
        allowed = cpus & current
        if len(allowed) >= 4 and allowed != current:   # never squeeze a rank onto a handful of cores on a parsing accident
            os.sched_setaffinity(0, allowed)
            return f"{how}: {len(allowed)} cpus"
    except Exception:
        pass
    return None


def pipelined_e2e(submit, wait, n_steps):
    """The serving loop of the asynchronous host API: submit step i, then wait for step i-1 (two calls in flight), so the
    H2D of step i+1 and the D2H of step i-1 overlap the forward of step i.  Every step's copies happen inside the timed
    region; returns wall-clock seconds for n_steps steps (all results delivered)."""
    import torch
    t0 = time.perf_counter()
    prev = None
    for i in range(n_steps):
        tk = submit(i)
        if prev is not None:
            wait(prev)
        prev = tk
    wait(prev)
    torch.cuda.synchronize()
    return time.perf_counter() - t0


def cpu_oracle_clips_per_s(model: str, steps: int, warmup: int, budget_s: float):
    """The reference's CPU path: fp32 PyTorch-CPU restatement of the Flax forward (oracle/), 1 clip per step."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import videoprism_oracle as O
    _restore_affinity()
    # all the host cores this process may use (torchrun exports OMP_NUM_THREADS=1, which would make this a 1-thread run)
    try:
        n_cores = len(os.sched_getaffinity(0))
    except AttributeError:
        n_cores = os.cpu_count() or 1
    torch.set_num_threads(max(1, n_cores))
    cfg = O.CONFIGS[MODEL_NAME[model]]
    W = O.to_torch(O.make_synthetic_weights(cfg))
    v = torch.from_numpy(O.make_video(1, 16, 288, seed=0))
    times = []
    t_begin = time.perf_counter()
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            O.encoder_forward(cfg, W, v)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
            if time.perf_counter() - t_begin > budget_s and times:
                break
    mean = sum(times) / len(times)
    return 1.0 / mean, len(times), torch.get_num_threads(), mean


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, n_steps, threads, mean = cpu_oracle_clips_per_s(args.model, args.steps, min(args.warmup, 1), budget_s=200.0)
    sample = f"1 clip (1x16x288x288x3) per step, {n_steps} timed steps, mean {mean:.2f} s/clip"
    line = {
        "impl": "reference", "metric": "clips/sec (16x288^2 encoder forward)", "value": value, "unit": "clips/s", "n_gpus": args.gpus,
        "steps": n_steps, "warmup": min(args.warmup, 1), "ms_per_step": mean * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the same workload as the b200 arm (same string, same global batch); each step is a bounded sample of it: one clip
        "config": {"workload": f"{MODEL_NAME[args.model]} encoder forward, 16x288x288x3 clips, random-init weights",
                   "global_batch": args.global_batch, "sample_clips_per_step": 1,
                   "parallelism": f"{threads} host threads (rank 0 only)"},
        "cpu_baseline": {"value": value, "unit": "clips/s", "cores": threads, "kind": "port", "sample": sample,
                         "note": "PyTorch-CPU fp32 restatement of the reference Flax path (jax/flax not installable offline)"},
        "e2e": {"value": value, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_oracle_retrieval(model: str, budget_s: float):
    """CPU baseline of the retrieval workload: the oracle's video-text forward on 1 clip + 4 queries per run."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import videoprism_oracle as O
    _restore_affinity()
    try:
        n_cores = len(os.sched_getaffinity(0))
    except AttributeError:
        n_cores = os.cpu_count() or 1
    torch.set_num_threads(max(1, n_cores))
    name = {"base": "videoprism_lvt_public_v1_base", "large": "videoprism_lvt_public_v1_large"}[model]
    cfg = O.CONFIGS[name]
    W = O.make_synthetic_weights(cfg)
    video = O.make_video(1, 16, 288, seed=0)
    ids, pad = O.make_text(4, vocab=cfg["vocabulary_size"])
    times = []
    t_begin = time.perf_counter()
    for i in range(6):
        t0 = time.perf_counter()
        O.run_clip(cfg, W, video, ids, pad)
        if i >= 1:
            times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_begin > budget_s and times:
            break
    mean = sum(times) / len(times)
    return 1.0 / mean, len(times), torch.get_num_threads(), mean


def run_retrieval(args):
    """BASELINE.json configs[3]/[4]: videoprism_lvt_public_v1_{base,large}; clips and text queries sharded over the
    ranks, pooled embeddings all-gathered (NCCL, one collective), similarity matrix [clips, queries] on every rank."""
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    numa_node = bind_to_gpu_numa_node(local)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=240))   # a mismatched collective aborts in minutes, not in NCCL's default 10
    import videoprism_b200 as vp
    from videoprism_b200.retrieval import gather_embedding_pair, retrieval_similarity, shard_range
    name = {"base": "videoprism_lvt_public_v1_base", "large": "videoprism_lvt_public_v1_large"}[args.model]
    model = vp.get_model(name)
    model.load_state(vp.synthetic_state(model, seed=1234))
    n_clips = args.global_batch if args.global_batch != 32 else 256
    n_q = args.global_queries
    lo, hi = shard_range(n_clips, rank, world)
    qlo, qhi = shard_range(n_q, rank, world)
    rng = np.random.default_rng(100 + rank)
    host_video = torch.from_numpy(rng.random((hi - lo, 16, 288, 288, 3), dtype=np.float32)).pin_memory()
    video = host_video.cuda()
    ids_all = np.random.default_rng(2).integers(1, 32000, (n_q, 64), dtype=np.int32)
    lens = np.random.default_rng(3).integers(4, 33, (n_q,))
    pad_all = (np.arange(64)[None, :] >= lens[:, None]).astype(np.float32)
    ids_all = np.where(pad_all > 0, 0, ids_all).astype(np.int32)
    host_ids = torch.from_numpy(ids_all[qlo:qhi].copy()).pin_memory(); host_pad = torch.from_numpy(pad_all[qlo:qhi].copy()).pin_memory()
    ids = host_ids.cuda(); pad = host_pad.cuda()
    warmup = max(args.warmup, 3)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        return retrieval_similarity(model, video, ids, pad, total_clips=n_clips, total_queries=n_q)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.1)
    t_load0 = time.time()
    for _ in range(warmup):
        sim = step()
    barrier()
    l0 = model.kernel_launches
    t_wall0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        sim = step()
    e1.record()
    barrier()
    clocks = sampler.stop(t_wall0, time.time(), t_load0) if rank == 0 else None
    launches = int(model.kernel_launches - l0)
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / args.steps

    # ---- e2e through the public API with HOST buffers: per step the clips (float32, pinned) go host->device through the
    # chunk-pipelined asynchronous entry point, ids / paddings are copied host->device, the embeddings are exchanged and the
    # similarity matrix is read back to the host.
    e2e = None
    if not args.no_e2e:
        hv = host_video.numpy()
        v_out = [vp.pinned_empty((hi - lo, model.config["model_dim"])) for _ in range(2)]

        def e2e_step(i):
            tk, v_host = model.embed_video_async(hv, out=v_out[i & 1])
            ids_d = host_ids.cuda(non_blocking=True); pad_d = host_pad.cuda(non_blocking=True)
            _, t_d, _ = model(None, ids_d, pad_d)
            model.wait(tk)
            v_d = torch.from_numpy(v_host).cuda(non_blocking=True)
            v_all, t_all = gather_embedding_pair(v_d, t_d, n_clips, n_q)
            return vp.compute_similarity_matrix(v_all, t_all).cpu()

        for i in range(2):
            e2e_step(i)
        barrier()
        n_e2e = max(2, min(args.steps, 5))
        t0 = time.perf_counter()
        for i in range(n_e2e):
            sim_host = e2e_step(i)
        torch.cuda.synchronize()
        tt = torch.tensor([time.perf_counter() - t0], device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        clip_bytes = 16 * 288 * 288 * 3 * 4
        e2e = {"value": n_clips * n_e2e / float(tt.item()), "unit": "clips/s",
               "h2d_bytes_per_step": (hi - lo) * clip_bytes + (qhi - qlo) * 64 * 8 + (hi - lo) * model.config["model_dim"] * 4,
               "d2h_bytes_per_step": (hi - lo) * model.config["model_dim"] * 4 + n_clips * n_q * 4, "steps": n_e2e,
               "queries_per_s": n_q * n_e2e / float(tt.item()),
               "how": "FactorizedVideoCLIP.embed_video_async(numpy pinned clips) -> vp_clip_video_forward_host_async (chunk-pipelined H2D + "
                      "forward), ids / paddings host->device, text forward, wait, all-gather (one collective), similarity matrix read back "
                      "to the host; wall clock, max over ranks"}

    peaks, src = measured_peaks()
    roof, cpu = None, None
    if rank == 0:
        D, F, Hh = model.config["model_dim"], model.config["mlp_dim"], model.config["num_heads"]
        # rank 0 ALONE runs these extra steps, so they must not contain the collective: the forward only (video + text)
        model.trace(True)
        for _ in range(2):
            model(video, ids, pad)
        rows = model.trace_report()
        model.trace(False)
        total_ms = [r for r in rows if r[0] == "TOTAL"][0][2]
        shares = {r[0]: round(r[2] / total_ms, 4) for r in rows if r[0] != "TOTAL"}
        M = (hi - lo) * 16 * 256
        ffn1 = [r for r in rows if r[0] in ("spatial.ffn1", "temporal.ffn1", "aux.ffn1")]
        k_ms = sum(r[2] for r in ffn1) / sum(r[1] for r in ffn1)
        flops = 2.0 * M * F * D
        ach = flops / (k_ms * 1e-3) / 1e12
        aux = [r for r in rows if r[0] == "aux.attn"]
        aux_ms = sum(r[2] for r in aux) / max(1, sum(r[1] for r in aux))
        aux_flops = 4.0 * 4096 * 4096 * 64 * Hh * (hi - lo)
        aux_tf = aux_flops / (aux_ms * 1e-3) / 1e12 if aux_ms > 0 else None
        peak = float(peaks["bf16_tflops_sustained"]); burst = float(peaks.get("bf16_tflops", peak))
        gf = {"base": 1231.04 - 38.71 + 0.15, "large": 3410.92 - 68.80 + 0.27}[args.model]   # pooler in its executed (collapsed) form
        gfq = {"base": 11.20, "large": 19.84}[args.model]
        tf = (n_clips * gf + n_q * gfq) / ms / world   # GF/ms = TF/s per GPU
        roof = {"bound": "tensor", "kernel": f"gemm_bf16_kernel<256,GELU> FFN1 (LayerNorm folded) [{M}x{D}]x[{D}x{F}]", "achieved": ach, "peak": peak,
                "unit": "TFLOP/s", "frac": ach / peak, "frac_of_burst_peak": ach / burst, "traffic": None, "peak_source": src + ", bf16_tflops_sustained",
                "ms_per_launch": k_ms, "flops_per_launch": flops, "share_of_step": round(sum(r[2] for r in ffn1) / total_ms, 4),
                "how": "CUDA events after every launch on the launch stream (vp_trace), 2 extra steps of the same workload",
                "kernel_shares": shares,
                "attention_aux": {"kernel": "attn_kloop_tcgen05_kernel (auxiliary encoder, S = 4096, dh = 64)", "ms_per_launch": aux_ms,
                                  "flops_per_launch": aux_flops, "achieved": aux_tf, "frac": (aux_tf / peak) if aux_tf else None,
                                  "frac_of_burst_peak": (aux_tf / burst) if aux_tf else None},
                "step": {"achieved": tf, "frac": tf / peak,
                         "note": "whole step per GPU; algorithmic GF with the pooling head in its executed single-query form"}}
        if world == 1 and not args.no_cpu_baseline:
            v, n_runs, threads, mean = cpu_oracle_retrieval(args.model, budget_s=25.0)
            cpu = {"value": v, "unit": "clips/s", "cores": threads, "kind": "port",
                   "sample": f"1 clip + 4 text queries per run, 1 warm-up + {n_runs} timed runs (<= 25 s of CPU work), mean {mean:.2f} s; fp32 "
                             "PyTorch-CPU restatement of FactorizedVideoCLIP (oracle/), all host threads torch uses"}
        print(json.dumps({
            "metric": "clips/sec (video-text retrieval step)", "value": n_clips / (ms / 1e3), "unit": "clips/s", "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{name}: {n_clips} clips x {n_q} text queries, all-gather of pooled embeddings, similarity matrix",
                       "clips_per_gpu": hi - lo, "queries_per_gpu": qhi - qlo, "parallelism": f"dp{world} + all-gather (NCCL, one collective per step)",
                       "l2": f"inputs larger than L2: {(hi - lo) * 16 * 288 * 288 * 3 * 4 / 2**20:.0f} MiB of clips per step", "numa_node_rank0": numa_node},
            "queries_per_s": n_q / (ms / 1e3), "similarity_shape": list(sim.shape), "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "roofline": roof, "cpu_baseline": cpu}), flush=True)
    if world > 1:
        dist.barrier()   # rank 0's trace / CPU-baseline legs run alone; nobody leaves before it is done
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="base", choices=["base", "large"])
    ap.add_argument("--global-batch", type=int, default=32)
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: --global-batch clips in total, sharded over the ranks (BASELINE configs[1]); "
                         "weak: --global-batch clips on every rank")
    ap.add_argument("--workload", default="encoder", choices=["encoder", "retrieval"],
                    help="encoder: BASELINE configs 2/3 (default); retrieval: configs 4/5 (LvT video+text, all-gather, similarity)")
    ap.add_argument("--global-queries", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "retrieval":
        return run_retrieval(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    numa_node = bind_to_gpu_numa_node(local)   # before any pinned allocation
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep NCCL's banner off stdout: stdout carries ONE JSON line
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=240))   # a mismatched collective aborts in minutes, not in NCCL's default 10
    warmup = max(args.warmup, 3)

    import videoprism_b200 as vp
    name = MODEL_NAME[args.model]
    model = vp.get_model(name)
    state = vp.synthetic_state(model, seed=1234)
    model.load_state(state)

    if args.scaling == "weak":
        b_local = args.global_batch
        args.global_batch = b_local * world
    else:
        if args.global_batch % world:
            raise SystemExit("--global-batch must be divisible by the number of GPUs")
        b_local = args.global_batch // world
    T, S = 16, 288
    # rotate device input buffers so the clips read by consecutive steps never sit in the 126 MB L2
    clip_bytes = T * S * S * 3 * 4
    n_bufs = max(2, -(-512 * 2**20 // (b_local * clip_bytes)))
    rng = np.random.default_rng(rank)
    host_in = torch.from_numpy(rng.random((b_local, T, S, S, 3), dtype=np.float32)).pin_memory()
    bufs = [host_in.cuda(non_blocking=True) for _ in range(n_bufs)]
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i):
        return model(bufs[i % n_bufs])[0]

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.1)   # let the first nvidia-smi sample land before the load starts
    t_load0 = time.time()
    for i in range(warmup):
        step(i)
    barrier()
    launches0 = model.kernel_launches
    t_wall0 = time.time()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for i in range(args.steps):
        step(i)
    end.record()
    barrier()
    ms = start.elapsed_time(end)
    clocks = sampler.stop(t_wall0, time.time(), t_load0) if rank == 0 else None
    launches = model.kernel_launches - launches0
    t = torch.tensor([ms], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = args.global_batch * args.steps / (ms_total / 1e3)

    # ---- e2e: host buffers through the public API (pinned H2D of the clips + D2H of the features every step)
    e2e = None
    if not args.no_e2e:
        host_np = host_in.numpy()
        dm = model.config["model_dim"]
        out_bytes = b_local * T * 256 * dm * 4
        host_out = [vp.pinned_empty((b_local, T * 256, dm)) for _ in range(2)]
        n_e2e = max(3, min(args.steps, 10))

        def timed(submit, n=n_e2e):
            pipelined_e2e(submit, model.wait, 2)   # warm-up of this variant (workspace, staging buffers)
            barrier()
            tt = torch.tensor([pipelined_e2e(submit, model.wait, n)], device="cuda")
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return args.global_batch * n / float(tt.item())

        # headline: float32 clips in, float32 features out, pinned host buffers, two calls in flight
        v_async = timed(lambda i: model.forward_async(host_np, out=host_out[i & 1])[0])
        e2e = {"value": v_async, "unit": "clips/s", "h2d_bytes_per_step": b_local * clip_bytes, "d2h_bytes_per_step": out_bytes, "steps": n_e2e,
               "how": "models.FactorizedEncoder.forward_async(numpy, out=pinned) + wait -> vp_encoder_forward_host_async / vp_wait: every step "
                      "copies its clips host->device and its features device->host (chunk-pipelined over 3 streams); two steps are in flight, "
                      "so a step's copies overlap its neighbours' forwards; wall clock over all steps until the last result has landed, max over ranks"}
        # the same with one blocking call per step (nothing overlaps across steps): models.FactorizedEncoder.__call__(numpy)
        barrier()
        t0 = time.perf_counter()
        for i in range(n_e2e):
            model(host_np, out=host_out[0])
        torch.cuda.synchronize()
        tt = torch.tensor([time.perf_counter() - t0], device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e["blocking_call"] = {"value": args.global_batch * n_e2e / float(tt.item()), "unit": "clips/s",
                                "h2d_bytes_per_step": b_local * clip_bytes, "d2h_bytes_per_step": out_bytes}
        # uint8 frames in (as decoded; the /255 of video_utils.load_video runs on the device): 4x smaller H2D
        u8 = torch.randint(0, 256, (b_local, T, S, S, 3), dtype=torch.uint8).pin_memory().numpy()
        e2e["uint8_frames"] = {"value": timed(lambda i: model.forward_async(u8, out=host_out[i & 1])[0]), "unit": "clips/s",
                               "h2d_bytes_per_step": b_local * clip_bytes // 4, "d2h_bytes_per_step": out_bytes}
        # uint8 frames in, bfloat16 features out: half the D2H as well
        out16 = [vp.pinned_empty((b_local, T * 256, dm), dtype=np.uint16) for _ in range(2)]
        e2e["uint8_frames_bf16_features"] = {
            "value": timed(lambda i: model.forward_async(u8, out=out16[i & 1], bf16_features=True)[0]), "unit": "clips/s",
            "h2d_bytes_per_step": b_local * clip_bytes // 4, "d2h_bytes_per_step": out_bytes // 2}

        # bytes per second that cross between host memory and the GPUs, all ranks together: at 8 ranks this (one host's
        # memory / PCIe complex), not the forward, is what bounds the float32-in / float32-out variant
        def _host_gbs(d):
            return (d["h2d_bytes_per_step"] + d["d2h_bytes_per_step"]) * world * (d["value"] / args.global_batch) / 1e9
        e2e["host_copy_gbs_all_ranks"] = _host_gbs(e2e)
        for k in ("blocking_call", "uint8_frames", "uint8_frames_bf16_features"):
            e2e[k]["host_copy_gbs_all_ranks"] = _host_gbs(e2e[k])

    # ---- roofline of the dominant kernel (FFN1 GEMM: folded LayerNorm + GELU epilogue), timed IN SITU: the engine
    # records a CUDA event after every launch on the launch stream (vp_trace), a few more steps of the same workload
    # run with that switched on, and the kernel's duration is the event-to-event time averaged over its launches
    # (warm L2, sustained power-capped clocks: the state it has inside the timed region above).
    peaks, peak_src = measured_peaks()
    roof = None
    if rank == 0:
        D, F = model.config["model_dim"], model.config["mlp_dim"]
        M = b_local * T * 256
        n_trace = max(2, min(args.steps, 5))
        model.trace(True)
        for i in range(n_trace):
            step(i)
        rows = model.trace_report()
        model.trace(False)
        total_ms = [r for r in rows if r[0] == "TOTAL"][0][2]
        ffn1 = [r for r in rows if r[0].endswith(".ffn1")]
        k_ms = sum(r[2] for r in ffn1) / sum(r[1] for r in ffn1)
        shares = {r[0]: round(r[2] / total_ms, 4) for r in rows if r[0] != "TOTAL"}
        flops = 2.0 * M * F * D
        ach = flops / (k_ms * 1e-3) / 1e12
        peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
        fwd_tf = value / world * GF_PER_CLIP[args.model] / 1e3
        traffic = None   # dram__bytes_read.sum + dram__bytes_write.sum of this launch from the committed `ncu --set full` capture
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get(f"ffn1_ln_gelu_{M}x{F}x{D}")
        burst = float(peaks.get("bf16_tflops", peak))
        sm_mhz = (clocks or {}).get("sm_mhz")
        attn = [r for r in rows if r[0] == "spatial.attn"]
        attn_ms = sum(r[2] for r in attn) / max(1, sum(r[1] for r in attn))
        attn_flops = 4.0 * 256 * 256 * 64 * model.config["num_heads"] * (b_local * T)
        attn_tf = attn_flops / (attn_ms * 1e-3) / 1e12 if attn_ms > 0 else None
        roof = {"bound": "tensor", "kernel": f"gemm_bf16_kernel<256,GELU> FFN1 (LayerNorm folded) [{M}x{D}]x[{D}x{F}]", "achieved": ach, "peak": peak,
                "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic,
                "traffic_source": "static: the committed `ncu --set full` capture of this launch shape (profiles/ncu_traffic.json), not measured in this run",
                "frac_of_burst_peak": ach / burst,
                "peak_note": ("sampled SM clock above 1.5 GHz: the kernel is not power-capped at this duty, compare with the burst peak "
                              f"({burst:.0f} TF)" if (sm_mhz or 0) > 1500 else "SM clock power-capped: the sustained peak is the denominator"),
                "attention": {"kernel": "attn_kloop_tcgen05_kernel (spatial, S = 256, dh = 64)", "ms_per_launch": attn_ms, "flops_per_launch": attn_flops,
                              "achieved": attn_tf, "frac": (attn_tf / peak) if attn_tf else None, "frac_of_burst_peak": (attn_tf / burst) if attn_tf else None,
                              "note": "4*S^2*dh flop per head-sequence; exponentials and the cap are not counted"},
                "peak_source": peak_src + ", bf16_tflops_sustained",
                "ms_per_launch": k_ms, "flops_per_launch": flops, "launches_timed": sum(r[1] for r in ffn1),
                "share_of_step": round(sum(r[2] for r in ffn1) / total_ms, 4),
                "how": "CUDA events after every launch on the launch stream (vp_trace), averaged over the kernel's launches in "
                       f"{n_trace} extra steps of the same workload",
                "kernel_shares": shares,
                "forward": {"achieved": fwd_tf, "frac": fwd_tf / peak, "gflop_per_clip": GF_PER_CLIP[args.model],
                            "note": "whole forward per GPU = clips/s/GPU x algorithmic GF/clip"}}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, n_steps, threads, mean = cpu_oracle_clips_per_s(args.model, 10, 1, budget_s=25.0)
        cpu = {"value": v, "unit": "clips/s", "cores": threads, "kind": "port",
               "sample": f"1 clip (1x16x288x288x3) per run, 1 warm-up + {n_steps} timed runs (<= 25 s of CPU work), mean {mean:.2f} s; "
                         "fp32 PyTorch-CPU restatement of the Flax path (oracle/), all host threads torch uses"}

    if rank == 0:
        line = {
            "metric": "clips/sec (16x288^2 encoder forward)", "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{name} encoder forward, 16x288x288x3 clips, random-init weights", "global_batch": args.global_batch,
                       "clips_per_gpu": b_local, "parallelism": f"dp{world} (batch shard, no collective)",
                       "l2": f"inputs larger than L2: {n_bufs} rotating device input buffers of {b_local * clip_bytes / 2**20:.0f} MiB",
                       "numa_node_rank0": numa_node},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
