#!/usr/bin/env python
"""Benchmark of the DGCNN EdgeConv hot path: point clouds / s for a DGCNNSeg training step
(forward + cross-entropy + backward + gradient all-reduce + Adam), N = 2048 points, k = 20.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29500 bench.py --gpus 8 --steps 20 --warmup 5
    python bench.py --impl reference --steps 5 --warmup 2      # the reference's own CPU path
    python bench.py --workload configC | static40 | chamfer | pointtransformer | inference   # BASELINE configs[2..4], f4

Prints ONE JSON line (rank 0). See DESIGN.md section "Measurement" for the definitions.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn.functional as F  # noqa: E402

METRIC = "point clouds/sec DGCNNSeg fwd+bwd N=2048 k=20"
UNIT = "clouds/s"
NCU_METRICS = os.path.join(ROOT, "profiles", "r02_ncu_metrics.json")     # written by tools/ncu_metrics.py from a capture


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train",
                    choices=["train", "configC", "static40", "chamfer", "pointtransformer", "inference"])
    ap.add_argument("--batch", type=int, default=None, help="clouds per GPU (default: per workload)")
    ap.add_argument("--points", type=int, default=None)
    ap.add_argument("--k", type=int, default=None)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--static", action="store_true", help="static coordinate graph (train.py --static)")
    ap.add_argument("--cpu-batch", type=int, default=None, help="clouds per CPU step (default: 8 for cpu_baseline, "
                    "the GPU batch for --impl reference when the host has the memory)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="do not capture the step in a CUDA graph")
    args = ap.parse_args()
    wl = {"train": (32, 2048, 20, 3), "configC": (8, 8192, 40, 9), "static40": (32, 2048, 40, 3),
          "chamfer": (64, 2048, 0, 3), "pointtransformer": (4, 4096, 16, 3), "inference": (1, 20000, 20, 3)}[args.workload]
    args.batch = args.batch or wl[0]
    args.points = args.points or wl[1]
    args.k = args.k or wl[2]
    args.in_features = wl[3]
    if args.workload in ("configC", "static40"):
        args.static = True          # the configuration the authors trained with k = 40 (bash_scripts/redo_dgcnn_seg.sh:6-8)
    return args


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p.get("bf16_tflops", 1590.0)), "measured"
    return 6650.0, 1590.0, "fallback"


def from_profile(kernel_substr):
    """Counter values of one kernel from the committed ncu capture of this round (profiles/r02_ncu_metrics.json,
    written by tools/ncu_metrics.py with the capture's file name and git revision). None when no capture exists: the
    bench line never carries a hand-copied constant."""
    if not os.path.exists(NCU_METRICS):
        return None
    with open(NCU_METRICS) as f:
        data = json.load(f)
    for name, rec in data.get("kernels", {}).items():
        if kernel_substr in name:
            out = dict(rec)
            out["kernel"] = name
            out["source"] = {"file": os.path.relpath(NCU_METRICS, ROOT), "git": data.get("git"), "capture": data.get("capture")}
            return out
    return None


# ------------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [s.strip() for s in line.split(",")]
            if len(parts) >= 6:
                self.samples.append(parts)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4) if s[2 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# the reference itself (baseline/_ref or /root/reference through oracle/reference_shim.py), else the oracle port
# ------------------------------------------------------------------------------------------------
def build_reference_model(args, device):
    """(step_fn(x, y) -> loss, kind): the UNMODIFIED reference DGCNNSeg with Adam + cross-entropy exactly as
    model_trainer.py:57, 154-195 drives it; falls back to the pinned oracle port when the staged files are absent."""
    from oracle import reference_shim
    torch.manual_seed(0)
    if reference_shim.available():
        ref_dgcnn, _, _ = reference_shim.load()
        model = ref_dgcnn.DGCNNSeg(k=args.k, in_features=args.in_features, num_classes=4, dynamic=not args.static).to(device)
        model.train()
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-5)

        def fwd(x):
            return model(x)
        return fwd, opt, "reference"
    from oracle import dgcnn_oracle as O
    p = O.make_params(O.dgcnn_seg_param_shapes(args.in_features, 4), 1, random_bn=False)
    p = {n: v.to(device) for n, v in p.items()}
    params = {n: v.clone().requires_grad_(True) for n, v in p.items() if v.dtype.is_floating_point and "running" not in n}
    state = {n: v for n, v in p.items() if n not in params}
    opt = torch.optim.Adam(list(params.values()), lr=1e-3, weight_decay=1e-5)

    def fwd(x):
        stats = {}
        out = O.dgcnn_seg({**state, **params}, x, args.k, dynamic=not args.static, training=True, stats_out=stats)
        state.update(stats)
        return out
    return fwd, opt, "port"


def host_memory_gb():
    try:
        import psutil
        return psutil.virtual_memory().available / 2 ** 30
    except Exception:
        return 0.0


def cpu_reference_steps(args, batch, steps, warmup, max_seconds):
    """DGCNNSeg forward + CE + backward + Adam on the host cores, `batch` clouds per step, fp32, autocast off
    (model_trainer.py:76 disables it on the CPU). Returns (clouds/s, s/step, threads, n_steps, kind)."""
    from fissure_segmentation_b200 import synth
    if torch.get_num_threads() < (os.cpu_count() or 1):
        torch.set_num_threads(os.cpu_count())          # torchrun exports OMP_NUM_THREADS=1: use the whole host anyway
    x, y = synth.make_batch(batch, args.points, seed=1234, n_features=args.in_features - 3)
    fwd, opt, kind = build_reference_model(args, "cpu")
    times = []
    t_begin = time.perf_counter()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        loss = F.cross_entropy(fwd(x), y)
        loss.backward()
        opt.step()
        opt.zero_grad()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if time.perf_counter() - t_begin > max_seconds and len(times) >= 2:
            break
    total = sum(times)
    return batch * len(times) / total, total / len(times), torch.get_num_threads(), len(times), kind


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path, all host threads, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload not in ("train", "configC", "static40"):
        print(json.dumps({"impl": "reference", "unavailable": "workload %s has no CPU reference arm (pytorch3d / "
                          "pointops_cuda are CUDA-only third-party dependencies)" % args.workload}), flush=True)
        return
    batch = args.cpu_batch
    if batch is None:
        # the GPU arm's per-step batch when the host can hold the reference's activations (~0.25 GB per cloud at
        # N=2048, k=20, fp32), else a smaller bounded sample
        need_gb = 0.25 * args.batch * (args.points / 2048.0) ** 2 * max(args.k, 20) / 20.0
        batch = args.batch if host_memory_gb() > 2.0 * need_gb + 8.0 else max(2, min(args.batch, 4))
    val, spstep, threads, n, kind = cpu_reference_steps(args, batch, args.steps, args.warmup, max_seconds=200.0)
    sample = ("%d clouds/step x %d timed steps of the N=%d k=%d DGCNNSeg training step on the host CPU (%s, fp32, "
              "autocast off)" % (batch, n, args.points, args.k,
                                 "unmodified reference module from baseline/_ref" if kind == "reference" else "oracle port"))
    cfg = workload_config(args, 1)
    cfg["cpu_step_batch"] = batch
    cfg["same_batch_as_gpu_arm"] = batch == args.batch
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": n,
        "warmup": args.warmup, "ms_per_step": spstep * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample,
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def gpu_eager_baseline(args, dev, steps=8, warmup=3):
    """The reference module in PyTorch eager on the same B200 (the practical bar): fp32 and fp16 autocast +
    GradScaler exactly like model_trainer.py:75-76, 154-195, same batch and shape as the measured step."""
    from fissure_segmentation_b200 import synth
    x, y = synth.make_batch(args.batch, args.points, seed=1234, n_features=args.in_features - 3)
    x, y = x.to(dev), y.to(dev)
    out = {}
    for mode in ("fp32", "fp16_autocast"):
        try:
            fwd, opt, kind = build_reference_model(args, dev)
            scaler = torch.amp.GradScaler("cuda", enabled=mode != "fp32")
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            for i in range(warmup + steps):
                if i == warmup:
                    torch.cuda.synchronize()
                    ev[0].record()
                with torch.autocast("cuda", dtype=torch.float16, enabled=mode != "fp32"):
                    loss = F.cross_entropy(fwd(x), y)
                scaler.scale(loss).backward()
                scaler.step(opt)
                scaler.update()
                opt.zero_grad()
            ev[1].record()
            torch.cuda.synchronize()
            ms = ev[0].elapsed_time(ev[1]) / steps
            out[mode] = {"value": args.batch / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "kind": kind,
                         "steps": steps, "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30}
        except Exception as exc:           # e.g. out of memory at a large shape: report, do not fail the bench
            out[mode] = {"unavailable": repr(exc)[:200]}
        torch.cuda.empty_cache()
    return out


def workload_config(args, world):
    names = {
        "train": "DGCNNSeg(k=%d,in_features=3,num_classes=4,%s) train step, batch %d/GPU, N=%d, CE loss, Adam",
        "configC": "DGCNNSeg(k=%d,in_features=9,num_classes=4,%s) train step (large-cloud kNN stress), batch %d/GPU, N=%d, CE loss, Adam",
        "static40": "DGCNNSeg(k=%d,in_features=3,num_classes=4,%s) train step (the authors' k=40 static setting), batch %d/GPU, N=%d, CE loss, Adam",
    }
    if args.workload in names:
        wl = names[args.workload] % (args.k, "static" if args.static else "dynamic", args.batch, args.points)
    elif args.workload == "chamfer":
        wl = "ChamferLoss forward+backward, %d cloud pairs/GPU, %d vs %d points (PC-AE / DG-SSM loss)" % (args.batch, args.points, args.points)
    elif args.workload == "pointtransformer":
        wl = ("reference PointTransformerCompatibility(in_features=3,num_classes=4) fwd+bwd through the reference's "
              "pointops.py on this package's pointops_cuda, batch %d/GPU, N=%d" % (args.batch, args.points))
    else:
        wl = ("predict_full_pointcloud: %d keypoints, 50 eval runs on %d-point subsets (point_seg_net.py:21-48), batched + "
              "CUDA graph" % (args.points, 2048))
    return {"workload": wl, "global_batch": args.batch * world, "points": args.points, "k": args.k,
            "parallelism": "dp%d" % world,
            "l2": "no explicit flush: each step streams > 1 GB of activations through the 126 MB L2 and rotates "
                  "through 4 distinct input batches"}


# ------------------------------------------------------------------------------------------------
# roofline bookkeeping: algorithmic work per call of the timed entry points
# ------------------------------------------------------------------------------------------------
def fma_peak_tflops(dev):
    """Measured FP32 FMA throughput (fs_fma_microbench), best of 5 launches."""
    from fissure_segmentation_b200 import _lib
    out = torch.zeros(1, device=dev)
    iters = 4096
    flops = 2.0 * 64 * iters * 256 * 148 * 8
    best = None
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.call("fs_fma_microbench", out, iters, out, None)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return flops / (best * 1e-3) / 1e12


TMEM_READ_BYTES_PER_CLK_PER_SM = 64      # tcgen05.ld: "LDTM throughput: TMEM-read 64 B/cyc" (B300_MICROARCH.md, TMEM table)


def entry_rooflines(args, timed_ms, hbm_peak, tensor_peak, fma_peak, peak_kind, sm_mhz=None):
    """One roofline object per timed entry point: algorithmic bytes / flops per call divided by the mean
    CUDA-event duration of its calls in eager steps of this workload."""
    B, N, k = args.batch, args.points, args.k
    P = B * N
    cp = 64
    defs = {
        # feature-space kNN (ec2/ec3 graphs): whole entry point (operand prep + tcgen05 sweeps + finalize)
        "fs_knn_feat_tc": ("tensor", 2.0 * B * N * N * 64, tensor_peak, "TFLOP/s",
                           "fs_knn_feat_tc entry point (operand prep + knn_tc_select [tcgen05] + finalize), C=64"),
        # 3-D kNN: (2C+2) = 8 flops per pair, against the measured FP32 FMA rate
        "fs_knn3d": ("fp32_fma", 8.0 * B * N * N, fma_peak, "TFLOP/s", "fs_knn3d (coordinate kNN, C=3, SIMT)"),
        # the same through the tensor-core kernels: the distances come out of ONE K = 16 tcgen05 step (hi/lo split of the
        # three coordinates + norms), the CUDA cores only select; reported against the FP32 FMA rate of the SIMT form
        "fs_knn3d_tc": ("fp32_fma", 8.0 * B * N * N, fma_peak, "TFLOP/s",
                        "fs_knn3d_tc entry point (coordinate kNN, C=3: operand prep + knn_tc_select [tcgen05] + finalize)"),
        # EdgeConv gather/max pass, fp32 table: [a|b] 2*Cp*4 + idx 4k + sel 4Cp + arg Cp + sum_y 4Cp bytes per point
        "fs_edgeconv_gather": ("hbm", float(P) * (2 * cp * 4 + 4 * k + 4 * cp + cp + 4 * cp), hbm_peak, "GB/s",
                               "fs_edgeconv_gather (ec2/ec3 gather/max pass, Cp=64, train)"),
        # fused two-layer EdgeConv on coordinates (ec1): layer-2 contraction 2*P*k*64*64 flops on the tensor pipe
        "fs_edge2_fwd": ("tensor", 2.0 * P * k * 64 * 64, tensor_peak, "TFLOP/s",
                         "fs_edge2_fwd (fused two-layer EdgeConv forward, tcgen05 64x64 contraction per edge)"),
        "fs_edge2_bwd": ("tensor", 3 * 2.0 * P * k * 64 * 64, tensor_peak, "TFLOP/s",
                         "fs_edge2_bwd (fused two-layer EdgeConv backward: recompute + dW2 + dH contractions)"),
    }
    # Every accumulator value a tcgen05 kernel looks at leaves TMEM through tcgen05.ld (64 B/clk/SM): for the kernels
    # whose epilogue must see EVERY product (kNN: two sweeps over all N^2 scores; pooled GEMM; per-edge max) that read
    # rate - not the tensor pipe - is the floor. Bytes per launch of the dominant kernel of the entry point:
    tmem_bytes = {"fs_knn_feat_tc": 2.0 * B * N * N * 4, "fs_knn3d_tc": 2.0 * B * N * N * 4,
                  "fs_edge2_fwd": float(P) * k * 64 * 4, "fs_pool_gemm": float(P) * 1024 * 4}
    defs["fs_pool_gemm"] = ("tensor", 2.0 * P * 1024 * 192, tensor_peak, "TFLOP/s",
                            "fs_pool_gemm (global feature 192 -> 1024: tcgen05 product + per-cloud max/arg epilogue, "
                            "product never written)")
    out = {}
    for name, ms_list in timed_ms.items():
        if name not in defs or not ms_list:
            continue
        bound, work, peak, unit, label = defs[name]
        ms = statistics.mean(ms_list)
        calls_per_step = len(ms_list) / 3.0
        achieved = work / (ms * 1e-3) / (1e9 if unit == "GB/s" else 1e12)
        out[name] = {"kernel": label, "bound": bound, "achieved": achieved, "peak": peak, "peak_kind": peak_kind if bound != "fp32_fma" else "measured (fs_fma_microbench, this run)",
                     "unit": unit, "frac": achieved / peak if peak else None, "launch_ms": ms,
                     "calls_per_step": calls_per_step, "ms_per_step": ms * calls_per_step,
                     ("algorithmic_bytes_per_launch" if unit == "GB/s" else "algorithmic_flops_per_launch"): work,
                     "launches_timed": len(ms_list), "traffic": None}
        if name in tmem_bytes and sm_mhz:
            floor_us = tmem_bytes[name] / (TMEM_READ_BYTES_PER_CLK_PER_SM * 148 * sm_mhz * 1e6) * 1e6
            out[name]["tmem_read"] = {"bytes_per_launch": tmem_bytes[name],
                                      "peak_bytes_per_clk_per_sm": TMEM_READ_BYTES_PER_CLK_PER_SM,
                                      "sm_mhz": sm_mhz, "floor_us": floor_us, "frac_of_floor": floor_us / (ms * 1e3),
                                      "note": "floor of the accumulator read-out alone, against the whole entry point's time"}
    return out


# ------------------------------------------------------------------------------------------------
# our arm: DGCNNSeg training step (workloads train / configC / static40)
# ------------------------------------------------------------------------------------------------
def run_train(args):
    import fissure_segmentation_b200 as fs
    from fissure_segmentation_b200 import _lib, synth
    from fissure_segmentation_b200.ddp import FlatAdam, FlatDataParallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True

    torch.manual_seed(0)
    model = fs.DGCNNSeg(k=args.k, in_features=args.in_features, num_classes=4, dynamic=not args.static).to(dev)
    model.precision = args.precision
    model.train()
    # the tail bucket = the EdgeConv gradients (3 % of the parameters): its all-reduce is the exposed one
    # FS_FUSED_TAIL: all (default: every gradient reduced from peer memory inside the Adam kernel, no NCCL call in the
    # step) | 1 (EdgeConv bucket from peer memory, head bucket by NCCL during backward) | 0 (NCCL only)
    ft = os.environ.get("FS_FUSED_TAIL", "all")
    dp = FlatDataParallel(model, n_buckets=int(os.environ.get("FS_DDP_BUCKETS", "2")), tail_share=0.04,
                          fused_tail="all" if ft == "all" else ft != "0")
    dp.skip_collectives = os.environ.get("FS_DDP_SKIP_COLLECTIVES", "0") == "1"      # diagnosis only
    opt = FlatAdam(dp, lr=1e-3, weight_decay=1e-5)

    # synthetic lung-keypoint clouds: 4 distinct batches per rank, pinned on the host
    n_pool = 4
    pool_h = [synth.make_batch(args.batch, args.points, seed=1234 + rank * 100 + i, n_features=args.in_features - 3)
              for i in range(n_pool)]
    pool_h = [(x.pin_memory(), y.pin_memory()) for x, y in pool_h]
    pool_d = [(x.to(dev), y.to(dev)) for x, y in pool_h]
    x_in = torch.empty_like(pool_d[0][0])
    y_in = torch.empty_like(pool_d[0][1])
    loss_h = torch.zeros(1).pin_memory()
    loss_d = torch.zeros(1, device=dev)

    def train_step(x, y):
        dp.zero_grad()
        logits = dp(x)
        loss = F.cross_entropy(logits, y)
        loss.backward()
        dp.finish_backward()
        opt.step()
        loss_d.copy_(loss.detach().reshape(1))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (eager), then capture the whole step (forward, loss, backward, all-reduce, Adam) in a CUDA graph
    for i in range(max(args.warmup, 3)):
        train_step(*pool_d[i % n_pool])
    barrier()
    graph = None
    launches_per_step = None
    if not args.no_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                x_in.copy_(pool_d[0][0]); y_in.copy_(pool_d[0][1])
                train_step(x_in, y_in)
            torch.cuda.current_stream().wait_stream(side)
            barrier()
            c0 = _lib.launch_count
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                train_step(x_in, y_in)
            launches_per_step = _lib.launch_count - c0
            graph = g
        except Exception as exc:      # capture not possible (e.g. a collective that cannot be captured): eager
            if rank == 0:
                print("# CUDA graph capture failed, running eagerly: %r" % (exc,), file=sys.stderr)
            graph = None
            torch.cuda.synchronize()

    def run_step(x, y):
        if graph is None:
            train_step(x, y)
        else:
            x_in.copy_(x, non_blocking=True)
            y_in.copy_(y, non_blocking=True)
            graph.replay()

    for i in range(args.warmup):
        run_step(*pool_d[i % n_pool])
    barrier()

    # ---- device-resident throughput -------------------------------------------------------------
    launches0 = _lib.launch_count
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        run_step(*pool_d[i % n_pool])
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = (launches_per_step * args.steps) if graph is not None else (_lib.launch_count - launches0)
    ms_total = ev0.elapsed_time(ev1)

    # ---- end to end: pinned host inputs -> H2D -> step -> loss D2H, every step ------------------
    def e2e_step(i):
        xh, yh = pool_h[i % n_pool]
        x_in.copy_(xh, non_blocking=True)
        y_in.copy_(yh, non_blocking=True)
        if graph is None:
            train_step(x_in, y_in)
        else:
            graph.replay()
        loss_h.copy_(loss_d, non_blocking=True)
        torch.cuda.current_stream().synchronize()      # the trainer reads the loss value each step
        return float(loss_h[0])

    for i in range(2):
        e2e_step(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        e2e_step(i)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)

    # ---- per-call duration of the roofline entry points: CUDA events around their calls in 3 eager steps of the same
    #      workload (events cannot be read back from inside a replayed graph)
    timed_names = {"fs_edgeconv_gather", "fs_knn_feat_tc", "fs_knn3d", "fs_knn3d_tc", "fs_edge2_fwd", "fs_edge2_bwd",
                   "fs_pool_gemm"}
    _lib.time_calls.update(timed_names)
    _lib.timed.clear()
    for i in range(3):
        train_step(*pool_d[i % n_pool])
    barrier()
    _lib.time_calls.clear()
    timed_ms = {n: [a.elapsed_time(b) for a, b in v] for n, v in _lib.timed.items()}
    _lib.timed.clear()

    t = torch.tensor([ms_total, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = float(t[0]), float(t[1])
    clouds = args.batch * world * args.steps

    if rank == 0:
        hbm_peak, tensor_peak, peak_kind = peaks()
        fma_peak = fma_peak_tflops(dev)
        roofs = entry_rooflines(args, timed_ms, hbm_peak, tensor_peak, fma_peak, peak_kind,
                                sm_mhz=(clocks or {}).get("sm_mhz"))
        prof = {"fs_knn_feat_tc": from_profile("knn_tc_select"), "fs_edgeconv_gather": from_profile("edgeconv_gather"),
                "fs_edge2_fwd": from_profile("edge2_fwd"), "fs_edge2_bwd": from_profile("edge2_bwd"),
                "fs_knn3d": from_profile("knn3d"), "fs_knn3d_tc": from_profile("knn_tc_select")}
        for name, r in roofs.items():
            p = prof.get(name)
            if p is not None:
                r["traffic"] = p.get("dram_bytes")
                r["from_profile"] = p
        metric = METRIC if args.workload == "train" else "point clouds/sec DGCNNSeg fwd+bwd N=%d k=%d" % (args.points, args.k)
        line = {
            "metric": metric, "value": clouds / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision if args.precision != "fp32" else "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "e2e": {"value": clouds / (ms_e2e * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(x_in.numel() * 4 + y_in.numel() * 8), "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches), "cuda_graph": graph is not None,
            "clocks": clocks, "fp32_fma_peak_tflops_measured": fma_peak,
        }
        # `roofline` = the entry point that takes the largest share of the step among this library's timed kernels;
        # the others follow as roofline_<entry point>
        if roofs:
            top = max(roofs, key=lambda n: roofs[n]["ms_per_step"])
            line["roofline"] = roofs[top]
            for name, r in roofs.items():
                if name != top:
                    line["roofline_" + name[3:]] = r
        if world == 1 and not args.no_cpu_baseline:
            cb = args.cpu_batch or (8 if args.points <= 2048 else 2)
            val, spstep, threads, n, kind = cpu_reference_steps(args, cb, 6, 1, max_seconds=30.0)
            line["cpu_baseline"] = {
                "value": val, "unit": UNIT, "cores": threads, "kind": kind, "host_cpus": os.cpu_count(),
                "sample": "%d clouds/step x %d steps of the same N=%d k=%d training step (%s, fp32)"
                          % (cb, n, args.points, args.k, "unmodified reference module, baseline/_ref" if kind == "reference"
                             else "oracle port of the reference's PyTorch CPU path")}
        if world == 1 and not args.no_eager_baseline:
            line["gpu_eager_baseline"] = gpu_eager_baseline(args, dev)
        print(json.dumps(line), flush=True)
    graph = g = None        # noqa: F841  (closure cells of run_step / e2e_step release the CUDA graph)
    teardown(world)


def teardown(world):
    """Ordered tear-down of a multi-rank run. The caller has dropped the captured graph (it holds NCCL work of the
    communicator); here: collect it, drain the device, meet the other ranks, then destroy the process group while every
    rank is still alive. Round 1 saw communicator destruction dead-lock while a captured graph was alive; a watchdog
    turns any repeat of that into a plain exit (the result line is already printed and flushed) instead of a hang."""
    if world <= 1:
        return
    import gc
    gc.collect()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    watchdog = threading.Timer(20.0, lambda: os._exit(0))
    watchdog.daemon = True
    watchdog.start()
    dist.destroy_process_group()
    watchdog.cancel()


# ------------------------------------------------------------------------------------------------
# other workloads (BASELINE configs[3], [4]): one JSON line each, same timing rules
# ------------------------------------------------------------------------------------------------
def _time_loop(fn, steps, warmup, world):
    for _ in range(max(warmup, 3)):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    return ms


def run_other(args):
    import fissure_segmentation_b200 as fs
    from fissure_segmentation_b200 import _lib, synth
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    hbm_peak, tensor_peak, peak_kind = peaks()
    sampler = ClockSampler(local)
    extra = {}
    steps = min(args.steps, 100)
    if args.workload == "chamfer":
        B, N = args.batch, args.points
        a, b = synth.make_chamfer_pair(B, N, seed=4 + rank)
        a = a.to(dev).requires_grad_(True)
        b = b.to(dev)
        loss_fn = fs.ChamferLoss()

        def fn():
            a.grad = None
            loss_fn(a, b).backward()
        sampler.start()
        ms = _time_loop(fn, steps, args.warmup, world)
        unit, value, metric = "cloud pairs/s", B * world * steps / (ms * 1e-3), "Chamfer loss fwd+bwd, %d vs %d points" % (N, N)
        fma_peak = fma_peak_tflops(dev)
        _lib.time_calls.add("fs_nn_points")
        _lib.timed.clear()
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        _lib.time_calls.clear()
        nn_ms = [x.elapsed_time(y) for x, y in _lib.timed.get("fs_nn_points", [])]
        flops = 8.0 * B * N * N
        t = statistics.mean(nn_ms)
        extra["roofline"] = {"kernel": "nn_points_kernel (one direction of the Chamfer nearest-neighbour reduction)",
                             "bound": "fp32_fma", "achieved": flops / (t * 1e-3) / 1e12, "peak": fma_peak,
                             "peak_kind": "measured (fs_fma_microbench, this run)", "unit": "TFLOP/s",
                             "frac": flops / (t * 1e-3) / 1e12 / fma_peak, "launch_ms": t, "traffic": None,
                             "algorithmic_flops_per_launch": flops,
                             "hbm_frac_secondary": B * 2 * N * 20.0 / (t * 1e-3) / 1e9 / hbm_peak}
    elif args.workload == "pointtransformer":
        from oracle import reference_shim            # the reference's own module drives our pointops_cuda kernels
        _, ref_seg = reference_shim.load_pointtransformer()
        from fissure_segmentation_b200 import pointops_cuda as pc
        torch.manual_seed(0)
        model = ref_seg.PointTransformerCompatibility(in_features=3, num_classes=4).to(dev).train()
        x, y = synth.make_batch(args.batch, args.points, seed=1234 + rank)
        x, y = x.to(dev), y.to(dev)
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)

        def fn():
            opt.zero_grad()
            F.cross_entropy(model(x), y).backward()
            opt.step()
        sampler.start()
        pc.clear_knn_cache()
        ms = _time_loop(fn, min(steps, 20), args.warmup, world)
        steps = min(steps, 20)
        extra["knn_cache"] = dict(pc.knn_cache_stats)
        unit, value, metric = UNIT, args.batch * world * steps / (ms * 1e-3), "point clouds/sec PointTransformer seg fwd+bwd N=%d" % args.points
    else:   # inference
        torch.manual_seed(0)
        model = fs.DGCNNSeg(k=args.k, in_features=3, num_classes=4).to(dev).eval()
        model.precision = args.precision
        pc_full, _ = synth.make_batch(1, args.points, seed=99)
        pc_full = pc_full.to(dev)
        import contextlib
        import io

        def fn():
            with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
                model.predict_full_pointcloud(pc_full, sample_points=2048, n_runs_min=50)
        sampler.start()
        ms = _time_loop(fn, min(steps, 20), args.warmup, world)
        steps = min(steps, 20)
        model.inference_cuda_graph = False
        import fissure_segmentation_b200.modelio as mio
        seq_ms = None
        try:
            batched = mio.PointSegmentationModelBase.predict_full_pointcloud

            def seq():
                # the reference's sequential loop: 50 B=1 forwards (forced by a batch-of-one per call)
                with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
                    acc = torch.zeros(1, 4, args.points, device=dev)
                    for _ in range(50):
                        sub = torch.randperm(args.points, device=dev)[:2048]
                        acc[..., sub] += torch.softmax(model(pc_full[..., sub]).float(), dim=1)
            seq_ms = _time_loop(seq, 5, 2, 1) / 5
            del batched
        except Exception as exc:
            extra["sequential_error"] = repr(exc)[:200]
        extra["sequential_loop_ms_per_cloud"] = seq_ms
        extra["note"] = ("performance_time_plot.py:31 quotes 0.0009 s for the reference's 'net' stage on other hardware and an "
                         "unstated configuration; not comparable")
        unit, value, metric = "full clouds/s", world * steps / (ms * 1e-3), "predict_full_pointcloud, %d keypoints, 50 runs x 2048" % args.points
    clocks = sampler.stop()
    if rank == 0:
        line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": steps, "warmup": args.warmup,
                "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32" if args.workload != "inference" else args.precision, "data": "synthetic",
                "config": workload_config(args, world), "gpu_launches": int(_lib.launch_count), "clocks": clocks}
        line.update(extra)
        print(json.dumps(line), flush=True)
    teardown(world)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload in ("train", "configC", "static40"):
        run_train(args)
    else:
        run_other(args)


if __name__ == "__main__":
    main()
