#!/usr/bin/env python
"""Benchmark of the DGCNN EdgeConv hot path: point clouds / s for a DGCNNSeg training step
(forward + cross-entropy + backward + gradient all-reduce + Adam), N = 2048 points, k = 20.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29500 bench.py --gpus 8 --steps 20 --warmup 5
    python bench.py --impl reference --steps 5 --warmup 2      # the reference's CPU path (oracle port)

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
GATHER_DRAM_BYTES = 41_340_416      # ncu --set full, one launch, B=32 N=2048 k=20 (profiles/r01_g_gather_smem_full.txt)


def _tensor_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f).get("bf16_tflops", 1626.3))
    return 1626.3


TENSOR_PEAK_TFLOPS = _tensor_peak()


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="clouds per GPU")
    ap.add_argument("--points", type=int, default=2048)
    ap.add_argument("--k", type=int, default=20)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--static", action="store_true", help="static coordinate graph (train.py --static)")
    ap.add_argument("--cpu-batch", type=int, default=2, help="clouds per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="do not capture the step in a CUDA graph")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


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
# reference arm / CPU baseline: the oracle's restatement of the reference's PyTorch CPU path
# ------------------------------------------------------------------------------------------------
def cpu_reference_steps(args, steps, warmup, max_seconds=120.0):
    """DGCNNSeg forward + CE + backward + Adam on the host cores, `cpu_batch` clouds per step, fp32,
    autocast off (model_trainer.py:76 disables it on CPU). Returns (clouds/s, s/step, threads, n_steps)."""
    from fissure_segmentation_b200 import synth
    from oracle import dgcnn_oracle as O
    torch.manual_seed(0)
    B = args.cpu_batch
    x, y = synth.make_batch(B, args.points, seed=1234)
    p = O.make_params(O.dgcnn_seg_param_shapes(3, 4), 1, random_bn=False)
    params = {n: v.clone().requires_grad_(True) for n, v in p.items() if v.dtype.is_floating_point and "running" not in n}
    state = {n: v for n, v in p.items() if n not in params}
    opt = torch.optim.Adam(list(params.values()), lr=1e-3, weight_decay=1e-5)
    times = []
    t_begin = time.perf_counter()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        stats = {}
        logits = O.dgcnn_seg({**state, **params}, x, args.k, dynamic=not args.static, training=True, stats_out=stats)
        loss = F.cross_entropy(logits, y)
        loss.backward()
        opt.step()
        opt.zero_grad()
        state.update(stats)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if time.perf_counter() - t_begin > max_seconds and len(times) >= 2:
            break
    total = sum(times)
    return B * len(times) / total, total / len(times), torch.get_num_threads(), len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    val, spstep, threads, n = cpu_reference_steps(args, args.steps, args.warmup, max_seconds=240.0)
    sample = "%d clouds/step x %d steps of the N=%d k=%d workload on the host CPU" % (args.cpu_batch, n, args.points, args.k)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": n,
        "warmup": args.warmup, "ms_per_step": spstep * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {"workload": "DGCNNSeg(k=%d,in_features=3,num_classes=4,%s) train step, batch %d/GPU, N=%d, CE loss, Adam"
                        % (args.k, "static" if args.static else "dynamic", args.batch, args.points),
            "global_batch": args.batch * world, "points": args.points, "k": args.k,
            "parallelism": "dp%d" % world,
            "l2": "no explicit flush: each step streams > 1 GB of activations through the 126 MB L2 and rotates "
                  "through 4 distinct input batches"}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
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
    model = fs.DGCNNSeg(k=args.k, in_features=3, num_classes=4, dynamic=not args.static).to(dev)
    model.precision = args.precision
    model.train()
    dp = FlatDataParallel(model, n_buckets=2)
    opt = FlatAdam(dp, lr=1e-3, weight_decay=1e-5)

    # synthetic lung-keypoint clouds: 4 distinct batches per rank, pinned on the host
    n_pool = 4
    pool_h = [synth.make_batch(args.batch, args.points, seed=1234 + rank * 100 + i) for i in range(n_pool)]
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

    # ---- per-launch duration of the roofline kernels: CUDA events around their launches in eager steps of the same
    #      workload (events cannot be read back from inside a replayed graph)
    _lib.time_calls.update({"fs_edgeconv_gather", "fs_knn_feat_tc"})
    _lib.timed.clear()
    for i in range(3):
        train_step(*pool_d[i % n_pool])
    barrier()
    _lib.time_calls.clear()
    gather_ms = [a.elapsed_time(b) for a, b in _lib.timed.get("fs_edgeconv_gather", [])]
    knn_ms = [a.elapsed_time(b) for a, b in _lib.timed.get("fs_knn_feat_tc", [])]
    _lib.timed.clear()

    t = torch.tensor([ms_total, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = float(t[0]), float(t[1])
    clouds = args.batch * world * args.steps

    if rank == 0:
        hbm_peak, peak_kind = peaks()
        # dominant HBM kernel: EdgeConv gather pass (ec2/ec3, Cp = 64). Algorithmic bytes per point:
        # table [a|b] 2*Cp*s + idx 4k + sel 4Cp + arg Cp + sy 4Cp   ("Kernels")
        s = 4        # the per-point tables are fp32 in every precision mode
        cp = 64
        bytes_per_point = 2 * cp * s + 4 * args.k + 4 * cp + cp + 4 * cp
        alg_bytes = bytes_per_point * args.batch * args.points
        g_ms = statistics.mean(gather_ms) if gather_ms else None
        achieved = alg_bytes / (g_ms * 1e-3) / 1e9 if g_ms else None
        line = {
            "metric": METRIC, "value": clouds / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision if args.precision != "fp32" else "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "e2e": {"value": clouds / (ms_e2e * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(x_in.numel() * 4 + y_in.numel() * 8), "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches), "cuda_graph": graph is not None,
            "clocks": clocks,
            "roofline": {"kernel": "edgeconv_gather_smem_kernel<16,5,0,float> (ec2/ec3 gather/max pass, Cp=64, train)",
                         "bound": "hbm", "achieved": achieved,
                         "peak": hbm_peak, "peak_kind": peak_kind, "unit": "GB/s",
                         "frac": (achieved / hbm_peak) if achieved else None,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch at this shape
                         # (profiles/r01_g_gather_smem_full.txt): the 37 MB of outputs stay in the 126 MB L2
                         "traffic": GATHER_DRAM_BYTES if (args.batch, args.points, args.k) == (32, 2048, 20) else None,
                         "launch_ms": g_ms, "algorithmic_bytes_per_launch": alg_bytes,
                         "launches_timed": len(gather_ms)},
        }
        if knn_ms:
            # feature-space kNN (ec2/ec3 graphs): the tcgen05 distance GEMM and its selection epilogue, whole entry
            # point (prep + tensor-core sweeps + finalize). Algorithmic flops 2*B*N^2*C; the kernel issues
            # 2 sweeps x K = 208 (bf16 hi/lo split) = 6.5x as many.
            kms = statistics.mean(knn_ms)
            alg_flops = 2.0 * args.batch * args.points * args.points * 64
            line["roofline_knn_gemm"] = {
                "kernel": "fs_knn_feat_tc entry point (tc_colsum + tc_split + knn_tc_select [tcgen05] + knn_tc_finalize)",
                "bound": "tensor", "achieved": alg_flops / (kms * 1e-3) / 1e12, "peak": TENSOR_PEAK_TFLOPS,
                "unit": "TFLOP/s", "frac": alg_flops / (kms * 1e-3) / 1e12 / TENSOR_PEAK_TFLOPS,
                "issued_tflops": alg_flops * 6.5 / (kms * 1e-3) / 1e12, "launch_ms": kms,
                "tensor_pipe_active_pct_ncu": 43.3, "launches_timed": len(knn_ms)}
        if world == 1 and not args.no_cpu_baseline:
            val, spstep, threads, n = cpu_reference_steps(args, 6, 2, max_seconds=30.0)
            line["cpu_baseline"] = {
                "value": val, "unit": UNIT, "cores": threads, "kind": "port", "host_cpus": os.cpu_count(),
                "sample": "%d clouds/step x %d steps of the same N=%d k=%d training step (oracle port of the "
                          "reference's PyTorch CPU path, fp32)" % (args.cpu_batch, n, args.points, args.k)}
        print(json.dumps(line), flush=True)
    if world > 1:
        # Tear-down: the captured CUDA graph holds NCCL work of this communicator, and destroying the process group
        # (or letting the interpreter finalise it) while the graph is alive has been seen to dead-lock after the
        # result line was printed. Drop the graph, drain the device, meet the other ranks once more and leave
        # without running the NCCL / CUDA finalisers.
        graph = None
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
