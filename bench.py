"""Benchmark of the cnn_linear hot path (BASELINE.json metric: train sequences/sec, 20x224 breaths).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--backbone resnet18|densenet18]

Workload (config.workload): BASELINE.json configs[1] -- cnn_linear / ResNet-18 training step on synthetic
256 x 20 x 1 x 224 batches per GPU, bf16 storage + tcgen05 convolutions, fp32 statistics / gradients / weights.
One "step" = forward + BCEWithLogits + backward + (N>1: NCCL gradient all-reduce) + clamp + SGD-Nesterov update,
i.e. everything `run_train_epoch` does per batch (deepards/train_ards_detector.py:139-173) except the host ETL.
N>1 (torchrun): weak scaling, 256 sequences per rank, sequences are independent so there is no other collective.

Printed JSON (rank 0, one line): see the contract in the task statement.  `value` = device-timed throughput with
the inputs already resident in HBM; `e2e` = the same step driven from pinned HOST buffers (H2D of the batch and
D2H of the loss inside the timed region).  `roofline` is measured live with CUDA events around every kernel call
of a few extra steps; `cpu_baseline` times the CPU oracle port on this box's host cores.

--impl reference: times the reference algorithm's CPU implementation (the oracle port of the reference's module
graph; the Python reference itself cannot travel to the GPU box) with all host threads, same metric and config.
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
sys.path.insert(0, ROOT)

SEQ_PER_GPU = 256
SUB_BATCH = 20
FLOP_PER_SEQ = {"resnet18": 4.5733e9, "densenet18": 0.66809e9}      # fwd+bwd, SURVEY.md section 8d
FLOP_FWD_PER_SEQ = {"resnet18": 1.5251e9, "densenet18": 0.22337e9}
BYTES_PER_SEQ_BF16 = {"resnet18": 17.29e6, "densenet18": 12.43e6}     # fwd+bwd activation traffic, bf16


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md).  One long-running
    `nvidia-smi -lms 25` process streams a line every 25 ms (starting a new process per sample takes ~100 ms, about as
    long as a short timed region); one process per sample runs next to it in case the stream is block-buffered.
    `mark()` brackets the timed regions: the summary uses the samples taken inside them (all samples if none fell inside)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.proc, self.windows = index, [], False, None, []

    def _parse(self, line):
        parts = [s.strip() for s in line.strip().split(",")]
        if len(parts) >= 7:
            self.rows.append((time.perf_counter(), parts))

    def _stream(self, cmd):
        try:
            self.proc = subprocess.Popen(cmd + ["-lms", "25"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
                                         bufsize=1)
            for line in self.proc.stdout:
                self._parse(line)
                if self.stop_flag:
                    break
        except Exception:
            pass

    def run(self):
        cmd = ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"]
        threading.Thread(target=self._stream, args=(cmd,), daemon=True).start()   # dense samples, if its output is unbuffered
        while not self.stop_flag:      # and one process per sample (~100 ms each), which always works
            try:
                self._parse(subprocess.run(cmd, capture_output=True, text=True, timeout=5).stdout)
            except Exception:
                pass
            time.sleep(0.05)

    def mark(self, t0, t1):
        self.windows.append((t0, t1))

    def stop(self):
        self.stop_flag = True
        try:
            if self.proc is not None:
                self.proc.terminate()
        except Exception:
            pass
        self.join(timeout=3)

    def summary(self):
        inside = [r for t, r in self.rows if any(a <= t <= b for a, b in self.windows)]
        rows = inside or [r for _, r in self.rows]
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        reasons = []
        for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5), ("sw_power_cap", 6)):
            if any(r[col].lower().startswith("active") for r in rows):
                reasons.append(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(rows), "samples_in_timed_region": len(inside)}


# ------------------------------------------------------------------------------------------------------------
# CPU reference arm (oracle port), also used for cpu_baseline
# ------------------------------------------------------------------------------------------------------------
def cpu_reference_steps(backbone, n_seq, steps, warmup, threads=None):
    """Times `steps` CPU training steps (forward + BCE + backward + SGD update) of the reference algorithm on
    n_seq sequences.  Returns (seq_per_s, ms_per_step, threads)."""
    import torch
    from oracle import cnn_linear_oracle as O
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = O.cnn_linear_state(backbone, seed=0)
    x = O.synthetic_breaths(n_seq, seed=1234)
    t = O.synthetic_targets(n_seq, seed=1234)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        _, _, grads = O.forward_backward(sd, x, t, clip_val=0.01, running_update=backbone.startswith("resnet"))
        with torch.no_grad():
            for k, g in grads.items():
                sd[k].add_(g + 1e-4 * sd[k], alpha=-1e-3)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    tot = sum(times)
    return n_seq * len(times) / tot, 1e3 * tot / len(times), threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_seq = 16  # bounded sample of the 256-sequence batch (BASELINE configs[0] shape); ~1 s per step on 8 cores
    v, ms, threads = cpu_reference_steps(args.backbone, n_seq, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "train sequences/sec (20x224 breaths)", "value": v, "unit": "sequences/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, cpu_sample=n_seq),
        "cpu_baseline": {"value": v, "unit": "sequences/s", "cores": threads, "kind": "port",
                         "sample": "%d of the %d sequences of one batch per step (oracle port of the reference's "
                                   "per-sequence loop, torch CPU fp32)" % (n_seq, SEQ_PER_GPU)},
        "e2e": {"value": v, "unit": "sequences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, cpu_sample=None):
    c = {"workload": "cnn_linear %s training step (fwd + BCEWithLogits + bwd + grad all-reduce + clamp + SGD-Nesterov), "
                     "%d x %d x 1 x 224 synthetic breaths per GPU (BASELINE.json configs[1])" % (args.backbone, SEQ_PER_GPU, SUB_BATCH),
         "backbone": args.backbone, "sequences_per_gpu": SEQ_PER_GPU, "sub_batch": SUB_BATCH, "precision": "bf16 storage / "
         "tcgen05 convolutions, fp32 statistics, gradients and weights", "parallelism": "dp%d" % args.gpus,
         "cuda_graph": (not args.no_graph) and ("whole step" if args.gpus == 1 else getattr(args, "dp_graph_desc", "per segment, NCCL between")),
         "l2": "working set per step (>1 GB of activations) exceeds the 126 MB L2; 4 resident input batches are rotated"}
    if cpu_sample:
        c["cpu_sample_sequences"] = cpu_sample
    return c


# ------------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    import deepards_b200 as D
    from deepards_b200 import _lib
    from deepards_b200.data_parallel import DataParallelTrainer
    from oracle import cnn_linear_oracle as O

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner on stdout at init; the contract is ONE JSON line there
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    lib = _lib.load()

    torch.manual_seed(0)
    bb = D.resnet18() if args.backbone == "resnet18" else D.densenet18()
    net = D.CNNLinearNetwork(bb, SUB_BATCH, 0).to(dev)
    net.precision = args.precision
    net.train()
    trainer = DataParallelTrainer(net, lr=1e-3, optimizer="sgd", weight_decay=1e-4, clip_val=0.01,
                                  use_graph=not args.no_graph)

    n_in = 4
    xs_host = [O.synthetic_breaths(SEQ_PER_GPU, seed=1234 + rank * 17 + i).pin_memory() for i in range(n_in)]
    ts_host = [O.synthetic_targets(SEQ_PER_GPU, seed=1234 + rank * 17 + i).pin_memory() for i in range(n_in)]
    xs = [x.to(dev) for x in xs_host]
    ts = [t.to(dev) for t in ts_host]

    sampler_ref = [None]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        if sampler_ref[0] is not None:
            sampler_ref[0].mark(t0, time.perf_counter())
        ms = e0.elapsed_time(e1)
        if world > 1:
            tms = torch.tensor([ms], device=dev)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms = float(tms)
        return ms

    def resident_step(i):
        trainer.train_step(xs[i % n_in], ts[i % n_in])

    # e2e: the batch starts in pinned HOST memory.  Like a DataLoader with pin_memory + non_blocking copies
    # (train_ards_detector.py:324-337, 150-152), the copy of batch i+1 is issued on a copy stream while step i runs; the
    # loss of EVERY step is read back to the host (the reference's Meter does, metrics.py:142-153), so each step ends
    # with a stream synchronisation.  H2D and D2H are inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    x_stage = [torch.empty_like(xs[0]) for _ in range(2)]
    t_stage = [torch.empty_like(ts[0]) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    loss_host = torch.zeros(1).pin_memory()

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            x_stage[i % 2].copy_(xs_host[i % n_in], non_blocking=True)
            t_stage[i % 2].copy_(ts_host[i % n_in], non_blocking=True)
            ready[i % 2].record(copy_stream)

    def e2e_step(i):
        cur = torch.cuda.current_stream()
        cur.wait_event(ready[i % 2])
        loss = trainer.train_step(x_stage[i % 2], t_stage[i % 2])
        loss_host.copy_(loss, non_blocking=True)
        prefetch(i + 1)          # overlaps this step's kernels; buffer (i+1)%2 was consumed by step i-1, which is complete
        cur.synchronize()        # the trainer reads the loss every step

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()            # started before the warm-up so that it is streaming by the time the clock starts
        sampler_ref[0] = sampler
    for i in range(max(args.warmup, 3)):
        resident_step(i)
    l0, g0 = lib.dards_launch_count(), trainer.graph_launches
    ms = timed(resident_step, args.steps)
    launches = (lib.dards_launch_count() - l0) + (trainer.graph_launches - g0)
    prefetch(0)
    for i in range(2):
        e2e_step(i)
    e2e_base = 2

    def e2e_timed_step(i):
        e2e_step(e2e_base + i)

    ms_e2e = timed(e2e_timed_step, args.steps)
    if sampler:
        sampler.stop()
    final_loss = float(trainer.loss_buf)

    # ---- per-kernel timing for the roofline (extra, untimed-for-throughput steps) ---------------------------
    roof = kernel_roofline(trainer, xs[0], ts[0], args) if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    if world > 1:
        args.dp_graph_desc = ("whole step incl. the bucketed NCCL all-reduces (captured on the communication stream)"
                              if trainer.dp_graph == "whole" else "per segment, NCCL between")
    seqs = SEQ_PER_GPU * world * args.steps
    value = seqs / (ms / 1e3)
    e2e_value = seqs / (ms_e2e / 1e3)
    cpu_v, cpu_ms, cpu_threads = cpu_reference_steps(args.backbone, 16, 3, 1) if world == 1 and not args.no_cpu else (None, None, None)
    line = {
        "metric": "train sequences/sec (20x224 breaths)", "value": value, "unit": "sequences/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32",
        "data": "synthetic", "config": workload_config(args),
        "e2e": {"value": e2e_value, "unit": "sequences/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(xs_host[0].numel() * 4 + ts_host[0].numel() * 4), "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches),
        "clocks": sampler.summary() if sampler else None,
        "roofline": roof,
        "final_loss": final_loss,
    }
    if roof:
        ceil = min(roof["step_ceiling_seq_per_s"].values())
        line["frac_of_step_ceiling"] = value / world / ceil
    if cpu_v is not None:
        line["cpu_baseline"] = {"value": cpu_v, "unit": "sequences/s", "cores": cpu_threads, "kind": "port",
                                "sample": "3 steps of 16 sequences (BASELINE configs[0] shape) of the same training step, "
                                          "oracle port of the reference loop, torch CPU fp32, %.0f ms/step" % cpu_ms}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def kernel_roofline(trainer, x, t, args, steps=3):
    """CUDA-event time of every recorded kernel call over a few steps, grouped by entry point.  Returns the
    roofline object of the dominant kernel class plus the per-class breakdown."""
    import torch
    plan = trainer.plan_for(x)
    P = peaks()
    st = torch.cuda.current_stream()
    agg = {}
    for _ in range(steps):
        plan.load_input(x)
        for rec in (plan.pack, plan.fwd, plan.bwd):
            evs = []
            for name, f, a in rec.calls:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                rc = f(*a, st.cuda_stream)
                e1.record(st)
                if rc != 0:
                    raise RuntimeError(name)
                evs.append((name, a, e0, e1))
            torch.cuda.synchronize()
            for name, a, e0, e1 in evs:
                key = name
                if name in ("dards_conv1d_fwd", "dards_conv1d_dgrad", "dards_conv1d_wgrad"):
                    key = name + (":tcgen05" if (a[-1] & 0xff) == 1 else ":simt")
                d = agg.setdefault(key, {"ms": 0.0, "launches": 0, "flops": 0.0, "bytes": 0.0})
                d["ms"] += e0.elapsed_time(e1)
                d["launches"] += 1
                esz = 2 if args.precision == "bf16" else 4
                if name == "dards_conv1d_fwd" or name == "dards_conv1d_dgrad":
                    n, l_in, l_out, cin, cout, k = a[4], a[5], a[6], a[7], a[8], a[12]
                    d["flops"] += 2.0 * n * l_out * cin * cout * k
                    d["bytes"] += esz * n * (l_in * cin + l_out * cout)
                elif name == "dards_conv1d_wgrad":
                    n, l_in, l_out, cin, cout, k = a[6], a[7], a[8], a[9], a[10], a[13]
                    d["flops"] += 2.0 * n * l_out * cin * cout * k
                    d["bytes"] += esz * n * (l_in * cin + l_out * cout)
                elif name == "dards_conv1d_bn_fwd":
                    # convolution + BatchNorm in one kernel: reads the input, writes y (kept for the backward) and, when the
                    # whole normalisation runs in the epilogue, the activation (+ reads the residual)
                    n, l_in, l_out, cin, cout, k = a[10], a[12], a[13], a[14], a[15], a[20]
                    d["flops"] += 2.0 * n * l_out * cin * cout * k
                    d["bytes"] += esz * n * (l_in * cin + l_out * cout * (1 + (1 if a[3] else 0) + (1 if a[4] else 0)))
                elif name == "dards_gbn_apply_fwd":
                    g, rows, c = a[16], a[17], a[18]
                    d["bytes"] += esz * g * rows * c * (2 + (1 if a[2] else 0) + (1 if a[9] else 0))
                elif name == "dards_gbn_fwd":
                    # algorithmic traffic: x read once, out written once (+ the residual read)
                    g, rows, c = a[7], a[8], a[9]
                    d["bytes"] += esz * g * rows * c * (2 + (1 if a[2] else 0))
                elif name == "dards_gbn_bwd":
                    # dout and x read once, dx written once (+ mask read, + residual-gradient write, + dx read when accumulating)
                    g, rows, c = a[12], a[13], a[14]
                    d["bytes"] += esz * g * rows * c * (3 + (1 if a[2] else 0) + (1 if a[9] else 0) + (1 if a[8] else 0))
        plan.fwd_serial += 1
        plan.bwd_serial = plan.fwd_serial
    total_ms = sum(d["ms"] for d in agg.values())
    top = max(agg.items(), key=lambda kv: kv[1]["ms"])
    name, d = top
    out = {"kernel": name, "share_of_step": d["ms"] / total_ms, "launches_per_step": d["launches"] / steps,
           "avg_launch_ms": d["ms"] / d["launches"], "traffic": None,
           "breakdown_ms_per_step": {k: round(v["ms"] / steps, 4) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])}}
    if d["flops"] > 0:
        ach = d["flops"] / (d["ms"] / 1e3) / 1e12
        out.update({"bound": "tensor", "achieved": ach, "peak": P["tf_sustained"], "unit": "TFLOP/s",
                    "frac": ach / P["tf_sustained"], "peak_source": P["src"] + " bf16 sustained (kernel timed inside a long step)"})
    else:
        ach = d["bytes"] / (d["ms"] / 1e3) / 1e9 if d["bytes"] > 0 else None
        out.update({"bound": "hbm", "achieved": ach, "peak": P["hbm"], "unit": "GB/s",
                    "frac": ach / P["hbm"] if ach else None, "peak_source": P["src"] + " copy bandwidth"})
    # measured DRAM traffic per launch of that kernel from the committed `ncu --set full` capture, if there is one
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        out["traffic"] = json.load(open(tpath)).get(args.backbone, {}).get(name)
    out["per_kernel"] = {}
    for k, v in agg.items():
        e = {"ms_per_step": round(v["ms"] / steps, 4), "launches_per_step": v["launches"] / steps}
        if v["flops"] > 0:
            e["tflops"] = round(v["flops"] / (v["ms"] / 1e3) / 1e12, 1)
            e["frac_of_tensor_peak"] = round(e["tflops"] / P["tf_sustained"], 3)
        if v["bytes"] > 0:
            e["algorithmic_gbs"] = round(v["bytes"] / (v["ms"] / 1e3) / 1e9, 1)
            e["frac_of_hbm_peak"] = round(e["algorithmic_gbs"] / P["hbm"], 3)
        out["per_kernel"][k] = e
    # whole-step view against the algorithmic ceilings of SURVEY.md section 8d
    conv_ms = sum(v["ms"] for k, v in agg.items() if k.startswith("dards_conv1d")) / steps
    conv_fl = sum(v["flops"] for k, v in agg.items() if k.startswith("dards_conv1d")) / steps
    out["conv_tflops_all"] = conv_fl / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else None
    out["conv_share_of_step"] = conv_ms / (total_ms / steps)
    out["step_ceiling_seq_per_s"] = {"tensor": P["tf_sustained"] * 1e12 / FLOP_PER_SEQ[args.backbone],
                                     "hbm": P["hbm"] * 1e9 / BYTES_PER_SEQ_BF16[args.backbone]}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--backbone", default="resnet18", choices=["resnet18", "densenet18"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-graph", action="store_true", help="issue every kernel from Python instead of one CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
