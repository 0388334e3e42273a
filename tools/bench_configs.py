"""Device-timed throughput of the BASELINE.json configs that are not the bench line (they are parity-test cases; this
records what they run at on one B200 for DESIGN.md / profiles/):

  configs[3]  padded_breath_by_breath single-breath classifier, forward only, 1K .. 64K breaths, both backbones
  configs[4]  cnn_linear DenseNet-18 inference + GradCAM maps for a synthetic 24-hour recording (720 sequences; 90 per
              GPU when sharded over 8) -- all maps from ONE forward plan + ONE dards_gradcam launch

    python tools/bench_configs.py [inference] [gradcam] [scaling]

One JSON object per line on stdout.  Timing: CUDA events around `steps` calls after 3 warm-up calls; inputs rotate over
buffers whose activations exceed the L2; `e2e` includes the pinned-host -> device copy of the windows and the
device -> host copy of the results.
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import deepards_b200 as D  # noqa: E402
from deepards_b200 import gradcam as G  # noqa: E402
from oracle import cnn_linear_oracle as O  # noqa: E402  (synthetic inputs + the CPU baseline leg only)

DEV = torch.device("cuda", 0)


def timed(fn, steps, warmup=3):
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def padded(x, seed):
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(60, 201, (x.shape[0], x.shape[1]), generator=g)
    return x * (torch.arange(224).view(1, 1, 1, 224) < lens.view(x.shape[0], x.shape[1], 1, 1))


def inference():
    for backbone in ("resnet18", "densenet18"):
        torch.manual_seed(0)
        bb = D.resnet18() if backbone == "resnet18" else D.densenet18()
        net = D.CNNSingleBreathLinearNetwork(bb).to(DEV)
        net.precision = "bf16"
        net = net.train() if backbone == "resnet18" else net.eval()   # train_ards_detector.py:448: no eval() BN switch
        flop = 76.26e6 if backbone == "resnet18" else 11.17e6        # per breath, forward
        for n_breaths in (1024, 4096, 16384, 65536):
            b = n_breaths // 20 + (1 if n_breaths % 20 else 0)
            hosts = [padded(O.synthetic_breaths(min(b, 512), seed=40 + i), 7 + i).repeat((b + 511) // 512, 1, 1, 1)[:b]
                     .contiguous().pin_memory() for i in range(2)]
            xs = [h.to(DEV) for h in hosts]
            out_host = torch.empty((b, 20, 2)).pin_memory()

            def resident(i):
                with torch.no_grad():
                    net(xs[i % 2], None)

            def e2e(i):
                with torch.no_grad():
                    x = hosts[i % 2].to(DEV, non_blocking=True)
                    out_host.copy_(net(x, None), non_blocking=True)
                torch.cuda.current_stream().synchronize()

            steps = 20 if n_breaths <= 16384 else 6
            ms = timed(resident, steps)
            ms_e2e = timed(e2e, steps)
            print(json.dumps({"config": "configs[3] inference sweep", "backbone": backbone, "breaths": b * 20,
                              "ms_per_call": round(ms, 4), "breaths_per_s": round(b * 20 / ms * 1e3),
                              "tflops": round(b * 20 * flop / ms / 1e9, 1),
                              "e2e_ms_per_call": round(ms_e2e, 4), "e2e_breaths_per_s": round(b * 20 / ms_e2e * 1e3),
                              "dtype": "bf16"}), flush=True)
            del xs, hosts
            bb.__dict__.pop("_dards_plans", None)
            net.__dict__.pop("_dards_plans", None)
            torch.cuda.empty_cache()


def gradcam():
    torch.manual_seed(0)
    net = D.CNNLinearNetwork(D.densenet18(), 20, 0).to(DEV).eval()
    for precision in ("bf16", "fp32"):
        net.precision = precision
        for n_seq in (90, 720):
            hosts = [O.synthetic_breaths(n_seq, seed=60 + i).pin_memory() for i in range(2)]
            xs = [h.to(DEV) for h in hosts]
            res_host = torch.empty((n_seq, 20, 224), dtype=torch.uint8).pin_memory()
            log_host = torch.empty((n_seq, 2)).pin_memory()

            def resident(i):
                G.compute_maps(net, xs[i % 2], None, resized_len=224)

            def e2e(i):
                x = hosts[i % 2].to(DEV, non_blocking=True)
                m = G.compute_maps(net, x, None, resized_len=224)
                res_host.copy_(m.read_resized, non_blocking=True)
                log_host.copy_(m.logits, non_blocking=True)
                torch.cuda.current_stream().synchronize()

            ms = timed(resident, 10)
            ms_e2e = timed(e2e, 10)
            print(json.dumps({"config": "configs[4] GradCAM over a recording", "backbone": "densenet18",
                              "sequences": n_seq, "precision": precision, "ms_per_call": round(ms, 4),
                              "sequences_per_s": round(n_seq / ms * 1e3), "e2e_ms_per_call": round(ms_e2e, 4),
                              "e2e_sequences_per_s": round(n_seq / ms_e2e * 1e3),
                              "outputs": "read maps resized to (n_seq, 20, 224) uint8 + logits"}), flush=True)
    # the reference's way on this box's host cores: forward + one-hot backward + numpy per sequence (oracle port)
    sd = O.cnn_linear_state("densenet18", seed=0)
    x = O.synthetic_breaths(4, seed=60)
    torch.set_num_threads(os.cpu_count() or 1)
    O.gradcam_read_cam(sd, x[0], None)
    t0 = time.perf_counter()
    for i in range(4):
        O.gradcam_read_cam(sd, x[i], None)
    dt = (time.perf_counter() - t0) / 4
    print(json.dumps({"config": "configs[4] GradCAM, CPU oracle port", "sequences_per_s": round(1 / dt, 2),
                      "cores": os.cpu_count(), "sample": "4 sequences"}), flush=True)


def scaling():
    z = O.synthetic_breaths(256, seed=1).double() * O.DATASET_STD + O.DATASET_MU
    for dt in (torch.float64, torch.float32):
        raws = [z.to(dt).to(DEV) for _ in range(8)]
        sc = D.WindowScaler(O.DATASET_MU, O.DATASET_STD)
        out = torch.empty(z.shape, dtype=torch.float32, device=DEV)
        ms = timed(lambda i: sc(raws[i % 8], out=out), 50)
        nbytes = z.numel() * (raws[0].element_size() + 4)
        print(json.dumps({"config": "input scaling (dataset.py:1379) of one 256-sequence batch", "raw_dtype": str(dt),
                          "us_per_call": round(ms * 1e3, 2), "gbs": round(nbytes / ms / 1e6, 1)}), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["inference", "gradcam", "scaling"]
    for w in which:
        {"inference": inference, "gradcam": gradcam, "scaling": scaling}[w]()
