"""BASELINE.json configs[4] on N GPUs: cnn_linear DenseNet-18 inference + GradCAM read maps over a synthetic 24-hour
recording (720 sequences of 20 x 224), sharded over the ranks (replicas only, no data-path collective).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_gradcam_dp.py [--sequences 720] [--steps 20]

Every rank holds the whole recording on its device, runs its contiguous shard through one forward plan + one
dards_gradcam launch, and all ranks gather the uint8 maps and logits in order.  Timed on the device, max over ranks
(barrier + synchronize on both sides); rank 0 also recomputes the whole recording alone and checks that the gathered maps
are bit-identical.  One JSON line on stdout (rank 0)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import deepards_b200 as D  # noqa: E402
from deepards_b200 import gradcam as G  # noqa: E402
from deepards_b200 import synthetic as S  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sequences", type=int, default=720)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--precision", default="bf16")
    a = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    torch.manual_seed(0)
    net = D.CNNLinearNetwork(D.densenet18(), 20, 0).to(dev).eval()
    net.precision = a.precision
    if world > 1:
        for p in net.parameters():
            dist.broadcast(p.data, src=0)
    xs = [S.synthetic_breaths(a.sequences, seed=60 + i).to(dev) for i in range(2)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i):
        return G.recording_maps(net, xs[i % 2], None, resized_len=224)

    for i in range(3):
        maps, logits = step(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        maps, logits = step(i)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    maps, logits = step(0)
    ok = None
    if rank == 0:
        solo = G.compute_maps(net, xs[0], None, resized_len=224)
        ok = bool(torch.equal(solo.read_resized, maps) and torch.equal(solo.logits, logits))
        print(json.dumps({"config": "configs[4] GradCAM over a 24-hour recording, sharded", "backbone": "densenet18",
                          "n_gpus": world, "sequences": a.sequences, "sequences_per_gpu": -(-a.sequences // world),
                          "precision": a.precision, "ms_per_recording": round(float(ms), 4),
                          "sequences_per_s": round(a.sequences / float(ms) * 1e3),
                          "outputs": "read maps (sequences, 20, 224) uint8 + logits, gathered on every rank",
                          "equals_single_gpu_bit_for_bit": ok}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if ok is False:
        sys.exit(3)


if __name__ == "__main__":
    main()
