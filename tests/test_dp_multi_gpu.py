"""GPU (needs 2 devices; skipped on a 1-GPU box): the data-parallel step over NCCL against a single-process step on the
whole batch -- sharding is exact (BatchNorm statistics are per sequence), the loss is the global mean, the clamp acts on
the REDUCED gradient.  Eager, per-segment CUDA graphs and a ragged split (3 + 2 sequences)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import cnn_linear_oracle as O  # noqa: E402


def _worker(rank, world, port, backbone, use_graph, q):
    import torch.distributed as dist
    import deepards_b200 as D
    from deepards_b200.data_parallel import DataParallelTrainer, shard_bounds
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    solo = dist.new_group([0])          # collective call on every rank; only rank 0 is a member
    dropout = backbone.endswith("+dropout")     # densenet.py:33-39 active: the masks are keyed by the global sequence index
    backbone = backbone.split("+")[0]
    sd = O.cnn_linear_state(backbone, seed=41, bn_perturb=0.1, **({"initial_planes": 16} if backbone == "resnet18" else {}))

    def make():
        bb = D.resnet18(initial_planes=16) if backbone == "resnet18" else D.densenet18(drop_rate=0.2 if dropout else 0.0)
        net = D.CNNLinearNetwork(bb, 20, 0)
        net.load_state_dict(sd)
        net = net.cuda().train()
        net.precision = "fp32"
        return net

    batches = [(O.synthetic_breaths(5, seed=500 + i).cuda(), O.synthetic_targets(5, seed=500 + i).cuda()) for i in range(4)]
    tr = DataParallelTrainer(make(), lr=1e-2, clip_val=0.01, use_graph=use_graph)
    losses = []
    for i in range(5):
        x, t = batches[i % 4]
        b, e = shard_bounds(5, world, rank)           # 3 + 2 sequences
        losses.append(float(tr.train_step(x[b:e], t[b:e], global_batch=5)))
    res = {"rank": rank, "params": tr.param_flat.cpu(), "losses": losses, "n": e - b}
    if rank == 0:
        ref = DataParallelTrainer(make(), lr=1e-2, clip_val=0.01, group=solo, use_graph=False)
        assert ref.world == 1
        ref_losses = [float(ref.train_step(*batches[i % 4])) for i in range(5)]
        res["ref_params"], res["ref_losses"] = ref.param_flat.cpu(), ref_losses
    q.put(res)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("backbone,use_graph", [("resnet18", False), ("resnet18", True), ("densenet18", True),
                                                ("densenet18+dropout", True)])
def test_two_rank_step_equals_single_process_step(backbone, use_graph):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, backbone, use_graph, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=600) for _ in procs), key=lambda r: r["rank"])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    r0, r1 = res
    assert torch.equal(r0["params"], r1["params"])                    # replicas stay identical
    # the global mean loss = sequence-weighted mean of the ranks' local means
    combined = [(l0 * r0["n"] + l1 * r1["n"]) / 5 for l0, l1 in zip(r0["losses"], r1["losses"])]
    diffs = [abs(lc - lr) for lc, lr in zip(combined, r0["ref_losses"])]
    # Step 1 runs on identical weights: with dropout on, a wrong mask moves the loss by > 1e-3, identical masks leave
    # rounding noise.  Later steps compare two fp32 trajectories whose gradients differ in summation order; with the
    # clamp at 0.01 and lr 1e-2 such trajectories part quickly (tools/dropout_sensitivity.py,
    # profiles/r02_dropout_sensitivity.txt: a 1e-7 perturbation of the initial weights grows to 1e-3 in the loss within
    # these 5 steps, with or without dropout), so the dropout case -- measured 2.4e-4 at step 5 -- gets a wider band; the
    # masks themselves are held bit for bit by tests/test_model_parity_gpu.py::test_dropout_*_ragged_split_*.
    assert diffs[0] <= 1e-6, diffs
    tol = 5e-4 if "dropout" in backbone else 1e-4
    assert max(diffs) <= tol, (diffs, combined, r0["ref_losses"])
    ref = r0["ref_params"]
    err = float((r0["params"] - ref).abs().max() / ref.abs().max())
    assert err <= tol, err
    print("%s graph=%s: 2-rank vs single-process: loss |diff| per step %s, parameters after 5 steps rel err %.2e" %
          (backbone, use_graph, " ".join("%.1e" % d for d in diffs), err))
