"""Worker of tests/test_boundary_gpu.py::test_native_ddp_two_ranks_match_oracle (launched by torch.distributed.run,
one rank per GPU, NCCL).  Also usable by hand:

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/ddp_worker.py

What it checks (boundary rows H6 / 8b, dfd/runners/train.py:402-406,621-637):
  * protocol path — the reference's own loop body on a NativeDDP-wrapped NativeModel: `output = model(input)`,
    `loss_fn(output, target)`, `loss.backward()` (gradient mean across ranks inside backward), `optimizer.step()`;
  * fused path — the mirrored `train_epoch` over two batches whose LAST one is smaller, so the runner switches to a second
    execution plan in mid-epoch (the configuration in which a reducer bound to the wrong plan all-reduced garbage);
  * both against the CPU oracle: per-rank gradients on each rank's batch (rank-local BN statistics, apex semantics),
    `grad_hook` = mean over ranks, one optimizer step per batch; replicas must hold bit-identical weights afterwards;
  * synchronised BatchNorm (`convert_syncbn_model`, train.py:388-394): the same protocol loop equals ONE oracle process
    training on the concatenation of both ranks' batches (global batch statistics, global-mean loss);
  * an unwrapped NativeModel refuses to run backward in a multi-rank job.
"""
import json
import os
import sys
from types import SimpleNamespace

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def rel(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def oracle_epoch(spec, sd0, rank_batches, world):
    """rank_batches[r] = list of (x, y). Every rank's BN statistics are local; gradients are averaged before the update.
    Running statistics: returns rank 0's."""
    from oracle import train as OT
    sds = [{k: v.clone() for k, v in sd0.items()} for _ in range(world)]
    opts = [OT.OptState(kind="sgd", lr=0.01, momentum=0.9, weight_decay=1e-4) for _ in range(world)]
    losses = []
    for i in range(len(rank_batches[0])):
        outs = [OT.train_step(spec, sds[r], rank_batches[r][i][0], rank_batches[r][i][1], None, act_dtype=torch.float16)
                for r in range(world)]
        mean = {k: sum(o["grads"][k] for o in outs) / world for k in outs[0]["grads"]}
        for r in range(world):
            params, _ = OT.split_state(spec, sds[r])
            OT.optimizer_step(opts[r], params, mean)
        losses.append(sum(float(o["loss"]) for o in outs) / world)
    return sds[0], losses


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    torch.set_num_threads(16)
    from deepfake_detection_b200 import loss as NL
    from deepfake_detection_b200._lib import NativeError
    from deepfake_detection_b200.arch import get_spec
    from deepfake_detection_b200.ddp import NativeDDP
    from deepfake_detection_b200.models import create_model
    from deepfake_detection_b200.optim import create_optimizer
    from deepfake_detection_b200.runners.train import train_epoch
    from oracle.weights import synth_batch, synth_state
    spec = get_spec("efficientnet_b0")
    sd0 = synth_state(spec, seed=7)
    args = SimpleNamespace(opt="sgd", lr=0.01, momentum=0.9, weight_decay=1e-4, opt_eps=1e-8, prefetcher=True, mixup=0.0,
                           mixup_off_epoch=0, num_classes=2, smoothing=0.0, distributed=True, world_size=world, local_rank=rank,
                           log_interval=1, save_images=False, recovery_interval=0, tta=0)
    # per-rank batches: two of 16 images and a LAST batch of 8 (another plan)
    sizes = (16, 16, 8)
    rank_batches = [[synth_batch(n, 3, 96, 96, seed=1000 * (r + 1) + i) for i, n in enumerate(sizes)] for r in range(world)]
    mine = rank_batches[rank]
    report = {}

    # an unwrapped model must refuse a multi-rank backward
    m0 = create_model("efficientnet_b0", num_classes=2, dtype="fp16")
    m0.load_state_dict(sd0)
    m0.train()
    try:
        torch.nn.CrossEntropyLoss()(m0(mine[2][0].cuda()), mine[2][1].cuda()).backward()
        report["raises_without_wrapper"] = False
    except NativeError:
        report["raises_without_wrapper"] = True
    del m0

    sd_o, losses_o = oracle_epoch(spec, sd0, rank_batches, world) if rank == 0 else (None, None)
    sd_sync, losses_sync = None, None
    if rank == 0:
        from oracle import train as OT
        sd_sync = {k: v.clone() for k, v in sd0.items()}
        ost = OT.OptState(kind="sgd", lr=0.01, momentum=0.9, weight_decay=1e-4)
        losses_sync = []
        for i in range(len(sizes)):
            gx = torch.cat([rank_batches[r][i][0] for r in range(world)])
            gy = torch.cat([rank_batches[r][i][1] for r in range(world)])
            losses_sync.append(float(OT.train_step(spec, sd_sync, gx, gy, ost, act_dtype=torch.float16)["loss"]))

    for flavour in ("protocol", "fused", "syncbn"):
        model = create_model("efficientnet_b0", num_classes=2, dtype="fp16")
        sd_r = {k: (v + 0.01 * rank if v.dtype.is_floating_point and rank else v) for k, v in sd0.items()}   # ranks start apart ...
        model.load_state_dict(sd_r)
        if flavour == "syncbn":
            from deepfake_detection_b200.ddp import convert_syncbn_model
            model = convert_syncbn_model(model)
        model = NativeDDP(model, delay_allreduce=True)                                                   # ... DDP broadcasts rank 0
        optimizer = create_optimizer(args, model)

        class Loader(list):
            mixup_enabled = False

        loader = Loader((x.cuda(), y.cuda()) for x, y in mine)
        if flavour in ("protocol", "syncbn"):
            # the reference loop body verbatim (train.py:621-637), apex loss scaling left out (use_amp=False)
            model.train()
            loss_fn = torch.nn.CrossEntropyLoss()
            losses = []
            for input, target in loader:
                output = model(input)
                loss = loss_fn(output, target)
                optimizer.zero_grad()
                loss.backward()
                optimizer.step()
                t = loss.detach().clone()
                dist.all_reduce(t)
                losses.append(float(t) / world)
            loss_mean = sum(losses) / len(losses)
        else:
            m = train_epoch(0, model, loader, optimizer, NL.CrossEntropyLoss(), args)
            # train_epoch weights each batch loss by its size (AverageMeter): redo that for the oracle below
            loss_mean = m["loss"]
        torch.cuda.synchronize()
        sd = model.module.state_dict()
        flat = torch.cat([v.flatten().float() for k, v in sd.items() if v.dtype.is_floating_point and "running" not in k])
        gathered = [torch.zeros_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        if rank == 0:
            ref_sd, ref_losses = (sd_sync, losses_sync) if flavour == "syncbn" else (sd_o, losses_o)
            worst = max(rel(sd[k], ref_sd[k]) for k in ref_sd if ref_sd[k].dtype.is_floating_point and ref_sd[k].dim() > 1)
            if flavour == "syncbn":     # running statistics are global too: identical on every rank and equal to the oracle's
                worst = max(worst, max(rel(sd[k], ref_sd[k]) for k in ref_sd if "running" in k))
            lo = sum(ref_losses) / len(ref_losses) if flavour != "fused" else \
                sum(l * n for l, n in zip(ref_losses, sizes)) / sum(sizes)
            report[flavour] = dict(weights_rel_worst=worst, ranks_identical=all(torch.equal(g, gathered[0]) for g in gathered),
                                   loss=loss_mean, loss_oracle=lo, plans=len(model.module._engines),
                                   reduce_calls=model.reducer.n_reduce_calls)
        del model, optimizer
    if rank == 0:
        out = os.environ.get("DFD_DDP_OUT", os.path.join(ROOT, "gpurun_out", "ddp_worker.json"))
        os.makedirs(os.path.dirname(out), exist_ok=True)
        with open(out, "w") as f:
            json.dump(report, f, indent=1)
        print(json.dumps(report))
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
