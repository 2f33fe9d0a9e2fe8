"""world_size-2 gloo tests (CPU) of the data-parallel host logic: bucket planning over the flat gradient arena,
the bucketed all-reduce, parameter broadcast, reduce_tensor / distribute_bn (dfd/timm/utils.py:256-274)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT)
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from deepfake_detection_b200.ddp import GradReducer, distribute_bn, reduce_tensor
        from deepfake_detection_b200.engine import Engine
        eng = Engine("efficientnet_b0", 2, 64, 64, device="plan-only")
        red = GradReducer(eng, bucket_mb=2.0)
        n = eng.n_params
        # every arena element is covered by exactly one bucket span, and buckets are issued in plan order
        cover = torch.zeros(n, dtype=torch.int32)
        last = -1
        for op_idx, spans in red.buckets:
            assert op_idx >= last
            last = op_idx
            for lo, hi in spans:
                cover[lo:hi] += 1
        assert int(cover.min()) == 1 and int(cover.max()) == 1, (int(cover.min()), int(cover.max()))
        assert red.buckets[-1][0] == len(eng.bwd_ops) - 1 and len(red.buckets) >= 2
        # a bucket may only be reduced after the last backward op that writes into it
        g0 = eng.grads32.data_ptr()
        for op_idx, spans in red.buckets:
            for j in range(op_idx + 1, len(eng.bwd_ops)):
                for a in eng.bwd_ops[j][2]:
                    if isinstance(a, int) and g0 <= a < g0 + 4 * n:
                        off = (a - g0) // 4
                        assert not any(lo <= off < hi for lo, hi in spans), (op_idx, j, eng.bwd_ops[j][1])
        # the bucketed all-reduce AVERAGES every gradient element exactly once (what DDP leaves in .grad, train.py:402-406)
        eng.grads32.copy_(torch.arange(n, dtype=torch.float32) % 1000 * (rank + 1))
        red.backward_and_reduce()
        expect = torch.arange(n, dtype=torch.float32) % 1000 * (sum(r + 1 for r in range(world)) / world)
        assert torch.equal(eng.grads32, expect)
        # a second plan (other batch size: the last partial batch of an epoch) over the SAME arenas gets its own cut points,
        # and the reducer replays THAT plan's backward, not the default engine's (reducer / engine mismatch)
        eng2 = Engine("efficientnet_b0", 1, 96, 96, device="plan-only", share_from=eng)
        assert eng2.grads32 is eng.grads32 and eng2.arena is eng
        b1, b2 = red.plan_for(eng), red.plan_for(eng2)
        assert b2[-1][0] == len(eng2.bwd_ops) - 1 and [sp for _, sp in b1] == [sp for _, sp in b2]
        eng.grads32.fill_(float(rank))
        calls = red.n_reduce_calls
        red.backward_and_reduce(eng2)
        assert float(eng.grads32.min()) == float(eng.grads32.max()) == (world - 1) / 2.0 and red.n_reduce_calls > calls
        other = Engine("efficientnet_b0", 1, 64, 64, device="plan-only")
        try:
            red.backward_and_reduce(other)
            raise AssertionError("foreign engine accepted")
        except RuntimeError:
            pass
        # parameter broadcast from rank 0
        eng.params32.fill_(float(rank + 5))
        red.broadcast_parameters()
        assert float(eng.params32.min()) == 5.0 and float(eng.params32.max()) == 5.0
        # reduce_tensor: mean over ranks; distribute_bn: mean (reduce=True) or rank-0 broadcast
        t = reduce_tensor(torch.tensor([float(rank + 1)]), world)
        assert abs(float(t) - sum(r + 1 for r in range(world)) / world) < 1e-6
        eng.buffers32.fill_(float(rank))
        distribute_bn(eng, world, reduce=True)
        assert abs(float(eng.buffers32[0]) - (world - 1) / 2.0) < 1e-6
        eng.buffers32.fill_(float(rank + 3))
        distribute_bn(eng, world, reduce=False)
        assert float(eng.buffers32[0]) == 3.0
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "FAIL: %r\n%s" % (e, traceback.format_exc()[-1500:])))


def test_ddp_host_logic_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 400
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


def test_plan_buckets_greedy():
    from deepfake_detection_b200.ddp import plan_buckets
    b = plan_buckets([(90, 100), (50, 90), (0, 50)], 30)
    assert b == [[(90, 100), (50, 90)], [(0, 50)]]
