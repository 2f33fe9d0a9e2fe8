"""Data-parallel gradient reduction for the native engine (replaces the DDP wrapper at dfd/runners/train.py:402,406
and `reduce_tensor`, dfd/timm/utils.py:256-260).

One process per GPU (`torch.distributed`, NCCL over NVLink/NVSwitch; gloo for the CPU-side tests).  The engine's
gradients live in ONE flat fp32 arena laid out in forward execution order, so:
  * the backward plan is cut into segments at block boundaries; after segment i the arena suffix that backward has
    finished writing is all-reduced (SUM) on a side stream while segment i+1 computes — bucketed, overlapped
    all-reduce without per-tensor hooks;
  * the 1/world_size of the mean is folded into the optimizer kernel (`grad_scale`), not a separate pass;
  * BN running statistics stay rank-local (apex DDP semantics); `distribute_bn` mirrors utils.py:263-274.
"""
import torch
import torch.distributed as dist

from .engine import _ptr


def reduce_tensor(tensor, n, group=None):
    """dfd/timm/utils.py:256-260"""
    rt = tensor.clone()
    dist.all_reduce(rt, op=dist.ReduceOp.SUM, group=group)
    rt /= n
    return rt


def distribute_bn(engine, world_size, reduce=False, group=None):
    """dfd/timm/utils.py:263-274 on the engine's flat running-stat arena (one collective instead of 2 per BN)."""
    if reduce:
        dist.all_reduce(engine.buffers32, op=dist.ReduceOp.SUM, group=group)
        engine.buffers32 /= float(world_size)
    else:
        dist.broadcast(engine.buffers32, 0, group=group)


def plan_buckets(spans, bucket_elems):
    """spans: list of (lo, hi) arena ranges in the order backward completes them. Greedily merges consecutive spans
    into buckets of at least `bucket_elems` elements. Returns list of lists of (lo, hi)."""
    buckets, cur, size = [], [], 0
    for lo, hi in spans:
        if hi <= lo:
            continue
        cur.append((lo, hi))
        size += hi - lo
        if size >= bucket_elems:
            buckets.append(cur)
            cur, size = [], 0
    if cur:
        buckets.append(cur)
    return buckets


class GradReducer:
    def __init__(self, engine, group=None, bucket_mb=8.0):
        self.engine = e = engine
        self.group = group
        self.world = dist.get_world_size(group)
        self.side = None if e._plan_only else torch.cuda.Stream(device=e.device)
        # cut the backward plan where a block's last gradient has been produced: op index -> arena ranges done.
        # Arena layout = [decay tensors in exec order | no-decay tensors in exec order]; backward walks both from
        # the end towards the start, so after the ops of a block the suffixes starting at that block's first
        # tensors are final.
        first_d, first_nd = {}, {}
        from .arch import is_no_decay
        for n in e.param_names:
            o, s, k = e.p_off[n]
            key = self._owner(n)
            (first_nd if is_no_decay(n, s) else first_d).setdefault(key, o)
        self._cuts = []          # (bwd op index after which [d_lo, prev_d) and [nd_lo, prev_nd) are final)
        owners = []
        for n in e.param_names:
            k = self._owner(n)
            if k not in owners:
                owners.append(k)
        # map plan positions: find the last bwd op that writes into each owner (by gradient pointer range)
        last_op = {}
        for idx, (_, name, args) in enumerate(e.bwd_ops):
            for a in args:
                if isinstance(a, int) and _ptr(e.grads32) <= a < _ptr(e.grads32) + e.grads32.numel() * 4:
                    off = (a - _ptr(e.grads32)) // 4
                    last_op[self._owner_of_offset(off)] = idx
        prev_d, prev_nd = e.n_decay, e.n_params
        spans_by_op = []
        for k in reversed(owners):
            d_lo = first_d.get(k, prev_d)
            nd_lo = first_nd.get(k, prev_nd)
            spans_by_op.append((last_op.get(k, len(e.bwd_ops) - 1), [(d_lo, prev_d), (nd_lo, prev_nd)]))
            prev_d, prev_nd = min(d_lo, prev_d), min(nd_lo, prev_nd)
        # ensure coverage of the arena heads (padding) by the final bucket
        spans_by_op.append((len(e.bwd_ops) - 1, [(0, prev_d), (e.n_decay, prev_nd)]))
        # monotone op order, then bucket by size
        bucket_elems = int(bucket_mb * 1024 * 1024 / 4)
        self.buckets = []        # (op index, [(lo, hi), ...])
        cur, size, cur_op = [], 0, 0
        for op_idx, spans in spans_by_op:
            cur_op = max(cur_op, op_idx)
            for lo, hi in spans:
                if hi > lo:
                    cur.append((lo, hi))
                    size += hi - lo
            if size >= bucket_elems:
                self.buckets.append((cur_op, self._merge(cur)))
                cur, size = [], 0
        if cur:
            self.buckets.append((len(e.bwd_ops) - 1, self._merge(cur)))
        if self.buckets:
            self.buckets[-1] = (len(e.bwd_ops) - 1, self.buckets[-1][1])

    @staticmethod
    def _merge(spans):
        spans = sorted(spans)
        out = []
        for lo, hi in spans:
            if out and lo <= out[-1][1]:
                out[-1] = (out[-1][0], max(out[-1][1], hi))
            else:
                out.append((lo, hi))
        return out

    @staticmethod
    def _owner(name):
        parts = name.split(".")
        if parts[0] == "blocks":
            return ".".join(parts[:3])
        if parts[0].startswith("layer"):
            return ".".join(parts[:2])
        if parts[0] in ("conv_head", "bn2", "classifier", "fc"):
            return "head"
        return "stem"

    def _owner_of_offset(self, off):
        e = self.engine
        if not hasattr(self, "_off_index"):
            self._off_index = sorted((o, o + k, self._owner(n)) for n, (o, s, k) in e.p_off.items())
        for lo, hi, k in self._off_index:
            if lo <= off < hi:
                return k
        return "stem"

    def broadcast_parameters(self):
        """Rank-0 weights to every rank at wrap time (DDP constructor semantics)."""
        e = self.engine
        dist.broadcast(e.params32, 0, group=self.group)
        dist.broadcast(e.buffers32, 0, group=self.group)
        if not e._plan_only:
            e.sync_weights()

    def backward_and_reduce(self):
        """Runs the engine's backward plan on the current stream, launching each bucket's all-reduce on the side
        stream as soon as the ops that produce it have been enqueued; joins the side stream at the end."""
        e = self.engine
        if e._plan_only:        # host-logic tests (gloo on CPU): no kernels, only the bucketed collectives
            for _, spans in self.buckets:
                for lo, hi in spans:
                    dist.all_reduce(e.grads32[lo:hi], op=dist.ReduceOp.SUM, group=self.group)
            return
        main = torch.cuda.current_stream()
        st = main.cuda_stream
        start = 0
        for op_idx, spans in self.buckets:
            e._run(e.bwd_ops[start:op_idx + 1], st, True)
            start = op_idx + 1
            ev = torch.cuda.Event()
            ev.record(main)
            self.side.wait_event(ev)
            with torch.cuda.stream(self.side):
                for lo, hi in spans:
                    dist.all_reduce(e.grads32[lo:hi], op=dist.ReduceOp.SUM, group=self.group)
        if start < len(e.bwd_ops):
            e._run(e.bwd_ops[start:], st, True)
        main.wait_stream(self.side)
