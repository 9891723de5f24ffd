"""Multi-GPU paths (new capabilities required by BASELINE.json; the reference has no distributed code):

  * data-parallel training: one process per GPU, identical replicas, rank-local BatchNorm statistics and loss; the only
    exchange is the gradient all-reduce, issued per ~25 MB bucket of the flat gradient buffer as soon as backward has
    completed that range (buckets are contiguous because the buffer is laid out in reverse-forward order), so that
    NCCL over NVLink/NVSwitch overlaps with the remaining dgrad/wgrad kernels.  Averaging (1/world) is folded into the
    fused Adam's grad_scale.
  * sliding-window inference: windows of a volume are independent -> dealt round-robin to ranks, per-rank accumulation
    of logits, one all-reduce(sum) of the logit volume, uniform averaging, sigmoid, threshold.
"""
import torch
import torch.distributed as dist


def plan_buckets(total: int, boundaries, bucket_elems: int):
    """Split [0, total) into buckets that end on parameter-group boundaries (sorted offsets where a layer's gradients
    are complete) and hold at least `bucket_elems` elements where possible.  Returns [(lo, hi), ...]."""
    out, lo = [], 0
    for b in boundaries:
        if b - lo >= bucket_elems:
            out.append((lo, b))
            lo = b
    if lo < total:
        out.append((lo, total))
    return out


class GradSync:
    """engine.grad_sync hook: .ready(hi) = gradients in flat[0:hi) are final; .finish() = end of backward."""

    def __init__(self, engine, group=None, bucket_mb: float = 25.0):
        self.engine = engine
        self.group = group
        self.world = dist.get_world_size(group)
        self.bucket_elems = int(bucket_mb * 1024 * 1024 / 4)
        self._buckets = None
        self._next = 0
        self._works = []
        self.launched = 0

    def _plan(self):
        eng = self.engine
        ends = sorted({o + n for (o, n) in eng._slots.values()})
        total = eng.flat_grad.numel()
        self._buckets = plan_buckets(total, ends, self.bucket_elems)
        self._total = total

    def ready(self, hi: int):
        if self._buckets is None or self._total != self.engine.flat_grad.numel():
            self._plan()
        while self._next < len(self._buckets) and self._buckets[self._next][1] <= hi:
            lo, b_hi = self._buckets[self._next]
            self._launch(lo, b_hi)
            self._next += 1

    def _launch(self, lo, hi):
        # NCCL's stream waits for the current stream at this point, i.e. for every kernel that produced flat[lo:hi)
        self._works.append(dist.all_reduce(self.engine.flat_grad[lo:hi], op=dist.ReduceOp.SUM, group=self.group,
                                           async_op=True))
        self.launched += 1

    def finish(self):
        if self._buckets is None:
            self._plan()
        while self._next < len(self._buckets):
            lo, hi = self._buckets[self._next]
            self._launch(lo, hi)
            self._next += 1
        for w in self._works:
            w.wait()  # stream-level dependency on the current stream; the host does not block
        self._works = []
        self._next = 0


def make_data_parallel(model, optimizer=None, group=None, bucket_mb: float = 25.0):
    """Turn a UNet3D replica into a data-parallel rank: broadcast rank 0's parameters and BatchNorm buffers, install
    the bucketed gradient all-reduce, fold 1/world into the fused optimizer."""
    eng = model.engine
    dev = next(model.parameters()).device
    eng.prepare(dev)
    dist.broadcast(eng.flat_param, src=0, group=group)
    for b in model.buffers():
        dist.broadcast(b, src=0, group=group)
    eng.external_epoch += 1
    sync = GradSync(eng, group, bucket_mb)
    eng.grad_sync = sync
    if optimizer is not None and hasattr(optimizer, "grad_scale"):
        optimizer.grad_scale = 1.0 / sync.world
    return sync


def shard_batch(n_items: int, rank: int, world: int):
    """contiguous split of a global batch: returns slice(lo, hi) of this rank"""
    per, rem = divmod(n_items, world)
    lo = rank * per + min(rank, rem)
    return slice(lo, lo + per + (1 if rank < rem else 0))


# ------------------------------------------------------------------------------------------------ sliding window
def window_origins(extent: int, window: int, stride: int):
    if extent <= window:
        return [0]
    o = list(range(0, extent - window + 1, stride))
    if o[-1] != extent - window:
        o.append(extent - window)
    return o


def window_schedule(shape, window, stride):
    """[(volume, d0, h0, w0)] in fixed (volume, d, h, w) order for a batch of volumes (N, C, D, H, W)"""
    n, _, D, H, W = shape
    wd, wh, ww = min(window[0], D), min(window[1], H), min(window[2], W)
    sched = []
    for v in range(n):
        for d0 in window_origins(D, wd, stride[0]):
            for h0 in window_origins(H, wh, stride[1]):
                for w0 in window_origins(W, ww, stride[2]):
                    sched.append((v, d0, h0, w0))
    return sched, (wd, wh, ww)


def rank_windows(sched, rank, world):
    """this rank's share of the window schedule: a contiguous block, so that a rank touches as few volumes as possible
    (one volume per rank when the volume count equals the world size: only that volume has to be uploaded there)"""
    return sched[shard_batch(len(sched), rank, world)]


def rank_volumes(shape, window, stride, rank, world):
    """indices of the volumes this rank reads in sliding_window_logits"""
    sched, _ = window_schedule(shape, window, stride)
    return sorted({w[0] for w in rank_windows(sched, rank, world)})


@torch.no_grad()
def sliding_window_logits(model, x, window=(128, 128, 64), stride=(64, 64, 64), rank=0, world=1, group=None,
                          windows_per_launch=9, reduce=True):
    """Average of window logits over a batch of volumes; with world > 1 every rank evaluates its contiguous block of
    the window schedule (rank_windows) and the partial sums are all-reduced.  x: (N, C, D, H, W) on this rank's GPU;
    only the volumes named by rank_volumes() are read on this rank.  Windows are evaluated `windows_per_launch` at a
    time as one batch (measured at 5x256x256x64 on one B200: 222 / 264 / 279 M voxels/s for 1 / 3 / 9 windows per
    launch — fuller waves on the deep levels); the accumulation order is the schedule order either way."""
    model.eval()
    n, _, D, H, W = x.shape
    sched, (wd, wh, ww) = window_schedule(x.shape, window, stride)
    acc = torch.zeros(n, model.n_classes, D, H, W, device=x.device, dtype=torch.float32)
    cnt = torch.zeros(n, 1, D, H, W, device=x.device, dtype=torch.float32)
    mine = rank_windows(sched, rank, world)
    for i in range(0, len(mine), windows_per_launch):
        chunk = mine[i:i + windows_per_launch]
        xb = torch.stack([x[v, :, d0:d0 + wd, h0:h0 + wh, w0:w0 + ww] for (v, d0, h0, w0) in chunk])
        lg = model(xb.contiguous())
        for j, (v, d0, h0, w0) in enumerate(chunk):
            acc[v, :, d0:d0 + wd, h0:h0 + wh, w0:w0 + ww] += lg[j]
    for (v, d0, h0, w0) in sched:  # the count map is deterministic: every rank builds the full one locally
        cnt[v, :, d0:d0 + wd, h0:h0 + wh, w0:w0 + ww] += 1
    if world > 1 and not reduce:
        return acc, cnt  # this rank's partial logit sums (caller reduces) and the full count map
    if world > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    return acc / cnt


@torch.no_grad()
def sliding_window_predict(model, x, window=(128, 128, 64), stride=(64, 64, 64), threshold=0.5, **kw):
    logits = sliding_window_logits(model, x, window, stride, **kw)
    probs = torch.sigmoid(logits)
    return probs, (probs > threshold).float()
