"""Multi-GPU paths (new capabilities required by BASELINE.json; the reference has no distributed code):

  * data-parallel training: one process per GPU, identical replicas, rank-local BatchNorm statistics and loss; the only
    exchange is the gradient all-reduce, issued per ~25 MB bucket of the flat gradient buffer as soon as backward has
    completed that range (buckets are contiguous because the buffer is laid out in reverse-forward order), so that
    NCCL over NVLink/NVSwitch overlaps with the remaining dgrad/wgrad kernels.  Averaging (1/world) is folded into the
    fused Adam's grad_scale.
  * sliding-window inference: windows of a volume are independent -> dealt round-robin to ranks, per-rank accumulation
    of logits, one all-reduce(sum) of the logit volume, uniform averaging, sigmoid, threshold.
"""
import os

import torch
import torch.distributed as dist

from . import ops


# ------------------------------------------------------------------------------------------------ process group
def dist_info(group=None):
    """(rank, world) of the initialised process group, (0, 1) without one"""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def init_distributed(backend=None):
    """One process per GPU, as launched by `python -m torch.distributed.run ...` (RANK / LOCAL_RANK / WORLD_SIZE /
    MASTER_* in the environment).  Binds this process to its GPU and creates the process group (NCCL over
    NVLink/NVSwitch; B200_DIST_BACKEND=gloo for the one-GPU test rig, where several ranks share a device).
    Returns (rank, world, device); a plain single-process run returns (0, 1, current cuda device)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    ndev = torch.cuda.device_count()
    dev = torch.device("cuda", local % max(ndev, 1))
    torch.cuda.set_device(dev)
    if world > 1 and not dist.is_initialized():
        backend = backend or os.environ.get("B200_DIST_BACKEND", "nccl")
        if backend == "nccl":
            dist.init_process_group("nccl", device_id=dev)
        else:
            dist.init_process_group(backend)
    rank, world = dist_info()
    return rank, world, dev


def backend_is_nccl(group=None) -> bool:
    return dist.is_initialized() and dist.get_backend(group) == "nccl"


def shutdown_distributed():
    if dist.is_available() and dist.is_initialized():
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


def all_mean(value: float, device, group=None) -> float:
    """mean over ranks of a host scalar (epoch losses: every rank must see the same number, it drives the LR
    scheduler and early stopping)"""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(t.item()) / dist.get_world_size(group)


def sync_buffers(model, group=None):
    """BatchNorm running statistics are rank-local during training; before validation and before a checkpoint every
    rank takes rank 0's (DistributedDataParallel's broadcast_buffers behaviour)"""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        for b in model.buffers():
            dist.broadcast(b, src=0, group=group)


def shard_dict_batch(batch, rank, world, keys=("image", "label")):
    """this rank's contiguous share of a global batch dict (tensors under `keys`, lists alongside); None when the
    batch holds fewer samples than ranks (skipped on every rank alike: a rank without samples could not take part
    in the gradient exchange)"""
    n = batch[keys[0]].shape[0]
    if world == 1:
        return batch
    if n < world:
        return None
    sl = shard_batch(n, rank, world)
    out = {}
    for k, v in batch.items():
        if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == n:
            out[k] = v[sl]
        elif isinstance(v, (list, tuple)) and len(v) == n:
            out[k] = list(v[sl])
        else:
            out[k] = v
    return out


class ShardedLoader:
    """iterates a loader of GLOBAL batches and yields this rank's share of each (every rank must iterate the same
    batches in the same order: data.get_dataloader shuffles with a seeded generator for that reason)"""

    def __init__(self, loader, rank, world):
        self.loader, self.rank, self.world = loader, rank, world

    def __len__(self):
        return len(self.loader)

    def __iter__(self):
        for batch in self.loader:
            part = shard_dict_batch(batch, self.rank, self.world)
            if part is not None:
                yield part


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


def window_cover(shape, window, stride):
    """per-axis number of windows covering each coordinate, concatenated [D + H + W] (int32): the windows of a volume
    are the product of the per-axis origin lists, so a voxel is covered by cover[d] * cover[D+h] * cover[D+H+w]"""
    _, _, D, H, W = shape
    out = []
    for extent, win, st in ((D, window[0], stride[0]), (H, window[1], stride[1]), (W, window[2], stride[2])):
        win = min(win, extent)
        c = [0] * extent
        for o in window_origins(extent, win, st):
            for i in range(o, o + win):
                c[i] += 1
        out += c
    return torch.tensor(out, dtype=torch.int32)


def rank_windows(sched, rank, world):
    """this rank's share of the window schedule: a contiguous block, so that a rank touches as few volumes as possible
    (one volume per rank when the volume count equals the world size: only that volume has to be uploaded there)"""
    return sched[shard_batch(len(sched), rank, world)]


def rank_volumes(shape, window, stride, rank, world):
    """indices of the volumes this rank reads in sliding_window_logits"""
    sched, _ = window_schedule(shape, window, stride)
    return sorted({w[0] for w in rank_windows(sched, rank, world)})


def volume_owners(shape, window, stride, world):
    """{volume: (owner rank, [ranks that evaluate some of its windows])}; the owner is the lowest such rank and is
    the one that ends up with the volume's result"""
    sched, _ = window_schedule(shape, window, stride)
    touch = {}
    for r in range(world):
        for w in rank_windows(sched, r, world):
            touch.setdefault(w[0], []).append(r)
    return {v: (min(rs), sorted(set(rs))) for v, rs in touch.items()}


def owned_volumes(shape, window, stride, rank, world):
    return sorted(v for v, (o, _) in volume_owners(shape, window, stride, world).items() if o == rank)


@torch.no_grad()
def sliding_window_logits(model, x, window=(128, 128, 64), stride=(64, 64, 64), rank=0, world=1, group=None,
                          windows_per_launch=9, reduce=True, exchange="owner", threshold=0.5, want=("logits",)):
    """Uniform average of window logits over a batch of volumes x (N, C, D, H, W) fp32 on this rank's GPU.

    With world > 1 every rank evaluates a contiguous block of the window schedule (rank_windows; only the volumes
    named by rank_volumes() are read here).  A volume whose windows all sit on one rank never leaves that rank; a
    volume split over several ranks has its partial logit sums reduced to its owner (`exchange="owner"`, the lowest
    rank that evaluates one of its windows) — the only exchange of the path.  `exchange="all"` all-reduces every
    volume instead, so that every rank holds the full result.  The returned tensors are valid for owned_volumes()
    (all volumes with world == 1 or exchange="all"), zeros elsewhere.

    Windows are evaluated `windows_per_launch` at a time as one batch (fuller waves on the deep levels); gather,
    accumulation (schedule order, no atomics), averaging, sigmoid and threshold are three small kernels
    (ops.window_gather / window_accumulate / window_finalize).  want: any of "logits", "probs", "mask"; returns them
    in that order (a single tensor when one is asked for).  reduce=False returns this rank's partial sums and the
    cover table instead (caller reduces)."""
    model.eval()
    if x.dtype != torch.float32 or not x.is_contiguous():
        x = x.float().contiguous()
    n, _, D, H, W = x.shape
    sched, win = window_schedule(x.shape, window, stride)
    acc = torch.zeros(n, model.n_classes, D, H, W, device=x.device, dtype=torch.float32)
    mine = rank_windows(sched, rank, world)
    if mine:
        org = torch.tensor(mine, dtype=torch.int32).pin_memory().to(x.device, non_blocking=True)
        for i in range(0, len(mine), windows_per_launch):
            chunk = mine[i:i + windows_per_launch]
            o = org[i:i + len(chunk)]
            lg = model(ops.window_gather(x, o, win))
            v_lo, v_hi = chunk[0][0], chunk[-1][0]
            ops.window_accumulate(lg, o, acc, v_lo, v_hi - v_lo + 1)
    cover = window_cover(x.shape, win, stride).to(x.device)
    if world > 1 and not reduce:
        return acc, cover
    if world > 1:
        if exchange == "all":
            dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
        else:
            for v, (owner, ranks) in sorted(volume_owners(x.shape, window, stride, world).items()):
                # every rank runs the same sequence of collectives: one reduce per split volume
                if len(ranks) > 1:
                    dist.reduce(acc[v], dst=owner, op=dist.ReduceOp.SUM, group=group)
                    if rank != owner:
                        acc[v].zero_()
    probs = torch.empty_like(acc) if "probs" in want else None
    mask = torch.empty_like(acc) if "mask" in want else None
    ops.window_finalize(acc, cover, threshold, probs, mask)
    res = tuple({"logits": acc, "probs": probs, "mask": mask}[k] for k in want)
    return res[0] if len(res) == 1 else res


@torch.no_grad()
def sliding_window_predict(model, x, window=(128, 128, 64), stride=(64, 64, 64), threshold=0.5, **kw):
    """(probabilities, 0/1 float mask) of the window-averaged logits; see sliding_window_logits"""
    return sliding_window_logits(model, x, window, stride, threshold=threshold, want=("probs", "mask"), **kw)
