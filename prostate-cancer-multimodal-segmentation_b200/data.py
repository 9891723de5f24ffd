"""Tensor contract of the reference's data pipeline (script/data_loader.py:415-419) without its file IO: batches are
dicts ``{'image': f32 (N,5,D,H,W), 'label': f32 (N,1,D,H,W) in {0,1}, 'case_id': [...]}``.  SimpleITK/NIfTI loading is
out of scope; this module provides the synthetic source used by benchmarks and tests, including the
``zero_fill`` missing-modality semantics (an absent modality is a whole channel of zeros, data_loader.py:320-322)."""
import torch
from torch.utils.data import DataLoader, Dataset

MODALITIES = ["ADC", "DWI", "T2 fs", "T2 not fs", "gaoqing-T2"]


class SyntheticProstateDataset(Dataset):
    def __init__(self, n_cases=8, size=(32, 32, 32), missing_strategy="zero_fill", missing_prob=0.2, seed=1234,
                 indices=None):
        if missing_strategy not in ("zero_fill", "skip", "duplicate"):
            raise ValueError(f"unknown missing_strategy {missing_strategy!r}")
        self.size = tuple(size)
        self.seed = seed
        self.missing_strategy = missing_strategy
        self.missing_prob = missing_prob
        g = torch.Generator().manual_seed(seed)
        # channel 0 (ADC) is always present: the reference keys its case list on ADC (data_loader.py:65-75)
        present = torch.rand(n_cases, 5, generator=g) >= missing_prob
        present[:, 0] = True
        cases = list(range(n_cases))
        if missing_strategy == "skip":
            cases = [c for c in cases if bool(present[c].all())]
        self.present = present
        self.cases = cases if indices is None else [cases[i] for i in indices]

    def __len__(self):
        return len(self.cases)

    def __getitem__(self, i):
        c = self.cases[i]
        g = torch.Generator().manual_seed(self.seed * 7919 + c)
        d, h, w = self.size
        img = torch.randn(5, d, h, w, generator=g)
        zz, yy, xx = torch.meshgrid(torch.linspace(-1, 1, d), torch.linspace(-1, 1, h), torch.linspace(-1, 1, w),
                                    indexing="ij")
        cx = (torch.rand(3, generator=g) - 0.5) * 0.6
        r = 0.25 + 0.2 * torch.rand(1, generator=g).item()
        lab = (((zz - cx[0]) ** 2 + (yy - cx[1]) ** 2 + (xx - cx[2]) ** 2) < r * r).float().unsqueeze(0)
        img = img + 1.5 * lab  # the lesion is visible in every modality
        for m in range(1, 5):
            if not bool(self.present[c, m]):
                if self.missing_strategy == "duplicate":
                    img[m] = img[0]
                else:
                    img[m] = 0.0
        return {"image": img, "label": lab, "case_id": f"case_{c:04d}"}


def resample_case(image, label, target_size):
    """Device half of MultimodalDataset.__getitem__ (script/data_loader.py:294-419) for one case that is already in
    CUDA memory: every modality is resampled to `target_size` (linear), the label with nearest-neighbour and then
    binarised (> 0).  image (5, D, H, W) fp32, label (1, D', H', W') fp32 -> the reference's tensors at target_size.
    Axis order is (D, H, W) throughout (the reference hands its (D,H,W) tuple to ITK's (x,y,z) SetSize, which only
    agrees for cubic targets)."""
    from . import ops
    target_size = tuple(int(v) for v in target_size)
    if tuple(image.shape[-3:]) != target_size:
        image = ops.resample3d(image.float().contiguous(), target_size)
    if tuple(label.shape[-3:]) != target_size:
        label = ops.resample3d(label.float().contiguous(), target_size, nearest=True, binarize=True)
    else:
        label = (label > 0).float()
    return image, label


class DevicePrefetcher:
    """Iterates a loader of batch dicts and hands out the same dicts with their tensors on `device`.

    The host->device copy of batch i+1 is issued on a copy stream as soon as batch i is handed out, so it runs under
    the kernels of step i (the reference's loop, utils/trainer.py:177-181, copies synchronously in front of every
    step).  Host tensors that are not pinned are staged through a pinned buffer first (a pageable source would make
    the copy synchronous).  Two device slots per key rotate; a slot is only overwritten after the compute stream has
    passed the event recorded when its successor was handed out, so no allocator traffic and no host sync."""

    def __init__(self, loader, device, keys=("image", "label")):
        self.loader, self.device, self.keys = loader, torch.device(device), tuple(keys)
        if self.device.type != "cuda":
            raise ValueError("DevicePrefetcher stages batches into CUDA memory; got device " + str(device))
        self.copy_stream = torch.cuda.Stream(self.device)
        self._consumer = torch.cuda.current_stream(self.device)
        self._slots = [{}, {}]
        self._pinned = [{}, {}]
        self._free = [None, None]   # event on the compute stream after which slot i may be overwritten
        self._ready = [None, None]  # event on the copy stream after which slot i's pinned staging buffer is free

    def __len__(self):
        return len(self.loader)

    def _main_stream(self):
        # the consumer's stream: torch's current stream outside the `with torch.cuda.stream(copy_stream)` block
        return self._consumer

    def over(self, loader):
        """iterate another loader through the same staging buffers and copy stream (e.g. one object per trainer,
        re-used for every epoch's loader)"""
        self.loader = loader
        return self

    def _stage(self, batch, i):
        out = dict(batch)
        slot, pinned = self._slots[i], self._pinned[i]
        if self._free[i] is not None:
            self.copy_stream.wait_event(self._free[i])
        with torch.cuda.stream(self.copy_stream):
            for k in self.keys:
                src = batch[k]
                if src.is_cuda and src.device == self.device:
                    out[k] = src   # already resident: nothing to stage
                    continue
                if not src.is_pinned():
                    buf = pinned.get(k)
                    if buf is None or buf.shape != src.shape or buf.dtype != src.dtype:
                        buf = pinned[k] = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
                    if self._ready[i] is not None:
                        self._ready[i].synchronize()
                    buf.copy_(src)
                    src = buf
                dst = slot.get(k)
                if dst is None or dst.shape != src.shape or dst.dtype != src.dtype:
                    dst = slot[k] = torch.empty(src.shape, dtype=src.dtype, device=self.device)
                    # the slot is allocated under the copy stream but read by the compute stream: tell the caching
                    # allocator, so the block is not handed out again while a step still reads it
                    dst.record_stream(self._main_stream())
                dst.copy_(src, non_blocking=True)
                out[k] = dst
            ready = self._ready[i] = self.copy_stream.record_event()
        return out, ready

    def __iter__(self):
        it = iter(self.loader)
        i = 0
        try:
            nxt = self._stage(next(it), i)
        except StopIteration:
            return
        while nxt is not None:
            cur, ready = nxt
            main = self._consumer = torch.cuda.current_stream(self.device)
            main.wait_event(ready)
            # the other slot was handed out one iteration ago: everything queued on the compute stream so far is what
            # read it, so it may be overwritten once the compute stream gets here
            self._free[1 - i] = main.record_event()
            try:
                nxt = self._stage(next(it), 1 - i)
            except StopIteration:
                nxt = None
            yield cur
            i = 1 - i


class AsyncScalarReader:
    """Device->host read of one scalar per step without stalling the launch queue: every pushed 0-dim CUDA tensor is
    copied into a pinned slot on the compute stream and collected `depth` pushes later, when its copy has long
    finished (the reference's `loss.item()` per step, utils/trainer.py:188, drains the GPU every step).
    `values` holds every scalar, in order, once `finish()` has returned."""

    def __init__(self, depth=2):
        self.depth = depth
        self.buf = torch.empty(depth, dtype=torch.float32, pin_memory=True)
        self.events = [None] * depth
        self.values = []
        self._n = 0

    def _collect(self, i):
        self.events[i].synchronize()
        self.values.append(float(self.buf[i]))
        self.events[i] = None

    def push(self, scalar):
        i = self._n % self.depth
        if self.events[i] is not None:
            self._collect(i)
        self.buf[i:i + 1].copy_(scalar.detach().reshape(1), non_blocking=True)
        self.events[i] = torch.cuda.current_stream(scalar.device).record_event()
        self._n += 1

    def finish(self):
        for k in range(self.depth):
            i = (self._n + k) % self.depth
            if self.events[i] is not None:
                self._collect(i)
        return self.values


def get_dataloader(data_dir=None, batch_size=2, shuffle=True, modalities=None, missing_strategy="zero_fill",
                   target_size=(128, 128, 128), num_workers=0, is_training=True, data_type="BPH", indices=None,
                   n_cases=8, seed=1234):
    """same signature as script/data_loader.py:421-423 (+ n_cases/seed); data_dir is ignored (synthetic volumes)"""
    ds = SyntheticProstateDataset(n_cases=n_cases, size=target_size, missing_strategy=missing_strategy, seed=seed,
                                  indices=indices)
    # the shuffle order comes from a seeded generator: data-parallel ranks iterate identical global batches and each
    # takes its share (parallel.ShardedLoader)
    return DataLoader(ds, batch_size=batch_size, shuffle=shuffle and is_training, num_workers=num_workers,
                      pin_memory=True, drop_last=False, generator=torch.Generator().manual_seed(seed + 17))


def get_kfold_splits(n_cases, n_splits=5, seed=42):
    """KFold(shuffle=True, random_state=42) index pairs as script/data_loader.py:468-497, JSON-safe lists"""
    from sklearn.model_selection import KFold
    kf = KFold(n_splits=n_splits, shuffle=True, random_state=seed)
    return [(tr.tolist(), va.tolist()) for tr, va in kf.split(list(range(n_cases)))]
