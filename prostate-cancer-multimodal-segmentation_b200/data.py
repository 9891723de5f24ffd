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


def get_dataloader(data_dir=None, batch_size=2, shuffle=True, modalities=None, missing_strategy="zero_fill",
                   target_size=(128, 128, 128), num_workers=0, is_training=True, data_type="BPH", indices=None,
                   n_cases=8, seed=1234):
    """same signature as script/data_loader.py:421-423 (+ n_cases/seed); data_dir is ignored (synthetic volumes)"""
    ds = SyntheticProstateDataset(n_cases=n_cases, size=target_size, missing_strategy=missing_strategy, seed=seed,
                                  indices=indices)
    return DataLoader(ds, batch_size=batch_size, shuffle=shuffle and is_training, num_workers=num_workers,
                      pin_memory=True, drop_last=False)


def get_kfold_splits(n_cases, n_splits=5, seed=42):
    """KFold(shuffle=True, random_state=42) index pairs as script/data_loader.py:468-497, JSON-safe lists"""
    from sklearn.model_selection import KFold
    kf = KFold(n_splits=n_splits, shuffle=True, random_state=seed)
    return [(tr.tolist(), va.tolist()) for tr, va in kf.split(list(range(n_cases)))]
