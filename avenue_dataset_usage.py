"""Drop-in for the ``avenue_dataset_usage`` module imported by avenue_training_script1.py / script2.py (s1:19, 86-92, 303;
s2:357-365) and absent from the reference.

    train_loader, test_loader = create_avenue_dataloaders(dataset_path, batch_size=4, num_workers=2, clip_length=8,
                                                          frame_size=(64, 64))

Each loader yields ``(videos float32 (B,3,T,H,W) in [0,1], labels (B,))``.  ``dataset_path`` is either a directory in
the CUHK Avenue frame layout (``<path>/{training,testing}/frames/<video>/*.jpg`` or ``<path>/{train,test}/<video>/*``),
or ``"synthetic"`` / ``"synthetic:<n_clips>"`` for seeded Avenue-shaped clips (no dataset is available offline).
Host-side data plumbing only: the hot path starts at the ``.to(device)`` of the batch."""
from __future__ import annotations

import glob
import os

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset

_IMG = (".jpg", ".jpeg", ".png", ".tif", ".bmp")


class AvenueFramesDataset(Dataset):
    """Sliding clips of ``clip_length`` frames (stride clip_length // 2) over per-video frame folders."""

    def __init__(self, split_dir, clip_length=8, frame_size=(64, 64), label=0):
        self.clip_length, self.frame_size, self.label = clip_length, tuple(frame_size), label
        self.clips = []
        for vdir in sorted(d for d in glob.glob(os.path.join(split_dir, "*")) if os.path.isdir(d)):
            frames = sorted(f for f in glob.glob(os.path.join(vdir, "*")) if f.lower().endswith(_IMG))
            step = max(1, clip_length // 2)
            for s in range(0, len(frames) - clip_length + 1, step):
                self.clips.append(frames[s:s + clip_length])

    def __len__(self):
        return len(self.clips)

    def __getitem__(self, i):
        import cv2
        out = np.empty((self.clip_length, self.frame_size[1], self.frame_size[0], 3), dtype=np.float32)
        for t, path in enumerate(self.clips[i]):
            img = cv2.imread(path)
            if img is None:
                raise IOError(f"cannot read frame {path}")
            img = cv2.cvtColor(cv2.resize(img, self.frame_size), cv2.COLOR_BGR2RGB)
            out[t] = img.astype(np.float32) / 255.0
        return torch.from_numpy(out).permute(3, 0, 1, 2).contiguous(), torch.tensor(float(self.label))


class SyntheticAvenueDataset(Dataset):
    """Seeded Avenue-shaped clips: uniform noise times a per-clip brightness."""

    def __init__(self, n_clips, clip_length=8, frame_size=(64, 64), seed=1234):
        self.n, self.T, self.size, self.seed = n_clips, clip_length, tuple(frame_size), seed

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed + i)
        x = torch.rand(3, self.T, self.size[1], self.size[0], generator=g) * torch.rand(1, generator=g)
        return x, torch.tensor(0.0)


def _split_dir(root, names):
    for n in names:
        for cand in (os.path.join(root, n, "frames"), os.path.join(root, n)):
            if os.path.isdir(cand):
                return cand
    raise FileNotFoundError(f"no {names} split under {root}")


def create_avenue_dataloaders(dataset_path, batch_size=4, num_workers=2, clip_length=8, frame_size=(64, 64)):
    if str(dataset_path).startswith("synthetic"):
        n = int(str(dataset_path).split(":")[1]) if ":" in str(dataset_path) else 64
        train = SyntheticAvenueDataset(n, clip_length, frame_size, seed=1234)
        test = SyntheticAvenueDataset(max(batch_size, n // 4), clip_length, frame_size, seed=99991)
        num_workers = 0
    else:
        train = AvenueFramesDataset(_split_dir(dataset_path, ("training", "train")), clip_length, frame_size)
        test = AvenueFramesDataset(_split_dir(dataset_path, ("testing", "test")), clip_length, frame_size)
    kw = dict(batch_size=batch_size, num_workers=num_workers, pin_memory=True, drop_last=False)
    return DataLoader(train, shuffle=True, **kw), DataLoader(test, shuffle=False, **kw)
