"""Synthetic stand-in for the reference's training dataset (codes/datasets/mul_dataset.py:180-253, loader
codes/datasets/dataloader.py:95-119) so that DDPMMulTrainer.train can be exercised and timed end to end without the
NTU RGB+D files (there is no dataset on the benchmark box).

`SyntheticText2MotionMulDataset.__getitem__` returns exactly what `Text2MotionMulDataset.__getitem__` returns in training mode
— (caption1, caption2, motion1 [91, 263], motion2 [91, 263], m_length, file_id) — and builds the 91 rows the same way: row 0
is the last raw frame (the 4-feature root initialisation), rows 1..90 a window of the raw motion (random shift) or the whole
motion padded with its last frame, z-normalised with (mean, std) / (init_mean, init_std).  Raw motions are seeded noise.
`build_dataloader` mirrors the reference's loader (DistributedSampler per rank, drop_last) with pinned host memory."""
import random

import numpy as np
import torch
from torch.utils import data

CAPTIONS = [
    ("a person pushes the other person", "a person is pushed by the other person"),
    ("a person kicks the other person", "a person is kicked by the other person"),
    ("a person pats the other person on the back", "a person is patted on the back by the other person"),
    ("a person points a finger at the other person", "a person is pointed at by the other person"),
    ("a person hugs the other person", "a person hugs the other person"),
    ("a person gives an object to the other person", "a person receives an object from the other person"),
    ("a person touches the pocket of the other person", "a person has the pocket touched by the other person"),
    ("two people shake hands", "two people shake hands"),
    ("a person walks towards the other person", "a person walks towards the other person"),
    ("a person punches the other person", "a person is punched by the other person"),
]


class SyntheticText2MotionMulDataset(data.Dataset):
    def __init__(self, n_items=1024, dim_pose=263, cap_id=False, with_label=True, times=1, seed=0, min_len=20, max_len=199):
        rs = np.random.RandomState(seed)
        self.cap_id, self.with_label, self.times = cap_id, with_label, times
        self.mean, self.std = np.zeros(dim_pose, np.float32), np.ones(dim_pose, np.float32)
        self.init_mean, self.init_std = np.zeros(4, np.float32), np.ones(4, np.float32)
        self.items = []
        for i in range(n_items):
            n = int(rs.randint(min_len, max_len + 1))            # raw frames incl. the trailing initialisation frame
            motion = rs.standard_normal((2, n, dim_pose)).astype(np.float32)
            self.items.append({"motion": motion, "length": n - 1, "caption": int(rs.randint(0, len(CAPTIONS))),
                               "label": int(rs.randint(0, 2))})

    def real_len(self):
        return len(self.items)

    def __len__(self):
        return self.real_len() * self.times

    def __getitem__(self, item):
        d = self.items[item % self.real_len()]
        motion, m_length = d["motion"], d["length"]
        num_frames = 90
        nframes = motion.shape[1] - 1
        if num_frames > nframes:                                 # :189-194 pad with the last frame
            frame_ix = np.concatenate(([nframes], np.arange(0, nframes), (nframes - 1) * np.ones(num_frames - nframes, dtype=int)))
        else:                                                    # :195-202 random window
            lastone = num_frames - 1
            shift_max = nframes - lastone - 1
            shift = random.randint(0, max(0, shift_max - 1))
            frame_ix = np.concatenate(([nframes], shift + np.arange(0, lastone + 1)))
        motion1, motion2 = motion[0][frame_ix].copy(), motion[1][frame_ix].copy()
        for m in (motion1, motion2):                             # :205-209
            m[1:] = (m[1:] - self.mean) / self.std
            m[0, :4] = (m[0, :4] - self.init_mean) / self.init_std
        c1, c2 = CAPTIONS[d["caption"]]
        if self.cap_id:                                          # :214-216 caption ids, one-element lists
            c1, c2 = [2 * d["caption"]], [2 * d["caption"] + 1]
        if self.with_label and d["label"]:                       # :247-251
            return c1, c2, motion2, motion1, m_length, str(item)
        return c1, c2, motion1, motion2, m_length, str(item)


def build_dataloader(dataset, rank, world_size, samples_per_gpu, drop_last=True, workers_per_gpu=4, shuffle=True, seed=0):
    """codes/datasets/dataloader.py:95-119 with pinned host memory (the reference passes pin_memory=False)."""
    sampler = None
    if world_size > 1:
        sampler = data.distributed.DistributedSampler(dataset, world_size, rank, shuffle=shuffle, seed=seed)
        shuffle = False
    return data.DataLoader(dataset, batch_size=samples_per_gpu, sampler=sampler, num_workers=workers_per_gpu,
                           pin_memory=torch.cuda.is_available(), shuffle=shuffle, drop_last=drop_last,
                           persistent_workers=workers_per_gpu > 0)


class DevicePrefetcher:
    """Iterates a loader one batch ahead of the training step: the two motion tensors and the lengths of batch i+1 travel
    host -> device on a side stream (pinned memory, asynchronous DMA) while batch i trains, so the copy (24 MB at the C4
    batch) never sits between two iterations' kernels.  Captions stay on the host (strings / small id tensors)."""

    def __init__(self, loader, device):
        self.loader, self.device = loader, torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.sampler = getattr(loader, "sampler", None)

    def __len__(self):
        return len(self.loader)

    def _stage(self, it):
        try:
            batch = next(it)
        except StopIteration:
            return None
        c1, c2, m1, m2, lens, fid = batch
        with torch.cuda.stream(self.stream):
            m1 = torch.as_tensor(m1).to(self.device, non_blocking=True)
            m2 = torch.as_tensor(m2).to(self.device, non_blocking=True)
            lens = torch.as_tensor(lens).to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return (c1, c2, m1, m2, lens, fid), ev

    def __iter__(self):
        it = iter(self.loader)
        nxt = self._stage(it)
        while nxt is not None:
            batch, ev = nxt
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            for t in batch[2:5]:
                t.record_stream(cur)
            nxt = self._stage(it)
            yield batch
