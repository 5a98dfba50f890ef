"""Tile loaders for TEMPO radiance tiles — the reference's `TEMPODataLoader.get_dataloader(...)` contract
(src/tempo_data.py:112-146) and `load_normalization_stats` (src/tempo_data.py:149-170).

On disk a tile file is a `[64, 64, 64, 1028]` fp32 tensor (64 tiles, channels last; written by
src/scripts/prepare_tempo_tiles.py:189-200). Each yielded sample is a `[1028, 64, 64]` tensor like the reference's
(src/tempo_data.py:98-99) — but it is a *view* of the channels-last tile (no HWC->CHW copy per sample); the default
collate then produces the `[B, 1028, 64, 64]` batch the model API expects.

Sampling semantics follow the reference: an unbounded stream; tiles are drawn uniformly without replacement from a
shuffle pool that is topped up with all tiles of a randomly chosen file whenever it drops below `min_buffer_size`.
"""
import glob
from pathlib import Path

import numpy as np
import torch
from tqdm import tqdm


class RandomBuffer:
    """Pool with O(1) uniform draw-and-remove (swap with the last element)."""

    def __init__(self):
        self.buffer = []

    def put(self, item):
        self.buffer.append(item)

    def get(self):
        if not self.buffer:
            raise IndexError("Buffer is empty")
        i = np.random.randint(0, len(self.buffer))
        self.buffer[i], self.buffer[-1] = self.buffer[-1], self.buffer[i]
        return self.buffer.pop()

    def __len__(self):
        return len(self.buffer)


class TEMPODataset(torch.utils.data.IterableDataset):
    """Infinite stream of spectral tiles, `[C, H, W]` each."""

    def __init__(self, data_dir: str, min_buffer_size: int = 200, verbose: bool = True):
        self.data_dir = Path(data_dir)
        self.min_buffer_size = min_buffer_size
        self.verbose = verbose
        self.files = sorted(glob.glob(str(self.data_dir / "*.pt")))
        if not self.files:
            raise ValueError(f"No .pt files found in {data_dir}")
        self.buffer = RandomBuffer()
        bar = tqdm(total=min_buffer_size, desc="Loading initial buffer") if verbose else None
        self._top_up(bar)
        if bar is not None:
            bar.close()
            print(f"Loaded dataset from {data_dir} with {len(self.files)} files")

    def load_file(self, file_idx: int):
        tiles = torch.load(self.files[file_idx], weights_only=False)
        tiles = tiles.cpu()
        if tiles.dim() == 3:          # a single [H, W, C] tile
            self.buffer.put(tiles)
        else:                         # [N, H, W, C]
            for t in tiles.unbind(0):
                self.buffer.put(t)

    def _top_up(self, bar=None):
        while len(self.buffer) < self.min_buffer_size:
            self.load_file(np.random.randint(0, len(self.files)))
            if bar is not None:
                bar.n = len(self.buffer)
                bar.refresh()

    def get_data(self):
        tile = self.buffer.get()
        self._top_up()
        return tile.permute(2, 0, 1) if tile.dim() == 3 else tile   # [C, H, W] view, channels-last strides

    def __iter__(self):
        while True:
            yield self.get_data()


class TEMPODataLoader:
    """Simplified data loader for TEMPO tiles."""

    @staticmethod
    def get_dataloader(data_dir: str, batch_size: int = 16, num_workers: int = 4, min_buffer_size: int = 200,
                       verbose: bool = True) -> torch.utils.data.DataLoader:
        dataset = TEMPODataset(data_dir=data_dir, min_buffer_size=min_buffer_size, verbose=verbose)
        return torch.utils.data.DataLoader(dataset, batch_size=batch_size, num_workers=num_workers, pin_memory=True)


def load_normalization_stats(stats_dir: str) -> tuple:
    """(mean_spectrum, std_spectrum) from <stats_dir>/mean_spectrum.pt and std_spectrum.pt."""
    stats_dir = Path(stats_dir)
    out = []
    for name, what in (("mean_spectrum.pt", "Mean"), ("std_spectrum.pt", "Std")):
        path = stats_dir / name
        if not path.exists():
            raise FileNotFoundError(f"{what} file not found: {path}")
        out.append(torch.load(path, weights_only=False))
    return tuple(out)


class DevicePrefetcher:
    """Feeds host batches to the GPU one step ahead: the host->device copy of batch i+1 runs on its own CUDA stream
    while step i computes (pinned host memory, two device buffers). Works for tensor batches and dict batches
    (the L2 loader). Yields device tensors that are safe to use on the current stream.

        for batch in DevicePrefetcher(loader, device):
            trainer.train_step(batch)
    """

    def __init__(self, loader, device, dtype=torch.float32):
        self.loader = loader
        self.device = torch.device(device)
        self.dtype = dtype
        self.stream = torch.cuda.Stream(device=self.device)
        self.h2d_bytes = 0

    def _to_device(self, batch):
        if torch.is_tensor(batch):
            if not batch.is_pinned() and batch.device.type == "cpu":
                batch = batch.pin_memory()
            self.h2d_bytes += batch.numel() * batch.element_size()
            return batch.to(self.device, non_blocking=True)
        if isinstance(batch, dict):
            return {k: self._to_device(v) for k, v in batch.items()}
        return batch

    def _record(self, batch):
        cur = torch.cuda.current_stream(self.device)
        for t in (batch.values() if isinstance(batch, dict) else [batch]):
            if torch.is_tensor(t):
                t.record_stream(cur)

    def __iter__(self):
        it = iter(self.loader)

        def fetch():
            try:
                b = next(it)
            except StopIteration:
                return None, None
            with torch.cuda.stream(self.stream):
                d = self._to_device(b)
                ev = torch.cuda.Event()
                ev.record(self.stream)
            return d, ev

        nxt, ev = fetch()
        while nxt is not None:
            torch.cuda.current_stream(self.device).wait_event(ev)
            cur = nxt
            self._record(cur)
            nxt, ev = fetch()
            yield cur
