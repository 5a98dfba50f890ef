"""Tile loaders with aligned L2 products — the reference's `TEMPODataLoaderWithL2.get_dataloader(...)` contract
(src/tempo_data_with_l2.py:139-176).

Directory layout (src/scripts/prepare_tempo_tiles_with_l2.py): `<data_dir>/<split>/*.pt` hold `[64, 64, 64, 1028]`
spectral tiles, `<data_dir>/<split>/l2_<PRODUCT>/<same name>.pt` hold `[64, 64, 64]` targets (NaN = invalid pixel)
for PRODUCT in NO2, O3TOT, HCHO, CLDO4. Each sample is a dict {'spectral': [1028,64,64], '<PRODUCT>': [64,64]}.
"""
from pathlib import Path
from typing import Dict

import numpy as np
import torch
from torch.utils.data import DataLoader, IterableDataset
from tqdm import tqdm

from . import ops
from .tempo_data import RandomBuffer, epoch_shard

L2_PRODUCTS = ('NO2', 'O3TOT', 'HCHO', 'CLDO4')


class TEMPODatasetWithL2(IterableDataset):
    def __init__(self, data_dir: str, split: str = 'train', min_buffer_size: int = 200, verbose: bool = True):
        self.data_dir = Path(data_dir) / split
        self.min_buffer_size = min_buffer_size
        self.verbose = verbose
        if not self.data_dir.exists():
            raise FileNotFoundError(f"FATAL: Data directory not found: {self.data_dir}")
        self.tile_files = sorted(self.data_dir.glob("*.pt"))
        if not self.tile_files:
            raise ValueError(f"FATAL: No .pt files found in {self.data_dir}")
        self.l2_products = list(L2_PRODUCTS)
        self.l2_dirs = {}
        for product in self.l2_products:
            d = self.data_dir / f'l2_{product}'
            if not d.exists():
                raise FileNotFoundError(f"FATAL: L2 directory not found: {d}")
            self.l2_dirs[product] = d
        self.buffer = RandomBuffer()
        self.tiles_per_file = 64
        bar = tqdm(total=min_buffer_size, desc=f"Loading initial buffer ({split})") if verbose else None
        self._top_up(bar)
        if bar is not None:
            bar.close()
            print(f"Loaded {split} dataset with {len(self.tile_files)} files")

    def load_file(self, file_idx: int):
        path = self.tile_files[file_idx]
        spectral = torch.load(path, weights_only=True).cpu()
        l2 = {}
        for product in self.l2_products:
            p = self.l2_dirs[product] / path.name
            if not p.exists():
                raise FileNotFoundError(f"FATAL: L2 file not found: {p}")
            l2[product] = torch.load(p, weights_only=True).cpu()
        n = min(self.tiles_per_file, spectral.shape[0])
        for i in range(n):
            tile = spectral[i]
            if tile.dim() == 3 and tile.shape[-1] == spectral.shape[-1]:
                tile = tile.permute(2, 0, 1)                      # [C, H, W] view of the channels-last tile
            sample: Dict[str, torch.Tensor] = {'spectral': tile}
            for product in self.l2_products:
                sample[product] = l2[product][i]
            self.buffer.put(sample)

    def _top_up(self, bar=None):
        while len(self.buffer) < self.min_buffer_size:
            self.load_file(np.random.randint(0, len(self.tile_files)))
            if bar is not None:
                bar.n = len(self.buffer)
                bar.refresh()

    def get_data(self):
        sample = self.buffer.get()
        self._top_up()
        return sample

    def __iter__(self):
        while True:
            yield self.get_data()


class TEMPODataLoaderWithL2:
    """DataLoader wrapper for TEMPO tiles with L2 products."""

    @staticmethod
    def get_dataloader(data_dir: str, split: str = 'train', batch_size: int = 32, num_workers: int = 4,
                       min_buffer_size: int = 200, verbose: bool = True) -> DataLoader:
        dataset = TEMPODatasetWithL2(data_dir=data_dir, split=split, min_buffer_size=min_buffer_size, verbose=verbose)
        return DataLoader(dataset, batch_size=batch_size, num_workers=num_workers, pin_memory=True,
                          persistent_workers=(num_workers > 0))


class DeviceTileCacheWithL2:
    """`tempo_data.DeviceTileCache` for the L2-supervised variant: the spectral tiles resident in HBM as channels-last
    bf16, the four product targets next to them as fp32 `[n, 64, 64]` (NaN = invalid pixel, kept as is). Batches are the
    dicts `L2SupervisedTrainer.train_step` / `VAEWithL2Supervision.compute_loss` take:
    `{'spectral': [B, C, H, W] bf16 channels-last view, 'NO2': [B, H, W], ...}`.

        cache = DeviceTileCacheWithL2.from_dir(data_dir, 'train', device)
        for batch in cache.batches(256, seed=0, rank=rank, world=world):
            trainer.train_step_device(batch)
    """

    def __init__(self, spectral_cache, targets: Dict[str, torch.Tensor]):
        self.spectral = spectral_cache
        self.targets = targets

    def __len__(self):
        return len(self.spectral)

    @classmethod
    def from_dir(cls, data_dir: str, split: str, device, max_tiles=None):
        from .tempo_data import DeviceTileCache
        root = Path(data_dir) / split
        files = sorted(root.glob("*.pt"))
        if not files:
            raise ValueError(f"FATAL: No .pt files found in {root}")
        spectral = DeviceTileCache.from_dir(str(root), device, max_tiles=max_tiles)
        targets = {}
        for product in L2_PRODUCTS:
            d = root / f'l2_{product}'
            if not d.exists():
                raise FileNotFoundError(f"FATAL: L2 directory not found: {d}")
            parts = []
            for f in files:
                p = d / f.name
                if not p.exists():
                    raise FileNotFoundError(f"FATAL: L2 file not found: {p}")
                t = torch.load(p, weights_only=True)
                parts.append(t.unsqueeze(0) if t.dim() == 2 else t)
            tg = torch.cat(parts, 0)[:len(spectral)].to(device, dtype=torch.float32)
            if tg.shape[0] != len(spectral):
                raise ValueError(f"{product}: {tg.shape[0]} targets for {len(spectral)} spectral tiles")
            targets[product] = tg.contiguous()
        return cls(spectral, targets)

    def batches(self, batch_size: int, seed: int = 0, epochs=None, rank: int = 0, world: int = 1):
        sp = self.spectral
        if sp.n < batch_size * world:
            raise ValueError(f"{sp.n} cached tiles cannot fill a batch of {batch_size} on {world} ranks")
        g = torch.Generator().manual_seed(seed)
        epoch = 0
        while epochs is None or epoch < epochs:
            perm = epoch_shard(torch.randperm(sp.n, generator=g), batch_size, rank, world).to(sp.device)
            for i in range(0, perm.numel(), batch_size):
                idx = perm[i:i + batch_size]
                batch = {'spectral': sp.gather(idx)[..., :sp.C].permute(0, 3, 1, 2)}
                for product, tg in self.targets.items():
                    out = torch.empty((idx.numel(),) + tuple(tg.shape[1:]), dtype=tg.dtype, device=tg.device)
                    ops.gather_rows(tg, idx, out)
                    batch[product] = out
                yield batch
            epoch += 1
