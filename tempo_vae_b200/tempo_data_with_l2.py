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

from .tempo_data import RandomBuffer

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
