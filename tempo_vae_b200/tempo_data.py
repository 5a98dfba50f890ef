"""Tile loaders for TEMPO radiance tiles — the reference's `TEMPODataLoader.get_dataloader(...)` contract
(src/tempo_data.py:112-146) and `load_normalization_stats` (src/tempo_data.py:149-170).

On disk a tile file is a `[64, 64, 64, 1028]` fp32 tensor (64 tiles, channels last; written by
src/scripts/prepare_tempo_tiles.py:189-200). Each yielded sample is a `[1028, 64, 64]` tensor like the reference's
(src/tempo_data.py:98-99) — but it is a *view* of the channels-last tile (no HWC->CHW copy per sample); the default
collate then produces the `[B, 1028, 64, 64]` batch the model API expects.

Sampling semantics follow the reference: an unbounded stream; tiles are drawn uniformly without replacement from a
shuffle pool that is topped up with all tiles of a randomly chosen file whenever it drops below `min_buffer_size`.

Feeding a B200. One GPU consumes ~2,300 tiles/s = 39 GB/s of fp32 tiles; Python worker
processes + collate + a pin-memory thread cannot move that, and at 8 GPUs the fp32 NCHW batches saturate the host's
PCIe fabric (measured: 185 GB/s aggregate, e2e efficiency 0.59). The loaders that keep up hold the split ONCE in the
engine's operand format -- channels-last bf16 rows, pitch 1032 (the on-disk tiles are already channels-last):
  * `HostTileStore`   : the split in PINNED host memory; a batch is assembled directly in HBM by one asynchronous DMA
                        per tile on a copy stream, double-buffered one step ahead (no host-side gather/collate at all);
                        2.16 GB per 256-tile step over PCIe instead of 4.31 GB
  * `DeviceTileCache` : the split in HBM (the January train split is 18 GB of 180); batches gathered on the device
Both yield NCHW-SHAPED `[B, C, H, W]` bf16 views with channels-last strides that the engine consumes in place, and
both come from `TEMPODataLoader.get_host_store(...)` / `.get_device_cache(...)` next to the reference's
`get_dataloader(...)`.
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

    @staticmethod
    def get_host_store(data_dir: str, max_tiles=None, verbose: bool = True) -> "HostTileStore":
        """The split as a pinned-host bf16 channels-last tile store (see HostTileStore)."""
        return HostTileStore.from_dir(data_dir, max_tiles=max_tiles, verbose=verbose)

    @staticmethod
    def get_device_cache(data_dir: str, device, max_tiles=None, verbose: bool = True) -> "DeviceTileCache":
        """The split resident in HBM (see DeviceTileCache)."""
        return DeviceTileCache.from_dir(data_dir, device, max_tiles=max_tiles, verbose=verbose)


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


class DeviceTileCache:
    """The tile set resident in HBM in the layout the conv kernels read.

    The reference streams fp32 tiles from worker processes, permutes HWC -> CHW, collates and copies 4.3 GB per
    256-sample step to the GPU (src/tempo_data.py:34-146). The whole January train split is 18 GB as bf16, a tenth of
    one B200's HBM: here every tile is cast ONCE to channels-last bf16 rows (`tvae_nhwc_f32_to_nhwc_bf16`, pitch
    rounded up to 8 channels for TMA) and batches are gathered on the device. A batch is yielded as an NCHW-SHAPED
    `[B, C, H, W]` bf16 view with channels-last strides, which `get_loss` / `encode` consume in place (no transpose,
    no host->device traffic in the step):

        cache = DeviceTileCache.from_dir(train_dir, device)
        for x in cache.batches(256, seed=0, rank=rank, world=world):
            trainer.train_step_device(x)

    Sampling: epoch-wise random permutation without replacement (the reference draws without replacement from a pool
    refilled file by file; both visit every tile equally often). `rank/world` shard each epoch's permutation.
    """

    def __init__(self, device, H: int, W: int, C: int, capacity: int):
        from . import ops
        self.device = torch.device(device)
        self.H, self.W, self.C = H, W, C
        self.pitch = ops.round_up(C, 8)
        self.data = torch.zeros((capacity, H, W, self.pitch), dtype=torch.bfloat16, device=self.device)
        self.n = 0

    def __len__(self):
        return self.n

    def add(self, tiles: torch.Tensor, chunk: int = 64):
        """tiles: [n, H, W, C] or [H, W, C] fp32, channels last (the on-disk format), on the host or the device."""
        from . import ops
        from ._lib import lib
        if tiles.dim() == 3:
            tiles = tiles.unsqueeze(0)
        if tuple(tiles.shape[1:]) != (self.H, self.W, self.C):
            raise ValueError(f"tile shape {tuple(tiles.shape[1:])} does not match the cache ({self.H}, {self.W}, {self.C})")
        if self.n + tiles.shape[0] > self.data.shape[0]:
            raise ValueError("DeviceTileCache capacity exceeded")
        for i in range(0, tiles.shape[0], chunk):
            t = tiles[i:i + chunk].to(self.device, dtype=torch.float32, non_blocking=True).contiguous()
            dst = self.data[self.n:self.n + t.shape[0]]
            with torch.cuda.device(self.device):
                ops.check(lib.tvae_nhwc_f32_to_nhwc_bf16(t.data_ptr(), self.C, t.shape[0] * self.H * self.W, self.C,
                                                         dst.data_ptr(), self.pitch, None, ops._stream()),
                          "tvae_nhwc_f32_to_nhwc_bf16")
            self.n += t.shape[0]

    @classmethod
    def from_dir(cls, data_dir: str, device, max_tiles=None, verbose: bool = False):
        files = sorted(glob.glob(str(Path(data_dir) / "*.pt")))
        if not files:
            raise ValueError(f"No .pt files found in {data_dir}")
        first = torch.load(files[0], weights_only=False)
        if first.dim() == 3:
            first = first.unsqueeze(0)
        per_file, H, W, C = first.shape
        cap = per_file * len(files) if max_tiles is None else min(max_tiles, per_file * len(files))
        cache = cls(device, H, W, C, cap)
        for i, f in enumerate(tqdm(files, desc="Caching tiles on the device") if verbose else files):
            t = first if i == 0 else torch.load(f, weights_only=False)
            if t.dim() == 3:
                t = t.unsqueeze(0)
            room = cap - cache.n
            if room <= 0:
                break
            if t.shape[0] > per_file and max_tiles is None:
                raise ValueError(f"{f} holds {t.shape[0]} tiles, more than the first file ({per_file})")
            cache.add(t[:room])
        return cache

    def gather(self, idx: torch.Tensor) -> torch.Tensor:
        """[len(idx), H, W, pitch] bf16: the tiles `idx` (int64, on the device) copied by `tvae_gather_rows`."""
        from . import ops
        out = torch.empty((idx.numel(), self.H, self.W, self.pitch), dtype=torch.bfloat16, device=self.device)
        ops.gather_rows(self.data, idx, out)
        return out

    def batches(self, batch_size: int, seed: int = 0, epochs=None, rank: int = 0, world: int = 1):
        """Yields [batch_size, C, H, W] bf16 channels-last views; the last partial batch of an epoch is dropped.
        Each yielded batch owns its memory (a device-side gather), so it may be kept across iterations. Every rank
        yields the same number of batches per epoch (see epoch_shard)."""
        if self.n < batch_size * world:
            raise ValueError(f"{self.n} cached tiles cannot fill a batch of {batch_size} on {world} ranks")
        g = torch.Generator().manual_seed(seed)
        epoch = 0
        while epochs is None or epoch < epochs:
            perm = epoch_shard(torch.randperm(self.n, generator=g), batch_size, rank, world).to(self.device)
            for i in range(0, perm.numel(), batch_size):
                yield self.gather(perm[i:i + batch_size])[..., :self.C].permute(0, 3, 1, 2)
            epoch += 1


def epoch_shard(perm: torch.Tensor, batch_size: int, rank: int, world: int) -> torch.Tensor:
    """This rank's tiles of one epoch: the permutation is first truncated to a whole number of GLOBAL batches
    (world * batch_size tiles) and then strided, so every rank gets the same number of full batches -- with
    `perm[rank::world]` alone rank 0 can end up with one batch more than the others when n % world != 0, and the
    ranks would issue different numbers of all-reduces (a hang in NCCL)."""
    usable = (perm.numel() // (world * batch_size)) * world * batch_size
    return perm[:usable][rank::world]


class HostTileStore:
    """The tile set in PINNED host memory in the layout the conv kernels read (channels-last bf16 rows, pitch rounded
    up to 8 channels), cast once at load time. `batches(...)` assembles each batch directly in device memory: one
    asynchronous host->device DMA per tile (8.4 MB each) on a dedicated copy stream into one of three rotating device
    buffers, one batch ahead of the step that is computing -- the "gather" is done by the copy engines, the host only
    enqueues `batch_size` memcpy descriptors (~2 ms). Results are bit-identical to feeding the fp32 tiles: the engine
    rounds its input to bf16 first thing.

        store = TEMPODataLoader.get_host_store(train_dir)            # or HostTileStore.from_dir(...)
        for x in store.batches(256, device, seed=0, rank=rank, world=world):
            trainer.train_step(x)                                    # x: [256, 1028, 64, 64] bf16 view, channels-last

    A yielded batch stays valid until the NEXT-BUT-ONE batch is requested (three buffers); copy it if it must live longer.
    """

    def __init__(self, H: int, W: int, C: int, capacity: int, pin: bool = True):
        self.H, self.W, self.C = H, W, C
        self.pitch = (C + 7) // 8 * 8
        pinned = bool(pin and torch.cuda.is_available())         # allocated pinned directly: no pageable copy of the split
        self.data = torch.zeros((capacity, H, W, self.pitch), dtype=torch.bfloat16, pin_memory=pinned)
        self.n = 0
        self.h2d_bytes = 0

    def __len__(self):
        return self.n

    def add(self, tiles: torch.Tensor):
        """tiles: [n, H, W, C] or [H, W, C], fp32 or bf16, channels last (the on-disk format), host or device."""
        if tiles.dim() == 3:
            tiles = tiles.unsqueeze(0)
        if tuple(tiles.shape[1:]) != (self.H, self.W, self.C):
            raise ValueError(f"tile shape {tuple(tiles.shape[1:])} does not match the store ({self.H}, {self.W}, {self.C})")
        k = tiles.shape[0]
        if self.n + k > self.data.shape[0]:
            raise ValueError("HostTileStore capacity exceeded")
        self.data[self.n:self.n + k, :, :, :self.C].copy_(tiles)          # cast (+ D2H) once, at load time
        self.n += k

    @classmethod
    def from_dir(cls, data_dir: str, max_tiles=None, verbose: bool = False):
        files = sorted(glob.glob(str(Path(data_dir) / "*.pt")))
        if not files:
            raise ValueError(f"No .pt files found in {data_dir}")
        first = torch.load(files[0], weights_only=False)
        if first.dim() == 3:
            first = first.unsqueeze(0)
        per_file, H, W, C = first.shape
        cap = per_file * len(files) if max_tiles is None else min(max_tiles, per_file * len(files))
        store = cls(H, W, C, cap)
        for i, f in enumerate(tqdm(files, desc="Loading tiles into pinned memory") if verbose else files):
            t = first if i == 0 else torch.load(f, weights_only=False)
            if t.dim() == 3:
                t = t.unsqueeze(0)
            room = cap - store.n
            if room <= 0:
                break
            if t.shape[0] > per_file and max_tiles is None:
                raise ValueError(f"{f} holds {t.shape[0]} tiles, more than the first file ({per_file})")
            store.add(t[:room])
        return store

    def batches(self, batch_size: int, device, seed: int = 0, epochs=None, rank: int = 0, world: int = 1,
                n_buffers: int = 3):
        device = torch.device(device)
        if device.type != "cuda":
            raise ValueError("HostTileStore.batches feeds a CUDA device")
        if self.n < batch_size * world:
            raise ValueError(f"{self.n} stored tiles cannot fill a batch of {batch_size} on {world} ranks")
        copy_stream = torch.cuda.Stream(device=device)
        bufs = [torch.empty((batch_size, self.H, self.W, self.pitch), dtype=torch.bfloat16, device=device)
                for _ in range(n_buffers)]
        free = [None] * n_buffers            # event: the consumer is done with buffer k
        tile_bytes = self.H * self.W * self.pitch * 2

        def order():
            g = torch.Generator().manual_seed(seed)
            epoch = 0
            while epochs is None or epoch < epochs:
                perm = epoch_shard(torch.randperm(self.n, generator=g), batch_size, rank, world).tolist()
                for i in range(0, len(perm), batch_size):
                    yield perm[i:i + batch_size]
                epoch += 1

        def issue(k, idx):
            with torch.cuda.stream(copy_stream):
                if free[k] is not None:
                    copy_stream.wait_event(free[k])
                dst = bufs[k]
                for j, t in enumerate(idx):
                    dst[j].copy_(self.data[t], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            self.h2d_bytes += len(idx) * tile_bytes
            return ev

        it = order()
        k = 0
        try:
            pending = (k, issue(k, next(it)))
        except StopIteration:
            return
        while pending is not None:
            cur_k, ev = pending
            nxt = next(it, None)
            k = (cur_k + 1) % n_buffers
            pending = (k, issue(k, nxt)) if nxt is not None else None
            cur = torch.cuda.current_stream(device)
            cur.wait_event(ev)
            yield bufs[cur_k][..., :self.C].permute(0, 3, 1, 2)
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream(device))      # everything the consumer enqueued on this batch
            free[cur_k] = done
