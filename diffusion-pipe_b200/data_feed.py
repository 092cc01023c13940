"""Micro-batch feed of the pipeline engine — the integer bookkeeping that must be bit-exact with the reference
(SURVEY.md section 8a rows H2-H6):

  seeded_shuffle             utils/dataset.py:41-45   (`shuffle_with_seed`: Python `random`, state restored)
  BucketBatches              utils/dataset.py:340-396 (`ConcatenatedBatchedDataset`: per-size-bucket global batches,
                             truncated to a multiple of the global batch, sliced per data-parallel rank)
  BatchedDataset             utils/dataset.py:953-1034 (`Dataset.post_init/__len__/__getitem__/_collate`)
  split_batch                utils/dataset.py:1273-1281
  PipelineDataLoader         utils/dataset.py:1302-1435 (prefetch-by-one epoch accounting, resume by skipping)
  SkipFirstNSampler          utils/dataset.py:1438-1449
  get_data_iterator_for_step train.py:167-173

Latent / text-embedding caching, directory scanning and aspect-ratio bucketing that PRODUCE the size-bucket datasets are
outside the hot path (north star: "latent/text-embed caching (utils/cache.py) unchanged"); this module consumes any
objects with `.size_bucket`, `__len__` and `__getitem__` returning feature dicts.

Pinned by tests/golden/datafeed_traces.json, produced by executing the reference's own source text
(tests/golden/make_golden_datafeed.py).
"""
import math
import os
import random
import re
from collections import defaultdict

import numpy as np
import torch

from .pipe import dist


def seeded_shuffle(items, seed=None):
    """In-place shuffle with Python's Mersenne Twister seeded by `seed`; the global RNG state is left untouched."""
    saved = random.getstate()
    try:
        random.seed(seed)
        random.shuffle(items)
    finally:
        random.setstate(saved)


def _interleave_order(group_sizes):
    """[(group, running index within group)] after a seed-0 shuffle of the multiset {group g repeated size_g times}."""
    order = []
    for g, n in enumerate(group_sizes):
        order += [g] * n
    seeded_shuffle(order, 0)
    seen = [0] * len(group_sizes)
    out = []
    for g in order:
        out.append((g, seen[g]))
        seen[g] += 1
    return out


class BucketBatches:
    """All datasets of ONE size bucket, served as per-rank slices of global batches."""

    def __init__(self, datasets):
        self.datasets = datasets
        self.ready = False

    def post_init(self, global_batch_size, global_batch_size_image, data_parallel_rank, data_parallel_world_size):
        self.data_parallel_rank = data_parallel_rank
        self.data_parallel_world_size = data_parallel_world_size
        bucket = self.datasets[0].size_bucket
        assert all(ds.size_bucket == bucket for ds in self.datasets)
        self.iteration_order = np.array(_interleave_order([len(ds) for ds in self.datasets]))
        # images (frame count 1) and videos may use different batch sizes; either may be per-resolution
        table = global_batch_size_image if bucket[-1] == 1 else global_batch_size
        if None in table:
            self.global_batch_size = table[None]
        else:
            side = math.sqrt(bucket[-2] * bucket[-3])
            best = float('inf')
            for size, bs in table.items():
                if abs(size - side) < best:
                    best = abs(size - side)
                    self.global_batch_size = bs
        assert self.global_batch_size % data_parallel_world_size == 0
        keep = (len(self.iteration_order) // self.global_batch_size) * self.global_batch_size
        self.iteration_order = self.iteration_order[:keep]
        self.batch_size = self.global_batch_size // data_parallel_world_size
        self.ready = True

    def __len__(self):
        assert self.ready
        return len(self.iteration_order) // self.global_batch_size

    def __getitem__(self, idx):
        assert self.ready
        lo = idx * self.global_batch_size + self.data_parallel_rank * self.batch_size
        return [self.datasets[int(d)][int(j)] for d, j in self.iteration_order[lo:lo + self.batch_size]]


class BatchedDataset:
    """Order of global batches across size buckets + collation (the hot-path half of the reference's `Dataset`)."""

    def __init__(self, size_bucket_datasets, dataset_config=None):
        self.size_bucket_datasets = list(size_bucket_datasets)
        self.dataset_config = dataset_config or {}
        self.ready = False
        self.eval_quantile = None

    def post_init(self, data_parallel_rank, data_parallel_world_size, per_device_batch_size, gradient_accumulation_steps,
                  per_device_batch_size_image):
        self.data_parallel_rank = data_parallel_rank
        self.data_parallel_world_size = data_parallel_world_size
        scale = gradient_accumulation_steps * data_parallel_world_size
        gbs = {k: v * scale for k, v in per_device_batch_size.items()}
        gbs_img = {k: v * scale for k, v in per_device_batch_size_image.items()}
        by_bucket = defaultdict(list)
        for ds in self.size_bucket_datasets:
            by_bucket[ds.size_bucket].append(ds)
        self.buckets = [BucketBatches(v) for v in by_bucket.values()]
        for b in self.buckets:
            b.post_init(gbs, gbs_img, data_parallel_rank, data_parallel_world_size)
        self.iteration_order = _interleave_order([len(b) for b in self.buckets])
        self.ready = True
        if ratio := self.dataset_config.get('subsample_ratio', None):
            self.iteration_order = self.iteration_order[:int(len(self) * ratio)]

    def set_eval_quantile(self, q):
        self.eval_quantile = q

    def __len__(self):
        assert self.ready
        return len(self.iteration_order)

    def __getitem__(self, idx):
        assert self.ready
        b, j = self.iteration_order[idx]
        return self.collate(self.buckets[b][j])

    @staticmethod
    def collate(examples):
        out = {}
        for key in examples[0]:
            if key == 'mask':
                continue
            vals = [ex[key] for ex in examples]
            if torch.is_tensor(vals[0]) and all(torch.is_tensor(v) and v.shape == vals[0].shape for v in vals):
                vals = torch.stack(vals)
            out[key] = vals
        masks = [ex['mask'] for ex in examples]
        shape = None
        for m in masks:
            if m is not None:
                assert shape is None or m.shape == shape
                shape = m.shape
        if shape is None:
            out['mask'] = None          # the loss skips masking entirely
        else:
            out['mask'] = torch.stack([m if m is not None else torch.ones(shape, dtype=torch.float16) for m in masks])
        return out


def split_batch(batch, pieces):
    """(features, label) of batch size B -> `pieces` micro-batches of B // pieces; None fields become empty tensors."""
    features, label = batch
    n = features[0].size(0) // pieces

    def cut(fields):
        cols = [torch.split(t, n) if t is not None else [torch.tensor([])] * pieces for t in fields]
        return list(zip(*cols))
    return list(zip(cut(features), cut(label)))


class SkipFirstNSampler(torch.utils.data.Sampler):
    def __init__(self, n, dataset_length):
        super().__init__()
        self.n, self.dataset_length = n, dataset_length

    def __len__(self):
        return self.dataset_length

    def __iter__(self):
        return iter(range(self.n, self.dataset_length))


class PipelineDataLoader:
    """Endless micro-batch iterator.  One batch ahead is always pulled so that `epoch` advances exactly when the last
    micro-batch of an epoch is handed out; resume re-creates the loader skipping the batches already consumed."""

    def __init__(self, dataset, model_engine, gradient_accumulation_steps, model, num_dataloader_workers=1):
        if len(dataset) == 0:
            raise RuntimeError(
                'Processed dataset was empty. Probably caused by rounding down for each size bucket.\n'
                'Try decreasing the global batch size, or increasing num_repeats.\n'
                f'The dataset config that triggered this error was:\n{getattr(dataset, "dataset_config", None)}')
        self.model = model
        self.dataset = dataset
        self.model_engine = model_engine
        self.gradient_accumulation_steps = gradient_accumulation_steps
        self.num_dataloader_workers = num_dataloader_workers
        self.iter_called = False
        self.eval_quantile = None
        self.epoch = 1
        self.num_batches_pulled = 0
        self.next_micro_batch = None
        self.recreate_dataloader = False
        self._epoch0, self._consumed0, self._steps = 1, 0, 0     # for the communication-free epoch agreement (sync_epoch)
        self._ever_pulled = False                                 # middle pipeline stages never pull
        self._create_dataloader()
        self.data = self._micro_batches()

    def reset(self):
        self.epoch = 1
        self.num_batches_pulled = 0
        self.next_micro_batch = None
        self._epoch0, self._consumed0, self._steps = 1, 0, 0
        self.data = self._micro_batches()

    def set_eval_quantile(self, q):
        self.eval_quantile = q

    def __iter__(self):
        self.iter_called = True
        return self

    def __len__(self):
        return len(self.dataset) * self.gradient_accumulation_steps

    def __next__(self):
        if self.next_micro_batch is None:
            self.next_micro_batch = next(self.data)
        current = self.next_micro_batch
        try:
            self.next_micro_batch = next(self.data)
        except StopIteration:
            if self.recreate_dataloader:
                self._create_dataloader()
                self.recreate_dataloader = False
            self.data = self._micro_batches()
            self.num_batches_pulled = 0
            self.next_micro_batch = None
            self.epoch += 1
        return current

    def _create_dataloader(self, skip_first_n_batches=None):
        sampler = SkipFirstNSampler(skip_first_n_batches, len(self.dataset)) if skip_first_n_batches is not None else None
        workers = self.num_dataloader_workers
        self.dataloader = torch.utils.data.DataLoader(
            self.dataset, pin_memory=False, batch_size=None, sampler=sampler, num_workers=workers,
            persistent_workers=(workers > 0), prefetch_factor=2 if workers > 0 else None)

    def _micro_batches(self):
        for batch in self.dataloader:
            features, label = self.model.prepare_inputs(batch, timestep_quantile=self.eval_quantile)
            *targets, mask = label
            # the target depends on the noise drawn on the first stage: ship it to the last stage
            label = (*[self._broadcast_target(t) for t in targets], mask)
            self.num_batches_pulled += 1
            self._ever_pulled = True
            yield from split_batch((features, label), self.gradient_accumulation_steps)

    def _broadcast_target(self, target):
        engine = self.model_engine
        if not engine.is_pipe_parallel:
            return target
        assert engine.is_first_stage() or engine.is_last_stage()
        grid = engine.grid
        src, dst = grid.stage_to_global(0), grid.stage_to_global(engine.num_stages - 1)
        assert src in grid.pp_group and dst in grid.pp_group
        target = target.to(engine.device)
        if engine.is_first_stage():
            dist.send(target, dst)
        else:
            dist.recv(target, src)
        return target

    def sync_epoch(self):
        """Call once per step on every rank (train.py:921).  Middle stages never touch the dataloader; the reference makes
        everyone adopt the largest epoch seen with an `all_gather_object` per step (utils/dataset.py:1410-1417) — a host
        round trip over every rank.  The epoch is a pure function of the number of steps taken: every step consumes exactly
        one item of the dataset (`gradient_accumulation_steps` micro-batches), and `epoch` advances during the step that
        hands out the last one.  So every rank counts steps and computes the same value locally; the ranks that do pull
        data check it against what they observed.  DPIPE_CHECK_EPOCH_SYNC=1 additionally runs the reference's all-gather
        and asserts agreement."""
        self._steps += 1
        expected = self._epoch0 + (self._consumed0 + self._steps) // len(self.dataset)
        pulls = self.model_engine.is_first_stage() or self.model_engine.is_last_stage()
        if pulls and self.iter_called and expected != self.epoch:
            raise RuntimeError(f'epoch bookkeeping diverged: counted {expected}, dataloader says {self.epoch} '
                               f'(sync_epoch must be called exactly once per step)')
        self.epoch = expected
        if os.environ.get('DPIPE_CHECK_EPOCH_SYNC') == '1' and dist.get_world_size() > 1:
            seen = [None] * dist.get_world_size()
            dist.all_gather_object(seen, self.epoch)
            assert len(set(seen)) == 1, f'ranks disagree on the epoch: {seen}'

    def set_epoch(self, epoch):
        """`--reset_dataloader` on resume (train.py:876-877): keep the checkpoint's epoch number but start the data order
        from the beginning; the step-counted epoch agreement restarts from here too."""
        self.epoch = int(epoch)
        self._epoch0, self._consumed0, self._steps = self.epoch, 0, 0

    def state_dict(self):
        """Identical on every rank of a pipeline.  Ranks that pull data report what they pulled (the reference's value,
        utils/dataset.py:1419-1423); middle stages never touch the dataloader, so they report the same number derived from
        the step count: k items consumed since the start of the epoch means k+1 pulled (one is always pre-pulled), and 0
        right after the roll-over that handed out the epoch's last micro-batch.  (A checkpoint taken exactly on an epoch
        boundary stores 0, which the reference's `- 1` resume rule turns into "replay the last batch first",
        utils/dataset.py:1430 — kept, and counted, here.)"""
        pulled = self.num_batches_pulled
        if not self._ever_pulled and self._steps > 0:
            n, k = len(self.dataset), self._steps
            first = n - self._consumed0               # steps the (possibly resumed) first epoch lasts
            if k < first:
                pulled = self._consumed0 + k + 1
            else:
                kk = (k - first) % n
                pulled = kk + 1 if kk > 0 else 0
        return {'epoch': self.epoch, 'num_batches_pulled': pulled}

    def load_state_dict(self, state):
        assert not self.iter_called
        self.epoch = state['epoch']
        # one batch is always pre-pulled, so one fewer has actually been consumed
        self.num_batches_pulled = state['num_batches_pulled'] - 1
        self._epoch0, self._consumed0, self._steps = self.epoch, self.num_batches_pulled, 0
        self._create_dataloader(skip_first_n_batches=self.num_batches_pulled)
        self.data = self._micro_batches()
        self.recreate_dataloader = True   # skip only on the first pass


def get_data_iterator_for_step(dataloader, engine, num_micro_batches=None):
    """All micro-batches of a step are pulled up front (pulling does stage-0 -> last-stage traffic that must not interleave
    with the schedule's own sends); middle stages get None."""
    n = num_micro_batches or engine.micro_batches
    if not (engine.is_first_stage() or engine.is_last_stage()):
        return None
    it = iter(dataloader)
    return iter([next(it) for _ in range(n)])


# ---------------------------------------------------------------------------------------------------------------------
# the reference's on-disk cache (utils/cache.py) — read side.  The north star leaves latent / text-embedding caching to
# the reference; this is how the hot path consumes what it wrote.
# ---------------------------------------------------------------------------------------------------------------------
class ReferenceCache:
    """Read-only view of one `utils.cache.Cache` directory: `metadata.db` (sqlite: `items(shard, shard_index)` in insertion
    order, one `shard_<n>(offset, size)` table per shard) + `shard_<n>.bin` (concatenated `torch.save` blobs)
    (utils/cache.py:10-36,40-76,107-128).  Nothing is written and the fingerprint is not checked: regenerating a stale
    cache is the caching stage's job."""

    def __init__(self, path):
        import sqlite3
        self.path = str(path)
        db = os.path.join(self.path, 'metadata.db')
        if not os.path.exists(db):
            raise FileNotFoundError(f'{db}: not a cache directory written by the reference (utils/cache.py)')
        con = sqlite3.connect(f'file:{db}?mode=ro', uri=True)
        try:
            self.items = con.execute('SELECT shard, shard_index FROM items').fetchall()
            self.shard_metadata = {}
            for (name,) in con.execute("SELECT name FROM sqlite_master WHERE type = 'table'").fetchall():
                if name.startswith('shard_'):
                    self.shard_metadata[int(name.split('_')[-1])] = con.execute(f'SELECT offset, size FROM {name}').fetchall()
            row = con.execute('SELECT value FROM fingerprint').fetchone()
            self.fingerprint = row[0] if row else None
        finally:
            con.close()
        self._files = {}

    def __len__(self):
        return len(self.items)

    def __getitem__(self, idx):
        import io
        shard, k = self.items[idx]
        offset, size = self.shard_metadata[shard][k]
        f = self._files.get(shard)
        if f is None:
            f = self._files[shard] = open(os.path.join(self.path, f'shard_{shard}.bin'), 'rb')
        f.seek(offset)
        return torch.load(io.BytesIO(f.read(size)), map_location='cpu', weights_only=False)

    def __getstate__(self):            # DataLoader workers reopen their own file handles
        d = dict(self.__dict__)
        d['_files'] = {}
        return d


class ReferenceCacheBucket:
    """One size-bucket cache directory of the reference (`<dataset>/cache/<model>/cache_<w>x<h>x<frames>/`,
    utils/dataset.py:217): `latents/` plus `text_embeddings_<i>/` caches, merged per example the way
    SizeBucketDataset.__getitem__ does (utils/dataset.py:309-333) for the plain layout — one caption per item, no
    unconditional dropout — where item i of every cache belongs to example i (the caches are filled in the same, already
    shuffled, order: utils/dataset.py:212,233-243,178-200).  Datasets with several captions per image or caption JSON
    files need the reference's metadata tables and are not read here."""

    def __init__(self, path, num_repeats=1):
        self.path = str(path)
        self.latents = ReferenceCache(os.path.join(self.path, 'latents'))
        names = sorted((d for d in os.listdir(self.path) if d.startswith('text_embeddings_') and os.path.isdir(os.path.join(self.path, d))),
                       key=lambda d: int(d.rsplit('_', 1)[1]))
        self.text = [ReferenceCache(os.path.join(self.path, d)) for d in names]
        for d, c in zip(names, self.text):
            if len(c) != len(self.latents):
                raise RuntimeError(f'{self.path}/{d} holds {len(c)} items for {len(self.latents)} latents: several captions per '
                                   'image (or a partial cache) need the reference\'s metadata tables')
        self.num_repeats = num_repeats
        m = re.fullmatch(r'cache_(\d+)x(\d+)x(\d+)', os.path.basename(os.path.normpath(self.path)))
        first = self.latents[0]['latents'] if len(self.latents) else None
        if m:
            w, h, frames = (int(x) for x in m.groups())
        else:
            h, w = (first.shape[-2] * 8, first.shape[-1] * 8) if first is not None else (0, 0)
            frames = first.shape[-3] if first is not None and first.dim() == 4 else 1
        self.size_bucket = (round(w / h, 3) if h else 1.0, w, h, frames)

    def __len__(self):
        return int(len(self.latents) * self.num_repeats)

    def __getitem__(self, idx):
        i = idx % len(self.latents)
        ex = dict(self.latents[i])
        for c in self.text:
            ex.update(c[i])
        ex.setdefault('mask', None)
        ex.setdefault('caption', '')
        return ex
