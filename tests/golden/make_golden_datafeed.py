"""Generates tests/golden/datafeed_traces.json by running the REFERENCE's own batch bookkeeping code.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_datafeed.py

The following definitions are taken verbatim (source text, via ast) from /root/reference/utils/dataset.py and executed:
  shuffle_with_seed (:41-45), ConcatenatedBatchedDataset (:340-396), Dataset.post_init/__len__/__getitem__/_collate
  (:953-1034), split_batch (:1273-1281), PipelineDataLoader (:1302-1435), SkipFirstNSampler (:1438-1449)
and /root/reference/train.py: get_data_iterator_for_step (:167-173).
Everything they touch outside themselves (deepspeed.comm, logger, the model's prepare_inputs, the directory datasets) is a
tiny stand-in defined here; examples are dicts carrying integer ids so the traces are pure integers.
"""
import ast
import json
import math
import os
import random
import types
from collections import defaultdict

import numpy as np
import torch

REF_DS = '/root/reference/utils/dataset.py'
REF_TRAIN = '/root/reference/train.py'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'datafeed_traces.json')


def extract(path, names, class_methods=None):
    tree = ast.parse(open(path).read())
    body = []
    for n in tree.body:
        if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names:
            if class_methods and isinstance(n, ast.ClassDef) and n.name in class_methods:
                n.body = [m for m in n.body if isinstance(m, ast.FunctionDef) and m.name in class_methods[n.name]]
                n.bases = []
            body.append(n)
    return compile(ast.Module(body=body, type_ignores=[]), path, 'exec')


class FakeDist:
    @staticmethod
    def get_world_group():
        return None

    @staticmethod
    def get_world_size(g=None):
        return 1

    @staticmethod
    def send(*a, **k):
        raise AssertionError

    recv = send


class FakeLogger:
    def warning(self, *a):
        pass


def load_reference():
    ns = {'random': random, 'np': np, 'torch': torch, 'math': math, 'defaultdict': defaultdict, 'dist': FakeDist,
          'logger': FakeLogger(), 'is_main_process': lambda: False, 'DEBUG': False}
    exec(extract(REF_DS, {'shuffle_with_seed', 'ConcatenatedBatchedDataset', 'split_batch', 'PipelineDataLoader',
                          'SkipFirstNSampler', 'Dataset'},
                 {'Dataset': {'post_init', '__len__', '__getitem__', '_collate'}}), ns)
    exec(extract(REF_TRAIN, {'get_data_iterator_for_step'}), ns)
    return types.SimpleNamespace(**ns)


class FakeSizeBucketDataset:
    """stands in for SizeBucketDataset: examples are dicts of small tensors tagged with (dataset id, index)"""

    def __init__(self, ds_id, size_bucket, n, with_mask=False):
        self.ds_id, self.size_bucket, self.n, self.with_mask = ds_id, size_bucket, n, with_mask

    def __len__(self):
        return self.n

    def __getitem__(self, idx):
        idx = idx % self.n
        ex = {'latents': torch.full((2, 2), float(self.ds_id * 1000 + idx)), 'id': self.ds_id * 1000 + idx, 'mask': None}
        if self.with_mask and idx % 3 == 0:
            ex['mask'] = torch.full((2, 2), 0.5)
        return ex


class FakeDirectoryDataset:
    def __init__(self, sbs):
        self.sbs = sbs

    def get_size_bucket_datasets(self):
        return self.sbs


class FakeModel:
    def prepare_inputs(self, batch, timestep_quantile=None):
        ids = torch.tensor(batch['id'])
        return (batch['latents'], ids), (batch['latents'] * 2, batch['mask'])


class FakeEngine:
    is_pipe_parallel = False
    micro_batches = 1

    def is_first_stage(self):
        return True

    def is_last_stage(self):
        return True


def make_dataset(ref, layout, with_mask=False):
    ds = ref.Dataset.__new__(ref.Dataset)
    ds.dataset_config = {}
    ds.post_init_called = False
    dirs, k = [], 0
    for d in layout:
        sbs = []
        for size_bucket, n in d:
            sbs.append(FakeSizeBucketDataset(k, tuple(size_bucket), n, with_mask))
            k += 1
        dirs.append(FakeDirectoryDataset(sbs))
    ds.directory_datasets = dirs
    return ds


LAYOUTS = {
    'single_bucket': [[((1.0, 512, 512, 1), 37)]],
    'two_dirs_mixed': [[((1.0, 512, 512, 1), 23), ((0.75, 448, 576, 1), 11)], [((1.0, 512, 512, 1), 9), ((1.33, 576, 448, 1), 17)]],
    'video_and_image': [[((1.0, 512, 512, 1), 20), ((1.0, 512, 512, 33), 14)], [((1.0, 1024, 1024, 1), 10)]],
}


def main():
    ref = load_reference()
    out = {'order': {}, 'loader': {}, 'split': {}}
    for lname, layout in LAYOUTS.items():
        for dp_world in (1, 2, 4):
            for mbs, gas, img_mbs in ((1, 1, 1), (2, 2, 2), (1, 4, 2), ({512: 2, 1024: 1}, 2, {512: 2, 1024: 1})):
                for dp_rank in range(dp_world):
                    ds = make_dataset(ref, layout)
                    pd = mbs if isinstance(mbs, dict) else {None: mbs}
                    pdi = img_mbs if isinstance(img_mbs, dict) else {None: img_mbs}
                    try:
                        ds.post_init(dp_rank, dp_world, pd, gas, pdi)
                    except AssertionError:
                        continue
                    batches = [[int(x) for x in ds[i]['id']] for i in range(len(ds))]
                    key = f'{lname}|{dp_world}|{dp_rank}|{json.dumps(mbs, sort_keys=True)}|{gas}|{json.dumps(img_mbs, sort_keys=True)}'
                    out['order'][key] = {'iteration_order': [list(map(int, x)) for x in ds.iteration_order], 'batches': batches}
    # PipelineDataLoader: epoch / pull counters and micro-batch contents, with a resume in the middle
    for lname, gas in (('single_bucket', 2), ('two_dirs_mixed', 3)):
        ds = make_dataset(ref, LAYOUTS[lname], with_mask=True)
        ds.post_init(0, 1, {None: 2}, gas, {None: 2})
        dl = ref.PipelineDataLoader(ds, FakeEngine(), gas, FakeModel(), num_dataloader_workers=0)
        trace = []
        n = int(len(dl) * 2.5)
        saved, saved_at = None, n // 3
        for i in range(n):
            mb = next(dl)
            (lat, ids), (tgt, mask) = mb
            trace.append([dl.epoch, dl.num_batches_pulled, [int(x) for x in ids], int(mask.numel())])
            if i == saved_at:
                saved = dict(dl.state_dict())
        ds2 = make_dataset(ref, LAYOUTS[lname], with_mask=True)
        ds2.post_init(0, 1, {None: 2}, gas, {None: 2})
        dl2 = ref.PipelineDataLoader(ds2, FakeEngine(), gas, FakeModel(), num_dataloader_workers=0)
        dl2.load_state_dict(saved)
        resumed = []
        for i in range(len(dl) + 2):
            mb = next(dl2)
            resumed.append([dl2.epoch, dl2.num_batches_pulled, [int(x) for x in mb[0][1]]])
        out['loader'][f'{lname}|{gas}'] = {'len': len(dl), 'trace': trace, 'saved_state': saved, 'saved_at': saved_at,
                                          'resumed': resumed}
    # split_batch incl. None -> empty tensor
    feats = (torch.arange(24).view(6, 4), None, torch.arange(6))
    label = (torch.arange(12).view(6, 2), None)
    pieces = ref.split_batch((feats, label), 3)
    out['split']['6x3'] = [[[t.tolist() for t in f], [t.tolist() for t in l]] for f, l in pieces]
    # get_data_iterator_for_step on first / middle stage
    class E(FakeEngine):
        micro_batches = 3
    class Mid(E):
        def is_first_stage(self):
            return False
        def is_last_stage(self):
            return False
    out['iter_for_step'] = {'first': [int(x) for x in ref.get_data_iterator_for_step(iter(range(100)), E())],
                            'middle': ref.get_data_iterator_for_step(iter(range(100)), Mid())}
    with open(OUT, 'w') as f:
        json.dump(out, f, separators=(',', ':'))
    print('wrote', OUT, os.path.getsize(OUT) // 1024, 'KiB', len(out['order']), 'orders')


if __name__ == '__main__':
    main()
