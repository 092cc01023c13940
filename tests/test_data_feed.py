"""CPU: batch-index bookkeeping and the micro-batch loader are BIT-EXACT with the reference's own code
(tests/golden/datafeed_traces.json, produced by tests/golden/make_golden_datafeed.py from utils/dataset.py and train.py)."""
import json
import os

import pytest
import torch

from diffusion_pipe_b200 import data_feed as DF


class SizeBucket:
    def __init__(self, ds_id, size_bucket, n, with_mask=False):
        self.ds_id, self.size_bucket, self.n, self.with_mask = ds_id, tuple(size_bucket), n, with_mask

    def __len__(self):
        return self.n

    def __getitem__(self, idx):
        idx = idx % self.n
        ex = {'latents': torch.full((2, 2), float(self.ds_id * 1000 + idx)), 'id': self.ds_id * 1000 + idx, 'mask': None}
        if self.with_mask and idx % 3 == 0:
            ex['mask'] = torch.full((2, 2), 0.5)
        return ex


LAYOUTS = {
    'single_bucket': [[((1.0, 512, 512, 1), 37)]],
    'two_dirs_mixed': [[((1.0, 512, 512, 1), 23), ((0.75, 448, 576, 1), 11)], [((1.0, 512, 512, 1), 9), ((1.33, 576, 448, 1), 17)]],
    'video_and_image': [[((1.0, 512, 512, 1), 20), ((1.0, 512, 512, 33), 14)], [((1.0, 1024, 1024, 1), 10)]],
}


def make(layout, with_mask=False):
    out, k = [], 0
    for d in layout:
        for sb, n in d:
            out.append(SizeBucket(k, sb, n, with_mask))
            k += 1
    return DF.BatchedDataset(out)


class Model:
    def prepare_inputs(self, batch, timestep_quantile=None):
        return (batch['latents'], torch.tensor(batch['id'])), (batch['latents'] * 2, batch['mask'])


class Engine:
    is_pipe_parallel = False
    micro_batches = 3

    def is_first_stage(self):
        return True

    def is_last_stage(self):
        return True


@pytest.fixture(scope='module')
def golden(golden_dir):
    with open(os.path.join(golden_dir, 'datafeed_traces.json')) as f:
        return json.load(f)


def test_iteration_order_and_rank_slices_are_bit_exact(golden):
    assert len(golden['order']) >= 80
    for key, want in golden['order'].items():
        lname, dp_world, dp_rank, mbs, gas, img = key.split('|')
        mbs, img = json.loads(mbs), json.loads(img)
        pd = {int(k): v for k, v in mbs.items()} if isinstance(mbs, dict) else {None: mbs}
        pdi = {int(k): v for k, v in img.items()} if isinstance(img, dict) else {None: img}
        ds = make(LAYOUTS[lname])
        ds.post_init(int(dp_rank), int(dp_world), pd, int(gas), pdi)
        assert [list(map(int, x)) for x in ds.iteration_order] == want['iteration_order'], key
        assert [[int(x) for x in ds[i]['id']] for i in range(len(ds))] == want['batches'], key


def test_pipeline_dataloader_epochs_and_resume(golden):
    for key, want in golden['loader'].items():
        lname, gas = key.split('|')
        gas = int(gas)
        ds = make(LAYOUTS[lname], with_mask=True)
        ds.post_init(0, 1, {None: 2}, gas, {None: 2})
        dl = DF.PipelineDataLoader(ds, Engine(), gas, Model(), num_dataloader_workers=0)
        assert len(dl) == want['len']
        saved = None
        for i, (epoch, pulled, ids, mask_numel) in enumerate(want['trace']):
            (lat, got_ids), (tgt, mask) = next(dl)
            assert (dl.epoch, dl.num_batches_pulled, [int(x) for x in got_ids], int(mask.numel())) == \
                (epoch, pulled, ids, mask_numel), (key, i)
            assert torch.equal(tgt, lat * 2)
            if i == want['saved_at']:
                saved = dl.state_dict()
        assert saved == want['saved_state']
        ds2 = make(LAYOUTS[lname], with_mask=True)
        ds2.post_init(0, 1, {None: 2}, gas, {None: 2})
        dl2 = DF.PipelineDataLoader(ds2, Engine(), gas, Model(), num_dataloader_workers=0)
        dl2.load_state_dict(saved)
        for i, (epoch, pulled, ids) in enumerate(want['resumed']):
            mb = next(dl2)
            assert (dl2.epoch, dl2.num_batches_pulled, [int(x) for x in mb[0][1]]) == (epoch, pulled, ids), (key, i)


def test_split_batch_and_none_fields(golden):
    feats = (torch.arange(24).view(6, 4), None, torch.arange(6))
    label = (torch.arange(12).view(6, 2), None)
    got = [[[t.tolist() for t in f], [t.tolist() for t in l]] for f, l in DF.split_batch((feats, label), 3)]
    assert got == golden['split']['6x3']
    assert DF.split_batch((feats, label), 3)[0][0][1].numel() == 0


def test_get_data_iterator_for_step(golden):
    class Mid(Engine):
        def is_first_stage(self):
            return False

        def is_last_stage(self):
            return False
    assert list(DF.get_data_iterator_for_step(iter(range(100)), Engine())) == golden['iter_for_step']['first']
    assert DF.get_data_iterator_for_step(iter(range(100)), Mid()) is golden['iter_for_step']['middle']


def test_seeded_shuffle_leaves_global_rng_untouched():
    import random
    random.seed(123)
    a = random.random()
    random.seed(123)
    x = list(range(50))
    DF.seeded_shuffle(x, 0)
    assert random.random() == a
    y = list(range(50))
    DF.seeded_shuffle(y, 0)
    assert x == y and x != list(range(50))


def test_empty_and_dropped_buckets():
    ds = DF.BatchedDataset([SizeBucket(0, (1.0, 512, 512, 1), 3)])
    ds.post_init(0, 1, {None: 2}, 2, {None: 2})          # global batch 4 > 3 examples: bucket dropped entirely
    assert len(ds) == 0
    with pytest.raises(RuntimeError):
        DF.PipelineDataLoader(ds, Engine(), 2, Model(), num_dataloader_workers=0)


def test_sync_epoch_is_a_pure_function_of_the_step_count(golden):
    """every rank — including middle stages that never see data — computes the epoch the pulling ranks observe, across
    epoch boundaries, after reset() and after a resume, without communication (utils/dataset.py:1410-1417 replaced)"""
    class Mid(Engine):
        def is_first_stage(self):
            return False

        def is_last_stage(self):
            return False
    for key in golden['loader']:
        lname, gas = key.split('|')
        gas = int(gas)

        def mk(engine):
            ds = make(LAYOUTS[lname], with_mask=True)
            ds.post_init(0, 1, {None: 2}, gas, {None: 2})
            return DF.PipelineDataLoader(ds, engine, gas, Model(), num_dataloader_workers=0)
        puller, middle = mk(Engine()), mk(Mid())
        n = len(puller.dataset)
        saved, saved_at = None, n + 1
        for step in range(1, 3 * n + 2):
            it = DF.get_data_iterator_for_step(puller, puller.model_engine, num_micro_batches=gas)
            assert len(list(it)) == gas
            assert DF.get_data_iterator_for_step(middle, middle.model_engine, num_micro_batches=gas) is None
            puller.sync_epoch()                      # raises if the count and the dataloader disagree
            middle.sync_epoch()
            assert middle.epoch == puller.epoch == 1 + step // n, (key, step)
            if step == saved_at:
                saved = (puller.state_dict(), middle.state_dict(), puller.epoch)
        # resume both from the saved state and keep going
        p2, m2 = mk(Engine()), mk(Mid())
        p2.load_state_dict(saved[0])
        m2.load_state_dict(saved[0])                 # every rank loads the same client_state (train.py:870-875)
        for step in range(saved_at + 1, saved_at + 2 * n):
            list(DF.get_data_iterator_for_step(p2, p2.model_engine, num_micro_batches=gas))
            p2.sync_epoch()
            m2.sync_epoch()
            assert m2.epoch == p2.epoch, (key, step)
        # reset() (evaluation loops) restarts the count
        puller.reset()
        middle.reset()
        for step in range(1, n + 1):
            list(DF.get_data_iterator_for_step(puller, puller.model_engine, num_micro_batches=gas))
            puller.sync_epoch()
            middle.sync_epoch()
            assert middle.epoch == puller.epoch
        assert puller.epoch == 2


def test_every_stage_saves_the_same_loader_state_and_resumes_in_step(golden):
    """every rank restores the loader from ITS OWN stage checkpoint (train.py:870-879).  Middle stages never
    pull data, so their saved state is derived from the step count and must equal what the pulling stages saved — mid-epoch,
    right at an epoch boundary and before the first step — or a >= 3-stage pipeline disagrees on the epoch after resume."""
    class Mid(Engine):
        def is_first_stage(self):
            return False

        def is_last_stage(self):
            return False
    key = sorted(golden['loader'])[0]
    lname, gas = key.split('|')
    gas = int(gas)

    def mk(engine):
        ds = make(LAYOUTS[lname], with_mask=True)
        ds.post_init(0, 1, {None: 2}, gas, {None: 2})
        return DF.PipelineDataLoader(ds, engine, gas, Model(), num_dataloader_workers=0)

    def run(loaders, steps):
        for _ in range(steps):
            for ld in loaders:
                it = DF.get_data_iterator_for_step(ld, ld.model_engine, num_micro_batches=gas)
                if it is not None:
                    assert len(list(it)) == gas
            for ld in loaders:
                ld.sync_epoch()
            assert len({ld.epoch for ld in loaders}) == 1
    n = len(mk(Engine()).dataset)
    assert n >= 3
    for save_after in (0, 1, n - 1, n, n + 2, 2 * n):
        first, middle, last = mk(Engine()), mk(Mid()), mk(Engine())
        run((first, middle, last), save_after)
        states = [ld.state_dict() for ld in (first, middle, last)]
        assert states[0] == states[1] == states[2], (save_after, states)
        # a second generation: resume, run across the next epoch boundary, save again, resume again
        for _generation in range(2):
            resumed = (mk(Engine()), mk(Mid()), mk(Engine()))
            for ld, st in zip(resumed, states):
                ld.load_state_dict(st)                          # each rank: its own checkpoint's client_state
            run(resumed, n + 1)
            states = [ld.state_dict() for ld in resumed]
            assert states[0] == states[1] == states[2], (save_after, states)


def test_reset_dataloader_keeps_the_checkpoint_epoch(golden):
    """`--reset_dataloader` (train.py:876-877) resumes at the checkpoint's epoch with the data order restarted; the
    step-counted epoch agreement must restart from that epoch instead of raising 'epoch bookkeeping diverged'."""
    key = sorted(golden['loader'])[0]
    lname, gas = key.split('|')
    gas = int(gas)
    ds = make(LAYOUTS[lname], with_mask=True)
    ds.post_init(0, 1, {None: 2}, gas, {None: 2})
    ld = DF.PipelineDataLoader(ds, Engine(), gas, Model(), num_dataloader_workers=0)
    ld.set_epoch(3)
    n = len(ds)
    for step in range(1, n + 2):
        list(DF.get_data_iterator_for_step(ld, ld.model_engine, num_micro_batches=gas))
        ld.sync_epoch()
        assert ld.epoch == 3 + step // n


def test_reference_cache_directories_are_read_as_the_reference_wrote_them(golden_dir):
    """tests/golden/ref_cache/ was written by the reference's own utils/cache.py (tests/golden/make_golden_cache.py): the
    read side (data_feed.ReferenceCache / ReferenceCacheBucket) returns the same items in the same order, across shards,
    and the cached examples feed the train CLI's dataset path"""
    import pickle
    import sys
    sys.path.insert(0, golden_dir)
    sys.path.insert(0, os.path.dirname(golden_dir).rsplit('/tests', 1)[0])
    from synth import synth_tensor
    from diffusion_pipe_b200 import data_feed
    root = os.path.join(golden_dir, 'ref_cache')
    lat = data_feed.ReferenceCache(os.path.join(root, 'cache_64x64x1', 'latents'))
    assert len(lat) == 5 and lat.fingerprint == 'fixture-latents' and sorted(lat.shard_metadata) == [0, 1]      # two shards
    assert lat.items == [(0, 0), (0, 1), (0, 2), (1, 0), (1, 1)]
    bucket = data_feed.ReferenceCacheBucket(os.path.join(root, 'cache_64x64x1'), num_repeats=2)
    assert len(bucket) == 10 and bucket.size_bucket == (1.0, 64, 64, 1)
    for i in (0, 3, 4, 7):
        ex = bucket[i]
        j = i % 5
        assert torch.equal(ex['latents'], synth_tensor((16, 8, 8), 1100 + j, 1.0)) and ex['mask'] is None and ex['caption'] == ''
        assert torch.equal(ex['t5_embed'], synth_tensor((6, 32), 1200 + j, 1.0).bfloat16())
        assert torch.equal(ex['clip_embed'], synth_tensor((16,), 1300 + j, 1.0).bfloat16())
    clone = pickle.loads(pickle.dumps(bucket))                      # DataLoader workers get a copy without open files
    assert torch.equal(clone[4]['latents'], bucket[4]['latents'])
    with pytest.raises(FileNotFoundError):
        data_feed.ReferenceCache(root)
    # through the driver's dataset config: a bucket directory, or the directory that holds the buckets
    import train as T
    for path in (os.path.join(root, 'cache_64x64x1'), root):
        buckets = T.load_size_buckets({'directory': [{'cache_dir': path, 'num_repeats': 1}]})
        assert len(buckets) == 1 and len(buckets[0]) == 5
        ds = data_feed.BatchedDataset(buckets, {})
        ds.post_init(0, 1, {None: 1}, 2, {None: 1})
        batch = ds[0]
        assert batch['latents'].shape == (2, 16, 8, 8) and batch['t5_embed'].shape == (2, 6, 32) and batch['mask'] is None
