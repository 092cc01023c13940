"""Generates tests/golden/ref_cache/: a tiny size-bucket cache written by the REFERENCE's own `utils/cache.py` (imported as it
is — it only needs sqlite3 and torch).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_cache.py

Layout as utils/dataset.py creates it for one size bucket (:217,:233-243,:178-200):
    ref_cache/cache_64x64x1/latents/            {'latents': [16, 8, 8], 'mask': None}           per item
    ref_cache/cache_64x64x1/text_embeddings_1/  {'t5_embed': [6, 32]}                            per item
    ref_cache/cache_64x64x1/text_embeddings_2/  {'clip_embed': [16]}                             per item
Five items, two shards for the latents (tiny shard size) so that the shard bookkeeping is exercised.  Tensors come from
tests/golden/synth.py, so the test regenerates what it expects to read.
"""
import importlib.util
import os
import shutil
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from synth import synth_tensor  # noqa: E402

OUT = os.path.join(HERE, 'ref_cache', 'cache_64x64x1')
N = 5


def main():
    spec = importlib.util.spec_from_file_location('reference_cache', '/root/reference/utils/cache.py')
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    shutil.rmtree(os.path.join(HERE, 'ref_cache'), ignore_errors=True)
    lat = mod.Cache(os.path.join(OUT, 'latents'), 'fixture-latents', shard_size_gb=1.5e-5)     # ~3 items per shard
    for i in range(N):
        lat.add({'latents': synth_tensor((16, 8, 8), 1100 + i, 1.0), 'mask': None})
    lat.finalize_current_shard()
    te1 = mod.Cache(os.path.join(OUT, 'text_embeddings_1'), 'fixture-te1')
    te2 = mod.Cache(os.path.join(OUT, 'text_embeddings_2'), 'fixture-te2')
    for i in range(N):
        te1.add({'t5_embed': synth_tensor((6, 32), 1200 + i, 1.0).bfloat16()})
        te2.add({'clip_embed': synth_tensor((16,), 1300 + i, 1.0).bfloat16()})
    te1.finalize_current_shard()
    te2.finalize_current_shard()
    for c in (lat, te1, te2):
        c.con.commit()
        c.con.close()
    for d, _, fs in os.walk(os.path.join(HERE, 'ref_cache')):
        for f in fs:
            print(os.path.relpath(os.path.join(d, f), HERE), os.path.getsize(os.path.join(d, f)))


if __name__ == '__main__':
    main()
