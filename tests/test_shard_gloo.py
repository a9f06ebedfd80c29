"""Multi-rank host logic on CPU (gloo, world_size 2 and 3): shard ranges, the single all-gather of token ids in
global image order, and world-size independence of the per-index synthetic inputs."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deephumor_b200.runtime import shard
from deephumor_b200.utils import synth


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


class _FakeCaptioner:
    """Stands in for a captioner on CPU: ids are a pure function of (global image index, label) like the real
    path under the injected noise model, so sharded and unsharded jobs must agree exactly."""

    def generate(self, images, labels, max_len=32, image_base=0, **kw):
        n = images.shape[0]
        g = torch.arange(image_base, image_base + n, dtype=torch.int64)
        lens = (g * 7 + labels[:, 0]) % max_len + 1
        ids = (g.unsqueeze(1) * 131 + torch.arange(max_len) + images[:, 0, 0, :1].mul(1000).long()) % 36541
        ids = torch.where(torch.arange(max_len).unsqueeze(0) < lens.unsqueeze(1), ids, torch.zeros_like(ids))
        return ids, lens


def _inputs(first, count):
    return synth.images(0, first, count, size=8), synth.labels(0, first, count, 36541)


def _worker(rank, world, port, total, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        ids, lens = shard.generate_sharded(_FakeCaptioner(), total, _inputs, max_len=32)
        torch.save((ids, lens), os.path.join(out_dir, f'r{rank}.pt'))
    finally:
        dist.destroy_process_group()


def test_shard_range_covers_everything_once():
    for total in (0, 1, 7, 512, 65536, 65537):
        for world in (1, 2, 3, 8):
            spans = [shard.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f1 == f0 + c0
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        shard.shard_range(8, 2, 2)


def test_synthetic_inputs_are_keyed_by_global_index():
    full = synth.images(0, 0, 6, size=8)
    parts = torch.cat([synth.images(0, 0, 4, size=8), synth.images(0, 4, 2, size=8)])
    assert torch.equal(full, parts)
    assert torch.equal(synth.labels(0, 0, 6, 36541)[4:], synth.labels(0, 4, 2, 36541))


@pytest.mark.parametrize('world,total', [(2, 10), (2, 7), (3, 8)])
def test_all_gather_of_ids_matches_single_rank(world, total, tmp_path):
    ref_ids, ref_lens = _FakeCaptioner().generate(*_inputs(0, total), max_len=32, image_base=0)
    mp.spawn(_worker, args=(world, _free_port(), total, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        ids, lens = torch.load(os.path.join(tmp_path, f'r{r}.pt'))
        assert ids.shape == (total, 32) and torch.equal(ids, ref_ids) and torch.equal(lens, ref_lens)


def test_gather_is_identity_without_process_group():
    ids, lens = torch.arange(64).view(2, 32), torch.tensor([3, 4])
    out = shard.gather_captions(ids, lens)
    assert out[0] is ids and out[1] is lens
