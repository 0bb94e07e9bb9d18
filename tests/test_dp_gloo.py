"""Data-parallel host logic on CPU with the gloo backend, world_size 2: bucketed SUM all-reduce over the flat gradient
arena + 1/W in the clip coefficient == mean of per-rank gradients (DDP semantics, scripts/pretrain_virtex.py:121-123)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import virtex_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _layout(spec):
    names, offsets, numels, off = [], {}, {}, 0
    for n, shape in O.unique_shapes(spec).items():
        if O.is_buffer(n):
            continue
        k = 1
        for d in shape:
            k *= d
        names.append(n)
        offsets[n], numels[n] = off, k
        off = (off + k + 63) // 64 * 64
    return names, offsets, numels, off


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    from virtex_b200.distributed import average_across_processes, get_rank, get_world_size, init_from_env
    from virtex_b200.trainer import BUCKET_ORDER, bucket_ranges
    init_from_env(backend="gloo")
    assert get_world_size() == world and get_rank() == rank
    spec = O.Spec(hidden=128, layers=1, heads=2, ffn=256)
    state = O.synth_state(spec, 3)
    names, offsets, numels, total = _layout(spec)
    batch = O.synth_batch(1, seed=50 + rank)  # per-rank shard, per-rank seed
    _, grads, _ = O.loss_and_grads(state, batch, spec)
    flat = torch.zeros(total)
    for n in names:
        flat[offsets[n]:offsets[n] + numels[n]] = grads[n].flatten()
    local = flat.clone()
    ranges = bucket_ranges(names, offsets, numels)
    covered = torch.zeros(total, dtype=torch.bool)
    for tag in BUCKET_ORDER:  # buckets fire in backward-completion order, each a contiguous slice
        b, e = ranges[tag]
        dist.all_reduce(flat[b:e], op=dist.ReduceOp.SUM)
        covered[b:e] = True
    for n in names:
        assert covered[offsets[n]:offsets[n] + numels[n]].all(), n
    mean = flat / world
    loss = average_across_processes(torch.tensor([float(rank + 1)]))
    assert abs(loss.item() - (1 + world) / 2) < 1e-6
    torch.save({"local": local, "mean": mean}, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_allreduce_equals_mean_of_rank_grads(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(tmp_path / f"rank{i}.pt") for i in range(world)]
    expect = (r[0]["local"] + r[1]["local"]) / world
    for i in range(world):
        assert torch.allclose(r[i]["mean"], expect, rtol=1e-6, atol=1e-8)
    assert not torch.allclose(r[0]["local"], r[1]["local"])  # the shards really differ


def test_bucket_ranges_partition_the_arena():
    from virtex_b200.trainer import BUCKET_ORDER, bucket_ranges
    spec = O.Spec()
    names, offsets, numels, total = _layout(spec)
    ranges = bucket_ranges(names, offsets, numels)
    spans = sorted(ranges[t] for t in BUCKET_ORDER)
    assert spans[0][0] == 0
    for (b0, e0), (b1, e1) in zip(spans, spans[1:]):
        assert e0 <= b1  # disjoint, ordered
    assert sum(e - b for b, e in spans) >= sum(numels.values())


def test_optimizer_segments_follow_name_rules():
    from virtex_b200.config import Config
    from virtex_b200.factories import param_group_hparams
    from virtex_b200.trainer import optimizer_segments
    cfg = Config()
    spec = O.Spec()
    names, offsets, numels, total = _layout(spec)
    segs = optimizer_segments(names, offsets, numels, lambda n: param_group_hparams(cfg, n))
    assert sum(e - b for b, e, _, _ in segs) == 69_482_320
    assert all(e - b <= 65536 for b, e, _, _ in segs)
    by_name = {n: param_group_hparams(cfg, n) for n in names}
    assert by_name["visual.cnn.conv1.weight"] == (0.2, 1e-4)
    assert by_name["textual.transformer.layers.0.norm1.weight"] == (0.001, 0.0)
    assert by_name["textual.output.bias"] == (0.001, 1e-4)     # not matched by the NO_DECAY regex
    assert by_name["textual.embedding.layer_norm.bias"] == (0.001, 0.0)
    assert sum(1 for v in by_name.values() if v[0] == 0.2) == 159
    assert sum(1 for v in by_name.values() if v[1] == 0.0) == 26
