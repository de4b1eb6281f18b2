"""N>1 path on CPU: world_size-2 gloo run of the spp-sharding + accumulator all-reduce + finalize logic.
Each rank's slice is rendered by the oracle (sample ranges of the same per-(pixel, sample) streams); the
reduced image must equal the oracle's full render."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_util
from miniraytracer_b200 import accfile, distributed as mdist


def test_shard_ranges_tile_exactly():
    for n in (1, 16, 121, 1024, 4096):
        for world in (1, 2, 3, 4, 8):
            r = [mdist.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        mdist.shard_range(16, 2, 2)


def _worker(rank, world, port, scene, w, h, spp, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = spp
    b, e = mdist.shard_range(n, rank, world)
    acc_np, _ = oracle_util.ref_render(scene, w, h, spp, s0=b, s1=e)
    acc = torch.from_numpy(np.ascontiguousarray(acc_np))
    mdist.allreduce_accumulator(acc)
    img = mdist.finalize_torch(acc)
    if rank == 0:
        np.save(out_path, np.concatenate([img.numpy(), acc[..., 3:4].numpy()], axis=-1))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(not oracle_util.have_ref(), reason="oracle/_ref/mrt_ref not built")
def test_two_rank_reduce_matches_full_render(tmp_path):
    scene, w, h, spp = 5, 64, 36, 16
    full, _ = oracle_util.ref_render(scene, w, h, spp)
    out = str(tmp_path / "reduced.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, scene, w, h, spp, out), nprocs=2, join=True)
    got = np.load(out)
    np.testing.assert_array_equal(got[..., 3], full[..., 3])
    res = accfile.compare(got[..., :3], accfile.finalize(full), rel=1e-5)
    assert res["n_bad"] == 0, res


def test_progressive_schedule_covers_slice():
    for n, world, passes in ((16, 2, 4), (1024, 8, 8), (121, 3, 5), (4, 2, 8)):
        seen = []
        for rank in range(world):
            sched = mdist.progressive_schedule(n, rank, world, passes)
            assert len(sched) == passes
            b, e = mdist.shard_range(n, rank, world)
            assert sched[0][0] == b and sched[-1][1] == e
            assert all(sched[i][1] == sched[i + 1][0] for i in range(passes - 1))
            seen += [s for x, y in sched for s in range(x, y)]
        assert sorted(seen) == list(range(n))


def _progressive_worker(rank, world, port, scene, w, h, spp, passes, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    acc = torch.zeros((h, w, 4), dtype=torch.float32)
    previews = []

    def render_pass(b, e, out):   # stand-in for Renderer.render_async(accumulate=True): the oracle renders the slice
        part, _ = oracle_util.ref_render(scene, w, h, spp, s0=b, s1=e)
        out += torch.from_numpy(np.ascontiguousarray(part))

    red = mdist.ProgressiveReducer(acc, render_pass, preview=lambda p, buf: previews.append((p, buf.clone())))
    final = red.run(mdist.progressive_schedule(spp, rank, world, passes))
    if rank == 0:
        np.savez(out_path, final=final.numpy(), **{f"p{p}": b.numpy() for p, b in previews})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(not oracle_util.have_ref(), reason="oracle/_ref/mrt_ref not built")
def test_progressive_two_rank_previews_and_final(tmp_path):
    """Sample-major passes on two ranks with the per-pass reduce overlapped with the next pass: every preview is the
    sum of exactly the samples rendered so far by both ranks, the last one is the full render."""
    scene, w, h, spp, passes = 5, 48, 27, 16, 4
    out = str(tmp_path / "prog.npz")
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_progressive_worker, args=(2, port, scene, w, h, spp, passes, out), nprocs=2, join=True)
    got = np.load(out)
    full, _ = oracle_util.ref_render(scene, w, h, spp)
    res = accfile.compare(accfile.finalize(got["final"]), accfile.finalize(full), rel=1e-5)
    assert res["n_bad"] == 0, res
    np.testing.assert_array_equal(got["final"][..., 3], full[..., 3])
    for p in range(passes):
        want = np.zeros_like(full)
        for rank in range(2):
            for b, e in mdist.progressive_schedule(spp, rank, 2, passes)[:p + 1]:
                if e > b:
                    want += oracle_util.ref_render(scene, w, h, spp, s0=b, s1=e)[0]
        res = accfile.compare(accfile.finalize(got[f"p{p}"]), accfile.finalize(want), rel=1e-5)
        assert res["n_bad"] == 0, (p, res)
        np.testing.assert_array_equal(got[f"p{p}"][..., 3], want[..., 3])
