"""The N > 1 path on CPU: world_size-2 `gloo` ranks, each owning a contiguous shard of the streams in
its own engine (the emu build here; the CUDA library on GPUs), no data-path collective, and the
optional all-streams bus as a per-rank partial + one all-reduce.

Sharding must not change any stream's output: rank r's stream s equals stream (offset_r + s) of a
single-engine run, bit for bit; the reduced bus equals the float64 sum over all streams.
"""
import os
import socket
import sys

import numpy as np
import pytest

import harness as H


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total_streams, frames, out_dir):
    sys.path.insert(0, os.path.join(H.ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import oalsfxpp_b200 as ox
    from oalsfxpp_b200 import ChannelFormat as F, EffectType as T

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = H.emu_lib()
    per = total_streams // world
    first = rank * per
    x = np.stack([H.noise(first + s, 2, frames) for s in range(per)])
    with ox.Engine(per, F.stereo, 48000, 2, lib=lib) as eng:
        eng.set_effect(0, T.echo)
        eng.set_effect(1, T.eax_reverb)
        y = eng.mix(x)
        bus = np.zeros((frames, 2), np.float32)
        eng.reduce_bus(frames, y, bus)
    total = torch.from_numpy(bus.copy())
    dist.all_reduce(total, op=dist.ReduceOp.SUM)
    np.save(os.path.join(out_dir, f"y{rank}.npy"), y)
    if rank == 0:
        np.save(os.path.join(out_dir, "bus.npy"), total.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_bus_reduce(tmp_path):
    import torch.multiprocessing as mp
    import oalsfxpp_b200 as ox
    from oalsfxpp_b200 import ChannelFormat as F, EffectType as T

    H.emu_lib()  # build once before the ranks start
    world, total_streams, frames = 2, 48, 700
    port = _free_port()
    mp.spawn(_worker, args=(world, port, total_streams, frames, str(tmp_path)), nprocs=world, join=True)

    x = np.stack([H.noise(s, 2, frames) for s in range(total_streams)])
    with ox.Engine(total_streams, F.stereo, 48000, 2, lib=H.emu_lib()) as eng:
        eng.set_effect(0, T.echo)
        eng.set_effect(1, T.eax_reverb)
        whole = eng.mix(x)
    per = total_streams // world
    for r in range(world):
        shard = np.load(tmp_path / f"y{r}.npy")
        assert np.array_equal(shard, whole[r * per:(r + 1) * per]), f"rank {r} differs from the unsharded run"
    bus = np.load(tmp_path / "bus.npy")
    want = whole.astype(np.float64).sum(axis=0)
    assert np.max(np.abs(bus - want)) <= 1e-5 * np.sqrt(total_streams)
