"""world_size-2 gloo test (CPU) of the N>1 path: packed-weight broadcast + contiguous batch sharding + sliced noise."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from surfd_b200 import synth, unet as U
from surfd_b200.decoder import pack_decoder
from surfd_b200.dist import broadcast_packed, conditioning_from_rank0, shard_range, sliced_noise


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L = 32
    a = U.arch(L)
    if rank == 0:
        blob_u, prog, _ = U.pack_unet(synth.synth_mdm(L), L)
        blob_d = pack_decoder(synth.synth_ae_poly(L)["decoder"], L)
    else:
        blob_u = torch.zeros(a.n_floats)
        prog = torch.zeros(16 + len(a.buffers) + len(a.prog) * U.REC, dtype=torch.int64)
        blob_d = torch.zeros(pack_decoder(synth.synth_ae_rand(L, 1)["decoder"], L).numel())
    broadcast_packed([blob_u, prog, blob_d], 0)
    lo, hi = shard_range(13, world, rank)
    noise = sliced_noise(10, 5, 13, L, lo, hi)
    calls = []

    def make():                     # stands for the CLIP pass: must run on rank 0 only
        calls.append(rank)
        return torch.arange(13 * 512, dtype=torch.float32).reshape(13, 512) * 0.5

    ctx = conditioning_from_rank0(make, 13, 512, "cpu")[lo:hi]
    torch.save({"ctx": ctx, "calls": calls, "sum_u": float(blob_u.double().sum()), "sum_d": float(blob_d.double().sum()), "prog": int(prog.sum()),
                "lo": lo, "hi": hi, "noise": noise}, out + f".{rank}")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_broadcast_and_sharding(tmp_path):
    world, port = 2, _free_port()
    out = str(tmp_path / "r")
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    r0, r1 = torch.load(out + ".0"), torch.load(out + ".1")
    assert r0["sum_u"] == r1["sum_u"] and r0["sum_d"] == r1["sum_d"] and r0["prog"] == r1["prog"]   # weights arrived bit-identical
    assert (r0["lo"], r0["hi"], r1["lo"], r1["hi"]) == (0, 7, 7, 13)                                  # contiguous shards cover the batch
    full = sliced_noise(10, 5, 13, 32, 0, 13)
    assert torch.equal(torch.cat([r0["noise"], r1["noise"]], 1), full)                                # same noise as a single-GPU run
    assert r0["calls"] == [0] and r1["calls"] == []                                                    # the encoder ran on rank 0 only
    assert torch.equal(torch.cat([r0["ctx"], r1["ctx"]], 0), torch.arange(13 * 512, dtype=torch.float32).reshape(13, 512) * 0.5)


def test_single_process_conditioning_is_a_plain_call():
    t = torch.ones(3, 512)
    assert conditioning_from_rank0(lambda: t, 3, 512, "cpu") is t


def test_shard_range_edges():
    assert shard_range(8, 8, 3) == (3, 4) and shard_range(3, 8, 5) == (3, 3) and shard_range(64, 8, 7) == (56, 64)
    assert [shard_range(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]
