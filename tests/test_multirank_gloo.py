"""CPU, world_size 2, gloo: the N>1 host path -- request sharding, checkpoint broadcast from rank 0
(tied tensors stay tied), max-over-ranks timing reduction."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vallex_b200.dist import broadcast_state_dict, max_over_ranks, shard_units


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(5)
    sd = None
    if rank == 0:
        tied = torch.randn(8, 4, generator=g)
        sd = {"a.weight": torch.randn(33, 7, generator=g), "nar_audio_embeddings.2.word_embeddings.weight": tied,
              "b.alpha": torch.ones(1), "nar_predict_layers.0.weight": tied, "c": torch.randn(1000, generator=g)}
    got = broadcast_state_dict(sd, src=0, device="cpu", chunk_elems=300)  # force several buckets
    ok = got["nar_predict_layers.0.weight"].data_ptr() == got["nar_audio_embeddings.2.word_embeddings.weight"].data_ptr()
    chk = float(sum(v.double().sum() for v in got.values()))
    t = max_over_ranks(1.0 + rank)
    q.put((rank, list(got.keys()), chk, ok, t, shard_units(7, rank, world)))
    dist.destroy_process_group()


def test_broadcast_and_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, k0, c0, tied0, t0, s0), (r1, k1, c1, tied1, t1, s1) = res
    assert k0 == k1 == ["a.weight", "nar_audio_embeddings.2.word_embeddings.weight", "b.alpha",
                        "nar_predict_layers.0.weight", "c"]
    assert c0 == c1 and tied0 and tied1
    assert t0 == t1 == 2.0
    assert s0 == [0, 2, 4, 6] and s1 == [1, 3, 5] and sorted(s0 + s1) == list(range(7))


def test_shard_units_edges():
    assert shard_units(0, 0, 4) == []
    assert shard_units(3, 3, 4) == []
    assert [len(shard_units(64, r, 8)) for r in range(8)] == [8] * 8
