"""N > 1 host path on CPU: two gloo ranks each own half of the streams, decode their share (oracle events ->
the product's host assembler), and rank 0 gathers the message records.  The merged result must equal the
single-process result: the multiset of (stream, freq, bbbb, message) does not depend on the sharding."""
import os
import socket
import sys

import pytest
import torch.multiprocessing as mp

import cases
import oracle_lib as ol
from navtex_b200 import engine, sharding

NAMES = ["clean518", "noisy490", "dropout", "noise", "weak518", "clean518"]


def decode_streams(lo, hi):
    msgs = []
    for s in range(lo, hi):
        r = ol.run_oracle(cases.build(NAMES[s]), record_taps=False)
        for c, tag in enumerate(ol.CHANNELS):
            msgs += engine.host_assemble(r.events[tag], stream=s, freq=(518, 490)[c])
    return msgs


def _rank_main(rank, world, port, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(len(NAMES), world, rank)
    merged = sharding.gather_messages(decode_streams(lo, hi))
    if rank == 0:
        q.put(merged)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions():
    for total in (0, 1, 5, 1024, 65536):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(8, 2, 2)


def test_two_rank_gloo_gather_matches_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    single = sorted(decode_streams(0, len(NAMES)), key=lambda m: m[0])
    assert merged == single
    assert [m[2] for m in merged] == ["PA12", "QB07", "MK33", "PA12"]      # clean518 twice, dropout's partial message
