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


def _bench_check_main(rank, world, port, q):
    """bench.py's all-rank verification (every rank checks its own bulletins, counts are summed, rank 0 compares the gathered
    multiset with the union) under gloo: rank 1 loses one of its messages in the second case."""
    import collections

    import torch
    import torch.distributed as dist

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    c = bench.Ctx()
    c.torch, c.dist, c.rank, c.world, c.device = torch, dist, rank, world, torch.device("cpu")
    steps = 3
    expect = [(rank * 4 + s, 518 if s % 2 == 0 else 490, "AB%02d" % (rank * 4 + s), "ZCZC AB%02d\nTEXT %d\nNNNN\n" % (rank * 4 + s, s)) for s in range(4)]
    out = []
    for drop in (False, True):
        msgs = [e for _ in range(steps) for e in expect]
        if drop and rank == 1:
            msgs = msgs[:-1]
        ok, multiset, check = bench.verify_all_ranks(c, msgs, expect, steps)
        sums = bench.allreduce_sum(c, [ok, len(expect), multiset])
        out.append((sums, check))
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_bench_all_rank_verification_two_gloo_ranks():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bench_check_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    (sums_ok, check_ok), (sums_bad, check_bad) = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sums_ok == [8, 8, 2] and check_ok["gathered_multiset_equals_union_of_expected"] and check_ok["messages_gathered_all_ranks"] == 24
    # one message missing on rank 1: every bulletin was still seen at least once, but that rank's multiset and the gathered one are off
    assert sums_bad == [8, 8, 1] and not check_bad["gathered_multiset_equals_union_of_expected"] and check_bad["messages_gathered_all_ranks"] == 23
