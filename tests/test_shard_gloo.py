"""CPU, world_size 2 (gloo): the multi-GPU host logic -- newline-aligned sharding, the
line-base / totals exchange and the ordered gather of records -- with the oracle
standing in for the per-rank GPU scan (allowed here: tests may use oracle/)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, path, out_path):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import pyoracle
    from seeq_b200 import binding as B, shard
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    buf = np.fromfile(path, dtype=np.uint8)
    b, e = shard.shard_ranges(buf, world)[rank]
    assert (b, e) == B.shard_range(buf, rank, world)
    orc = pyoracle.Oracle()
    keys, _ = orc.parse("GATCGGAAGAGC")
    recs, nl, nm = orc.buffer_scan(buf[b:e], keys, 2, pyoracle.SQ_ALL)
    local = np.zeros(recs.shape[0], dtype=B.REC_DTYPE)
    local["line"], local["start"], local["end"], local["dist"] = recs[:, 0] - 1, recs[:, 1], recs[:, 2], recs[:, 3]
    base, tot_lines, tot_matched, tot_recs = shard.exchange(nl, nm, recs.shape[0], dist)
    merged = shard.gather_records(local, base, dist)
    if rank == 0:
        np.savez(out_path, merged=merged, totals=np.array([tot_lines, tot_matched, tot_recs]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_exchange_and_gather(tmp_path, oracle, world):
    from seeq_b200 import binding as B
    g = B.make_gen(seed=9, line_len=120, plant="GATCGGAAGAGC", plant_per_1024=250, max_edits=2)
    buf = B.gen_host(g, 5000)
    buf = np.concatenate([buf, np.frombuffer(b"TTGATCGGAAGAGCTT", dtype=np.uint8)])   # no final newline
    path = str(tmp_path / "reads.txt")
    buf.tofile(path)
    out = str(tmp_path / "merged.npz")
    mp.spawn(_worker, args=(world, _free_port(), path, out), nprocs=world, join=True)
    keys, _ = oracle.parse("GATCGGAAGAGC")
    exp, nl, nm = oracle.buffer_scan(buf, keys, 2, 2)
    got = np.load(out)
    assert list(got["totals"]) == [nl, nm, exp.shape[0]]
    exp0 = exp.astype(np.int64)
    exp0[:, 0] -= 1
    assert np.array_equal(got["merged"], exp0)
