"""world_size-2 gloo tests of the sharded host logic (partition, scatterv, gatherv) on CPU.
The local compute is a stand-in (2 'ids' per input byte pair) -- the kernels are covered by the
-m gpu tests; this file checks that sharding preserves order and content."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tokenmonster_b200.dist import RootPipeline, shard_by_bytes, tokenize_from_root


def fake_compute(data, off):
    """Deterministic stand-in tokenizer: one 2-byte id per input byte (id = byte value * 3 + 1)."""
    ids = (data.to(torch.int32) * 3 + 1).to(torch.int16)
    toks = ids.view(torch.uint8) if ids.numel() else torch.zeros(0, dtype=torch.uint8)
    counts = (off[1:] - off[:-1]).to(torch.int32)
    return toks, counts, 2


def _worker(rank, world, port, lens, seed, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(seed)
        off = torch.zeros(len(lens) + 1, dtype=torch.int64)
        off[1:] = torch.cumsum(torch.tensor(lens, dtype=torch.int64), 0)
        off += 7                                            # offsets need not start at 0
        data = torch.randint(0, 255, (int(off[-1]),), dtype=torch.uint8, generator=g)
        out = tokenize_from_root(data if rank == 0 else None, off if rank == 0 else None, fake_compute,
                                 root=0, device=torch.device("cpu"))
        if rank == 0:
            toks, tok_off, enc = out[:3]
            want, want_off, _ = fake_compute(data[7:], off - 7)
            ok = torch.equal(toks, want) and torch.equal(tok_off, (off - 7)) and enc == 2
            q.put(bool(ok))
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


def sparse_compute(data, off, stride):
    """Stand-in tokenizer with a realistic id density: one 2-byte id per `stride` bytes of each string
    (the bytes at positions 0, stride, 2*stride, ...), missing = 2, status 13 for empty strings."""
    lens = (off[1:] - off[:-1]).tolist()
    base = int(off[0])
    pieces, counts, at = [], [], 0
    for l in lens:
        seg = data[base + at: base + at + l: stride]
        pieces.append((seg.to(torch.int32) * 3 + 1).to(torch.int16))
        counts.append(seg.numel())
        at += l
    ids = torch.cat(pieces) if pieces else torch.zeros(0, dtype=torch.int16)
    toks = ids.view(torch.uint8) if ids.numel() else torch.zeros(0, dtype=torch.uint8)
    n = len(lens)
    return (toks, torch.tensor(counts, dtype=torch.int32), 2, torch.full((n,), 2, dtype=torch.int32),
            torch.tensor([13 if l == 0 else 0 for l in lens], dtype=torch.uint8))


class FakeChunkCompute:
    """Stand-in for GpuChunkCompute, laid out as the real one does."""

    def __init__(self, stride):
        self.stride = stride
        self.whole = lambda d, o: sparse_compute(d, o, stride)

    def prepare(self, chunks):
        self.chunks = chunks

    def run(self, k, data, off, msg, layout, stream):
        oc, om, os_, ot = layout
        n = off.numel() - 1
        raw, counts, _, missing, status = sparse_compute(data, off, self.stride)
        msg[0:8].copy_(torch.tensor([raw.numel() // 2], dtype=torch.int64).view(torch.uint8))
        if n:
            msg[oc:oc + 4 * n].copy_(counts.view(torch.uint8))
            msg[om:om + 4 * n].copy_(missing.view(torch.uint8))
            msg[os_:os_ + n].copy_(status)
        m = min(msg.numel() - ot, raw.numel())
        msg[ot:ot + m].copy_(raw[:m])


def _pipe_worker(rank, world, port, lens, chunks, stride, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(11)
        off = torch.zeros(len(lens) + 1, dtype=torch.int64)
        off[1:] = torch.cumsum(torch.tensor(lens, dtype=torch.int64), 0)
        off += 5
        data = torch.randint(0, 255, (int(off[-1]),), dtype=torch.uint8, generator=g)
        pipe = RootPipeline(None, None, 2, torch.device("cpu"), chunks=chunks, compute=FakeChunkCompute(stride))
        for _ in range(2):   # buffers are reused across calls
            out = pipe.run(data if rank == 0 else None, off if rank == 0 else None, off_cpu=off if rank == 0 else None)
        if rank == 0:
            toks, tok_off, enc, missing, status = out
            want, wc, _, wm, ws = sparse_compute(data, off, stride)
            woff = torch.zeros(len(lens) + 1, dtype=torch.int64)
            woff[1:] = torch.cumsum(wc.to(torch.int64), 0)
            q.put(bool(torch.equal(toks, want) and torch.equal(tok_off, woff) and enc == 2 and torch.equal(missing, wm)
                       and torch.equal(status, ws)))
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


# stride 4: ids fit the padded result messages (the pipelined path proper); stride 1: one id per byte is
# more than the messages hold, so the exact-size fallback must take over
@pytest.mark.parametrize("lens,chunks,stride", [([5, 0, 17, 300, 1, 1, 64, 9], 2, 4), ([0, 0, 0], 2, 4), ([1000], 3, 4),
                                                ([3] * 50, 4, 4), ([40, 7, 0, 90, 300, 300, 2], 2, 1)])
def test_root_pipeline_two_ranks(lens, chunks, stride):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() + len(lens) * 13 + chunks) % 2000
    procs = [ctx.Process(target=_pipe_worker, args=(r, 2, port, lens, chunks, stride, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


@pytest.mark.parametrize("lens", [[5, 0, 17, 300, 1, 1, 64, 9], [0, 0, 0], [1000], [3] * 50])
def test_root_scatter_gather_two_ranks(lens):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + len(lens) * 7) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lens, 5, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_shard_by_bytes_is_balanced_and_ordered():
    rng = np.random.default_rng(1)
    lens = rng.integers(0, 2000, 5000)
    off = torch.zeros(5001, dtype=torch.int64)
    off[1:] = torch.cumsum(torch.tensor(lens), 0)
    for world in (1, 2, 4, 8):
        ranges = shard_by_bytes(off, world)
        assert ranges[0][0] == 0 and ranges[-1][1] == 5000
        assert all(ranges[k][1] == ranges[k + 1][0] for k in range(world - 1))
        sizes = [int(off[b] - off[a]) for a, b in ranges]
        assert max(sizes) - min(sizes) <= 2 * 2000


def _score_worker(rank, world, port, q):
    """Each rank scores its own shard with the oracle's worker loop (the stand-in for the CUDA
    path, which tests/test_training_scores.py covers); the all-reduce must give both ranks the
    totals of the whole set."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle
        from tokenmonster_b200.dist import score_sharded
        from workloads import corpus
        with open(os.path.join(os.path.dirname(__file__), "golden", "small_cap2.vocab"), "rb") as f:
            chk = oracle.OracleVocab(f.read(), "port")
        base = bytes(corpus.base_text(8, 1 << 16, 0.4))
        strips = [base[k * 500:(k + 1) * 500 - (k % 7)] for k in range(60)]
        mine = strips[rank::world]
        s, h, t = chk.score(mine)
        gs, gh, gt = score_sharded(torch.from_numpy(s.astype(np.int64)), torch.from_numpy(h.astype(np.int64)), t,
                                   device=torch.device("cpu"))
        ws, wh, wt = chk.score(strips)
        q.put(bool(np.array_equal(gs.numpy(), ws.astype(np.int64)) and np.array_equal(gh.numpy(), wh.astype(np.int64))
                   and gt == wt and wt > 0))
    finally:
        dist.destroy_process_group()


def test_training_scores_all_reduce_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_score_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True and q.get(timeout=5) is True
