"""Sharded multi-GPU host: one process per GPU (torchrun), strings partitioned by bytes.

Tokenizing a string needs nothing from any other string (go/tokenmonster.go:1018-1026: all
state is local), so the batch shards trivially.  Two modes:

* shards already resident on every rank (what bench.py scales): no data-path collective at all;
* one root holds the batch (the server case): ONE scatter of input bytes and ONE gather of token
  ids over NCCL (grouped send/recv = scatterv / gatherv; the vocabulary tables are uploaded by
  every rank from the same file, nothing else is exchanged).  Order is preserved, so the gather
  is a concatenation.

torch.distributed is plumbing only; `compute` is the local tokenizer (the CUDA path on a GPU
rank; the gloo CPU tests inject a stand-in because they exercise the partition / exchange logic,
not the kernels).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_by_bytes(off: torch.Tensor, world: int) -> List[Tuple[int, int]]:
    """Contiguous string ranges [s0, s1) per rank with ~equal bytes, cut at string boundaries."""
    n = off.numel() - 1
    rel = (off - off[0]).to(torch.int64).cpu()
    total = int(rel[-1])
    cuts = [0]
    for r in range(1, world):
        target = total * r // world
        s = int(torch.searchsorted(rel, torch.tensor(target), right=False))
        s = min(max(s, cuts[-1]), n)
        cuts.append(s)
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def _exchange(ops):
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def tokenize_from_root(data: Optional[torch.Tensor], off: Optional[torch.Tensor], compute: Callable,
                       root: int = 0, device=None, group=None):
    """Root holds (data u8, off i64[n+1]); every rank tokenizes its share; root gets everything back.

    compute(data_u8, off_i64) -> (tokens_u8, counts_i32, enc_len): serialized ids of the local strings.
    Returns on root (tokens_u8, tok_off_i64[n+1], enc_len); None elsewhere.
    """
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = device if device is not None else (data.device if data is not None else torch.device("cpu"))
    # 1. plan: string cuts and byte cuts, broadcast as one small tensor
    plan = torch.zeros(2 * (world + 1), dtype=torch.int64, device=dev)
    if rank == root:
        ranges = shard_by_bytes(off, world)
        cuts = [r[0] for r in ranges] + [ranges[-1][1]]
        rel = (off - off[0]).to(torch.int64)
        plan[: world + 1] = torch.tensor(cuts, dtype=torch.int64)
        plan[world + 1:] = rel[torch.tensor(cuts, dtype=torch.int64, device=rel.device)].to(dev)
    dist.broadcast(plan, src=root, group=group)
    cuts = plan[: world + 1].tolist()
    bcuts = plan[world + 1:].tolist()
    my_n = cuts[rank + 1] - cuts[rank]
    my_bytes = bcuts[rank + 1] - bcuts[rank]
    # 2. scatterv of bytes and (rebased) offsets
    if rank == root:
        rel = (off - off[0]).to(torch.int64)
        base = int(off[0])
        ops, keep = [], []
        for r in range(world):
            d = data[base + bcuts[r]: base + bcuts[r + 1]]
            o = (rel[cuts[r]: cuts[r + 1] + 1] - bcuts[r]).contiguous()
            if r == root:
                my_data, my_off = d.contiguous(), o
            else:
                keep += [d.contiguous(), o]
                if d.numel():
                    ops.append(dist.P2POp(dist.isend, keep[-2], r, group))
                ops.append(dist.P2POp(dist.isend, keep[-1], r, group))
        _exchange(ops)
    else:
        my_data = torch.empty(my_bytes, dtype=torch.uint8, device=dev)
        my_off = torch.empty(my_n + 1, dtype=torch.int64, device=dev)
        ops = []
        if my_bytes:
            ops.append(dist.P2POp(dist.irecv, my_data, root, group))
        ops.append(dist.P2POp(dist.irecv, my_off, root, group))
        _exchange(ops)
    # 3. local compute
    toks, counts, enc = compute(my_data, my_off)
    # 4. sizes
    sizes = torch.zeros(world, dtype=torch.int64, device=dev)
    mine = torch.tensor([toks.numel()], dtype=torch.int64, device=dev)
    gathered = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(gathered, mine, group=group)
    sizes = torch.cat(gathered)
    # 5. gatherv of serialized ids and per-string counts
    if rank == root:
        n = cuts[-1]
        all_toks = torch.empty(int(sizes.sum()), dtype=torch.uint8, device=dev)
        all_counts = torch.empty(n, dtype=torch.int32, device=dev)
        starts = torch.cumsum(sizes, 0) - sizes
        ops = []
        for r in range(world):
            a, b = int(starts[r]), int(starts[r] + sizes[r])
            if r == root:
                all_toks[a:b] = toks
                all_counts[cuts[r]:cuts[r + 1]] = counts
            else:
                if b > a:
                    ops.append(dist.P2POp(dist.irecv, all_toks[a:b], r, group))
                if cuts[r + 1] > cuts[r]:
                    ops.append(dist.P2POp(dist.irecv, all_counts[cuts[r]:cuts[r + 1]], r, group))
        _exchange(ops)
        tok_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        tok_off[1:] = torch.cumsum(all_counts.to(torch.int64), 0)
        return all_toks, tok_off, enc
    ops = []
    if toks.numel():
        ops.append(dist.P2POp(dist.isend, toks.contiguous(), root, group))
    if my_n:
        ops.append(dist.P2POp(dist.isend, counts.contiguous(), root, group))
    _exchange(ops)
    return None


def gpu_compute(workspace, enc_len: int = 2):
    """compute() for a GPU rank: the device-resident CUDA path of this rank's workspace."""
    def run(data: torch.Tensor, off: torch.Tensor):
        workspace.launch_tokenize(data, off, enc_len, 0)
        res = workspace.results()
        return res["tokens_u8"], res["counts"].to(torch.int32), res["enc_len"]
    return run


def score_sharded(local_scores: torch.Tensor, local_hist: torch.Tensor, local_tokens: int, device=None, group=None):
    """Training scores over a dataset sharded across ranks (training/trainvocab.go:932-1170 per shard).

    Every rank scores its resident data strips for the same candidate vocabulary
    (`Vocab.score_batch`); the per-token histograms, the histogram of uncovered byte values and
    tokensInText are sums over strings, so ONE all-reduce (n_reverse + 257 int64 values; NCCL over
    NVLink on GPU ranks) gives every rank the totals the master's pruning step needs.
    Returns (scores int64[n_reverse], hist int64[256], tokens_in_text)."""
    dev = device if device is not None else local_scores.device
    n = local_scores.numel()
    vec = torch.empty(n + 257, dtype=torch.int64, device=dev)
    vec[:n] = local_scores.to(device=dev, dtype=torch.int64)
    vec[n:n + 256] = local_hist.to(device=dev, dtype=torch.int64)
    vec[n + 256] = int(local_tokens)
    dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    return vec[:n], vec[n:n + 256], int(vec[n + 256])


def gpu_score(vocab, data, off, device=None, group=None):
    """score_sharded() for a GPU rank: scores this rank's strips with the CUDA path, then reduces."""
    import numpy as np
    scores, hist, tokens, _ = vocab.score_batch(data, off)
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    return score_sharded(torch.from_numpy(scores.astype(np.int64)), torch.from_numpy(hist.astype(np.int64)), tokens,
                         device=dev, group=group)
