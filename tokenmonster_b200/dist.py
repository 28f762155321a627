"""Sharded multi-GPU host: one process per GPU (torchrun), strings partitioned by bytes.

Tokenizing a string needs nothing from any other string (go/tokenmonster.go:1018-1026: all
state is local), so the batch shards trivially.  Two modes:

* shards already resident on every rank (what bench.py scales): no data-path collective at all;
* one root holds the batch (the server case, training/tokenmonsterserver.go:364-377): the root
  SCATTERS the input bytes and GATHERS the token ids over NCCL (send/recv over NVLink; the
  vocabulary tables are uploaded by every rank from the same file, nothing else is exchanged).
  Order is preserved, so the gather is a concatenation.
    - `tokenize_from_root`  one synchronous scatterv, compute, one gatherv (exact sizes; also the
                            fallback of the pipelined form)
    - `RootPipeline`        every rank's share is cut into C byte-balanced chunks; scatter of chunk
                            k+1, the kernels of chunk k and the gather of chunk k-1 run on three
                            streams / two NCCL communicators.  Result messages have a size both
                            sides can compute from the plan (ids are padded to a bound), so the
                            root posts every receive up front and never waits for a size.

torch.distributed is plumbing only; `compute` is the local tokenizer (the CUDA path on a GPU
rank; the gloo CPU tests inject a stand-in because they exercise the partition / exchange logic,
not the kernels).
"""
from __future__ import annotations

import contextlib
from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_by_bytes(off: torch.Tensor, world: int) -> List[Tuple[int, int]]:
    """Contiguous string ranges [s0, s1) per rank with ~equal bytes, cut at string boundaries."""
    n = off.numel() - 1
    rel = (off - off[0]).to(torch.int64).cpu()
    total = int(rel[-1])
    targets = torch.tensor([total * r // world for r in range(1, world)], dtype=torch.int64)
    inner = torch.searchsorted(rel, targets, right=False).tolist() if world > 1 else []
    cuts = [0]
    for s in inner:
        cuts.append(min(max(int(s), cuts[-1]), n))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def _exchange(ops):
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def tokenize_from_root(data: Optional[torch.Tensor], off: Optional[torch.Tensor], compute: Callable,
                       root: int = 0, device=None, group=None):
    """Root holds (data u8, off i64[n+1]); every rank tokenizes its share; root gets everything back.

    compute(data_u8, off_i64) -> (tokens_u8, counts_i32, enc_len[, missing_i32, status_u8]): serialized
    ids of the local strings.  Returns on root (tokens_u8, tok_off_i64[n+1], enc_len, missing, status);
    None elsewhere.  A string the normalizer refused comes back with status 13 and no ids, as from
    Vocab.tokenize.
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
    host_plan = plan.cpu()
    cuts = host_plan[: world + 1].tolist()
    bcuts = host_plan[world + 1:].tolist()
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
                my_data, my_off = d.clone(), o   # a fresh, aligned tensor (the slice may start anywhere)
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
    out = compute(my_data, my_off)
    toks, counts, enc = out[0], out[1], out[2]
    missing = out[3] if len(out) > 3 else torch.zeros(my_n, dtype=torch.int32, device=dev)
    status = out[4] if len(out) > 4 else torch.zeros(my_n, dtype=torch.uint8, device=dev)
    # 4. sizes
    mine = torch.tensor([toks.numel()], dtype=torch.int64, device=dev)
    gathered = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(gathered, mine, group=group)
    sizes = torch.cat(gathered).cpu().tolist()
    # 5. gatherv of serialized ids and per-string counts / missing / status
    if rank == root:
        n = cuts[-1]
        all_toks = torch.empty(int(sum(sizes)), dtype=torch.uint8, device=dev)
        all_counts = torch.empty(n, dtype=torch.int32, device=dev)
        all_missing = torch.empty(n, dtype=torch.int32, device=dev)
        all_status = torch.empty(n, dtype=torch.uint8, device=dev)
        ops, a = [], 0
        for r in range(world):
            b = a + sizes[r]
            c0, c1 = cuts[r], cuts[r + 1]
            if r == root:
                all_toks[a:b] = toks
                all_counts[c0:c1] = counts
                all_missing[c0:c1] = missing
                all_status[c0:c1] = status
            else:
                if b > a:
                    ops.append(dist.P2POp(dist.irecv, all_toks[a:b], r, group))
                if c1 > c0:
                    ops.append(dist.P2POp(dist.irecv, all_counts[c0:c1], r, group))
                    ops.append(dist.P2POp(dist.irecv, all_missing[c0:c1], r, group))
                    ops.append(dist.P2POp(dist.irecv, all_status[c0:c1], r, group))
            a = b
        _exchange(ops)
        tok_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        tok_off[1:] = torch.cumsum(all_counts.to(torch.int64), 0)
        return all_toks, tok_off, enc, all_missing, all_status
    ops = []
    if toks.numel():
        ops.append(dist.P2POp(dist.isend, toks.contiguous(), root, group))
    if my_n:
        ops.append(dist.P2POp(dist.isend, counts.contiguous(), root, group))
        ops.append(dist.P2POp(dist.isend, missing.to(torch.int32).contiguous(), root, group))
        ops.append(dist.P2POp(dist.isend, status.contiguous(), root, group))
    _exchange(ops)
    return None


def gpu_compute(workspace, enc_len: int = 2):
    """compute() for a GPU rank: the device-resident CUDA path of this rank's workspace."""
    def run(data: torch.Tensor, off: torch.Tensor):
        workspace.launch_tokenize(data, off, enc_len, 0)
        res = workspace.results()
        return (res["tokens_u8"], res["counts"].to(torch.int32), res["enc_len"], res["missing"].to(torch.int32),
                res["status"])
    return run


# ------------------------------------------------------------------------------------------------
# pipelined single-root path
# ------------------------------------------------------------------------------------------------
def ids_bound(n_bytes: int, n_strings: int) -> int:
    """Ids a chunk's result message has room for: ~1.5x what English / code text yields (0.21 ids per
    byte on the bench configurations).  A chunk that yields more is fetched again with exact sizes."""
    return n_bytes // 3 + 2 * n_strings + 64


def plan_pieces(off_cpu: torch.Tensor, pieces: int):
    """Byte-balanced cuts of the batch into `pieces` contiguous runs of strings (host arithmetic on
    the offsets: `pieces` binary searches, no pass over the array, no device work)."""
    n = off_cpu.numel() - 1
    off0, total = int(off_cpu[0]), int(off_cpu[-1]) - int(off_cpu[0])
    targets = torch.tensor([off0 + total * p // pieces for p in range(1, pieces)], dtype=off_cpu.dtype)
    inner = torch.searchsorted(off_cpu, targets, right=False) if pieces > 1 else torch.zeros(0, dtype=torch.int64)
    cuts = torch.cat([torch.zeros(1, dtype=torch.int64), torch.clamp(inner, 0, n), torch.tensor([n])])
    cuts = torch.cummax(cuts, 0).values
    return cuts, off_cpu[cuts].to(torch.int64) - off0


def _msg_layout(n_strings: int, n_bytes: int, enc: int):
    """One result message per chunk: [total ids i64][counts i32 x n][missing i32 x n][status u8 x n,
    padded to 8][ids, enc bytes each, padded to the bound].  Returns (offsets, size)."""
    o_counts = 8
    o_missing = o_counts + 4 * n_strings
    o_status = o_missing + 4 * n_strings
    o_tokens = (o_status + n_strings + 7) & ~7
    return (o_counts, o_missing, o_status, o_tokens), o_tokens + ids_bound(n_bytes, n_strings) * enc


class GpuChunkCompute:
    """One DeviceWorkspace per chunk in flight (results stay valid until the gather has read them)."""

    def __init__(self, vocab, enc: int, device):
        self.vocab, self.enc, self.device = vocab, enc, device
        self.ws = []

    def prepare(self, chunks):
        """chunks: list of (n_strings, n_bytes) this rank will tokenize."""
        from .device import DeviceWorkspace
        for k, (n, b) in enumerate(chunks):
            if k >= len(self.ws) or self.ws[k].max_strings < n or self.ws[k].max_bytes < b:
                w = DeviceWorkspace(self.vocab, n + n // 4 + 1024, b + b // 4 + (1 << 16))
                if k < len(self.ws):
                    self.ws[k].close()
                    self.ws[k] = w
                else:
                    self.ws.append(w)

    def run(self, k, data, off, msg, layout, stream):
        """Tokenize chunk k on `stream` and lay its result message out in `msg` (uint8 view)."""
        ws = self.ws[k]
        n = off.numel() - 1
        ws.launch_tokenize(data, off, self.enc, 0, stream=stream)
        (oc, om, os_, ot) = layout
        cap = msg.numel() - ot
        v = ws.raw_views(n, cap)
        msg[0:8].copy_(v["n_tokens_u8"], non_blocking=True)
        if n:
            msg[oc:oc + 4 * n].copy_(v["counts_u8"], non_blocking=True)
            msg[om:om + 4 * n].copy_(v["missing_u8"], non_blocking=True)
            msg[os_:os_ + n].copy_(v["status"], non_blocking=True)
        if cap:
            msg[ot:ot + cap].copy_(v["tokens_u8"], non_blocking=True)

    def exact(self, k, total_ids):
        return self.ws[k].raw_views(0, total_ids * self.enc)["tokens_u8"]


class RootPipeline:
    """Pipelined scatter / tokenize / gather of one root-held batch (see the module docstring).

    run(data, off, off_cpu=None) -> on the root (tokens_u8, tok_off_i64[n+1], enc, missing_i32, status_u8),
    None elsewhere.  `off_cpu` (the offsets on the host, as a request handler has them) avoids one
    device-to-host copy of the offsets for the plan.
    """

    def __init__(self, vocab, ws_unused, enc: int, device, max_strings: int = 0, max_bytes: int = 0, chunks: int = 0,
                 root: int = 0, group=None, compute=None):
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.enc, self.dev, self.root, self.group = enc, torch.device(device), root, group
        self.cuda = self.dev.type == "cuda"
        if not chunks:
            # a chunk has to be >= 2 "waves" of its GPU (one lane per string, ~2000 strings per SM) or the
            # kernels run at a fraction of their rate: measured at N = 2, four chunks of 125 k strings
            # took 11.1 ms where one GPU needs 8.2 ms for the whole batch
            sms = torch.cuda.get_device_properties(self.dev).multi_processor_count if self.cuda else 1
            chunks = max(1, min(4, (max_strings // max(1, self.world)) // (sms * 2048)))
        self.chunks = chunks
        # two more communicators: scatter and gather must not queue behind each other
        self.g_in = dist.new_group(list(range(self.world)))
        self.g_out = dist.new_group(list(range(self.world)))
        self.compute = compute if compute is not None else GpuChunkCompute(vocab, enc, self.dev)
        if self.cuda:
            self.s_in, self.s_out = torch.cuda.Stream(self.dev), torch.cuda.Stream(self.dev)
        self.in_data = torch.empty(0, dtype=torch.uint8, device=self.dev)
        self.in_off = torch.empty(0, dtype=torch.int64, device=self.dev)
        self.stage = torch.empty(0, dtype=torch.uint8, device=self.dev)

    def _stream(self, s):
        return torch.cuda.stream(s) if self.cuda else contextlib.nullcontext()

    def run(self, data, off, off_cpu=None):
        import os
        import time
        trace = bool(os.environ.get("TMGPU_TRACE")) and self.cuda
        marks = []

        def mark(what):   # diagnostic: wall-clock phase times with the device drained at every mark
            if trace:
                torch.cuda.synchronize(self.dev)
                marks.append((what, time.perf_counter()))
        mark("start")
        W, C, root, rank, dev = self.world, self.chunks, self.root, self.rank, self.dev
        P = W * C
        # ---- plan: one broadcast, one small D2H ------------------------------------------------
        plan = torch.zeros(2 * (P + 1), dtype=torch.int64, device=dev)
        if rank == root:
            oc = off_cpu if off_cpu is not None else off.cpu()
            cuts, bcuts = plan_pieces(oc, P)
            plan.copy_(torch.cat([cuts, bcuts]))
        dist.broadcast(plan, src=root, group=self.group)
        hp = plan.cpu().tolist()
        cuts, bcuts = hp[: P + 1], hp[P + 1:]
        pieces = [(cuts[p + 1] - cuts[p], bcuts[p + 1] - bcuts[p]) for p in range(P)]
        layouts = [_msg_layout(n, b, self.enc) for (n, b) in pieces]
        mine = range(rank * C, (rank + 1) * C)
        self.compute.prepare([pieces[p] for p in mine])
        cur = torch.cuda.current_stream(dev) if self.cuda else None
        mark("plan")
        # ---- staging for the result messages (root: all pieces; others: their own) --------------
        keep = range(P) if rank == root else mine
        sizes = {p: layouts[p][1] for p in keep}
        starts, at = {}, 0
        for p in keep:
            starts[p] = at
            at += (sizes[p] + 255) & ~255
        if self.stage.numel() < at:
            self.stage = torch.empty(at + at // 8, dtype=torch.uint8, device=dev)
        msg = {p: self.stage[starts[p]: starts[p] + sizes[p]] for p in keep}
        # ---- scatter ---------------------------------------------------------------------------
        recv_in, sends = [], []
        if rank == root:
            b0 = int(off_cpu[0]) if off_cpu is not None else int(off[:1].item())
            if self.cuda:
                self.s_in.wait_stream(cur)
            with self._stream(getattr(self, "s_in", None)):
                ops, keep = [], []
                for r in range(W):   # every rank's offsets, rebased to the start of its share
                    if r == root:
                        continue
                    o = (off[cuts[r * C]: cuts[(r + 1) * C] + 1] - (b0 + bcuts[r * C])).contiguous()
                    keep.append(o)
                    ops.append(dist.P2POp(dist.isend, o, r, self.g_in))
                for w in (dist.batch_isend_irecv(ops) if ops else []):
                    sends.append((keep, w))
                for k in range(C):   # then the bytes, chunk by chunk: ONE grouped NCCL launch per chunk
                    ops = []
                    for r in range(W):
                        p = r * C + k
                        if r == root or pieces[p][1] == 0:
                            continue
                        d = data[b0 + bcuts[p]: b0 + bcuts[p + 1]]
                        ops.append(dist.P2POp(dist.isend, d, r, self.g_in))
                    for w in (dist.batch_isend_irecv(ops) if ops else []):
                        sends.append((data, w))
            my_data, my_off, origin = data, off, 0
        else:
            nb = bcuts[(rank + 1) * C] - bcuts[rank * C]
            ns = cuts[(rank + 1) * C] - cuts[rank * C]
            if self.in_data.numel() < nb + 64:
                self.in_data = torch.empty(nb + nb // 8 + 64, dtype=torch.uint8, device=dev)
            if self.in_off.numel() < ns + 1:
                self.in_off = torch.empty(ns + ns // 8 + 1, dtype=torch.int64, device=dev)
            with self._stream(getattr(self, "s_in", None)):
                w_off = dist.irecv(self.in_off[: ns + 1], root, group=self.g_in)
                for k in range(C):
                    p = rank * C + k
                    a, b = bcuts[p] - bcuts[rank * C], bcuts[p + 1] - bcuts[rank * C]
                    recv_in.append(dist.irecv(self.in_data[a:b], root, group=self.g_in) if b > a else None)
            my_data, my_off, origin = self.in_data, self.in_off, cuts[rank * C]
        mark("scatter posted")
        # ---- the root posts every gather receive up front --------------------------------------
        recv_out = []
        if rank == root:
            with self._stream(getattr(self, "s_out", None)):
                for k in range(C):
                    ops = [dist.P2POp(dist.irecv, msg[r * C + k], r, self.g_out) for r in range(W) if r != root]
                    if ops:
                        recv_out += dist.batch_isend_irecv(ops)
        # ---- compute chunk by chunk; each result goes out as soon as it exists ------------------
        sent = []
        if rank != root:
            w_off.wait()
        for k in range(C):
            p = rank * C + k
            if rank != root and recv_in[k] is not None:
                recv_in[k].wait()          # the compute stream waits for this chunk's bytes only
            o = my_off[cuts[p] - origin: cuts[p + 1] - origin + 1]
            self.compute.run(k, my_data, o, msg[p], layouts[p][0], cur)
            if rank != root:
                if self.cuda:
                    self.s_out.wait_stream(cur)
                with self._stream(getattr(self, "s_out", None)):
                    sent.append(dist.isend(msg[p], root, group=self.g_out))
        mark("computed")
        for w in sent:
            w.wait()
        for _, w in sends:
            w.wait()
        if self.cuda and rank == root:
            cur.wait_stream(self.s_in)
        mark("sends done")
        # ---- overflow check (a chunk with more ids than its message has room for) ---------------
        tot = torch.stack([msg[p][0:8].view(torch.int64)[0] for p in mine]) if C else torch.zeros(0, dtype=torch.int64, device=dev)
        over = torch.stack([(tot[k] > ids_bound(pieces[rank * C + k][1], pieces[rank * C + k][0])).to(torch.int64)
                            for k in range(C)]).max().reshape(1)
        dist.all_reduce(over, op=dist.ReduceOp.MAX, group=self.group)
        if rank != root:
            if int(over.item()):
                return self._exact(data, off)
            return None
        for w in recv_out:
            w.wait()
        mark("gathered")
        if int(over.item()):
            return self._exact(data, off)
        # ---- assemble: one D2H of the P totals, then slices ------------------------------------
        totals = torch.stack([msg[p][0:8].view(torch.int64)[0] for p in range(P)]).cpu().tolist()
        n = cuts[-1]
        toks, counts, missing, status = [], [], [], []
        for p in range(P):
            (oc, om, os_, ot) = layouts[p][0]
            np_ = pieces[p][0]
            toks.append(msg[p][ot: ot + totals[p] * self.enc])
            counts.append(msg[p][oc: oc + 4 * np_].view(torch.int32))
            missing.append(msg[p][om: om + 4 * np_].view(torch.int32))
            status.append(msg[p][os_: os_ + np_])
        all_counts = torch.cat(counts)
        tok_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        tok_off[1:] = torch.cumsum(all_counts.to(torch.int64), 0)
        out = torch.cat(toks), tok_off, self.enc, torch.cat(missing), torch.cat(status)
        mark("assembled")
        if trace:
            import sys
            print("[RootPipeline] " + ", ".join(f"{b[0]} +{(b[1] - a[1]) * 1e3:.2f} ms" for a, b in zip(marks, marks[1:])),
                  file=sys.stderr, flush=True)
        return out

    def _exact(self, data, off):
        """Fallback with exact message sizes (the synchronous scatterv / gatherv)."""
        comp = getattr(self.compute, "whole", None)
        if comp is None:
            from .device import DeviceWorkspace
            vocab = self.compute.vocab
            holder = {}

            def comp(d, o):
                n, b = o.numel() - 1, int(d.numel())
                w = holder.get("ws")
                if w is None or w.max_strings < n or w.max_bytes < b:
                    holder["ws"] = w = DeviceWorkspace(vocab, n + 1024, b + (1 << 16))
                return gpu_compute(w, self.enc)(d, o)
        return tokenize_from_root(data, off, comp, self.root, self.dev, self.group)


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
