#!/usr/bin/env python
"""bench.py -- MB/s tokenized (input bytes) on the BASELINE.json configurations.

    python bench.py --gpus N --steps K --warmup W            (this repo's CUDA path, config c2)
    python bench.py --impl reference --gpus N --steps K ...  (the reference's own C++ CPU runtime)
    torchrun --nproc-per-node N bench.py --gpus N ...        (one rank per GPU, weak scaling)
    python bench.py --config c3|c4|c5 [--gpus N]             (the other named configurations)

A "step" = one pass of the hot path (capcode text, position table, ungreedy chain, offset scan, id
packing) over one batch of synthetic strings resident in HBM before the timed region.  Every batch
is far larger than the 126 MB L2, so no flush is needed between steps.
`value`   : whole-job MB/s (10^6 B) with inputs already in HBM, device-timed, max over ranks.
`e2e`     : the same metric through the C-ABI host-buffer call tmgpu_tokenize_batch with pinned HOST
            buffers -- H2D of the strings and D2H of the ids inside the timed region; `e2e.pageable`
            repeats it with ordinary (pageable) host memory, which is what a cgo caller passes.
`roofline`: the tokenize pipeline against the measured HBM copy bandwidth; algorithmic bytes = 1 B
            read per input byte; per-stage times from CUDA events inside the library.
`parity`  : checksum of the ids of the CPU sample (per-string FNV-1a, summed) against the same fold of
            the reference runtime's own output.
`decode`  : Vocab.Decode of the ids just produced (ids read + bytes written against HBM).
`cpu_baseline`: oracle/_ref (the reference's unmodified C++ runtime) on the box's host cores, on a
            bounded sample of the same batch.
Prints exactly one JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "MB/s tokenized (input bytes) at 32K vocab, 1/2/4/8 B200; bit-exact token IDs"
UNIT = "MB/s"
WORKLOAD = "c2"
# BASELINE.json configs[k]: which one, on how many GPUs it is named, id width of the server path
NAMED = {
    "c2": dict(idx=1, gpus=1, enc=2, name="englishcode-32000-consistent-v1"),
    "c3": dict(idx=2, gpus=1, enc=2, name="gpt2 imported vocab (50257)"),
    "c4": dict(idx=3, gpus=4, enc=2, name="code-65536-unfiltered-nocapcode-v1"),
    "c5": dict(idx=4, gpus=8, enc=4, name="englishcode-100256-clean-v1"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--strings", type=int, default=0, help="strings per GPU (default: the config's share)")
    ap.add_argument("--config", default=WORKLOAD, choices=sorted(NAMED))
    ap.add_argument("--e2e-steps", type=int, default=5, help="timed steps of the host-buffer (e2e) legs, >= 1")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-decode", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for cpu_baseline / parity")
    args = ap.parse_args()
    args.e2e_steps = max(1, args.e2e_steps)
    return args


def strings_per_gpu(args):
    from workloads import configs
    if args.strings:
        return args.strings
    return configs.CONFIGS[args.config]["n"] // NAMED[args.config]["gpus"]


def config_block(args, n_strings, total_bytes, world):
    from workloads import configs
    cfg, nm = configs.CONFIGS[args.config], NAMED[args.config]
    return {
        "workload": f"{cfg['vocab']} (stand-in for BASELINE configs[{nm['idx']}] {nm['name']}), "
                    f"{n_strings} strings/GPU ~U[{cfg['lo']},{cfg['hi']}] B, ASCII prose+code",
        "baseline_config": nm["idx"],
        "strings_per_gpu": n_strings,
        "bytes_per_gpu": int(total_bytes),
        "vocab_ids": cfg["ids"],
        "capcode": cfg["capcode"],
        "enc_len": nm["enc"],
        "parallelism": f"strings sharded over {world} GPU(s), no data-path collective",
        "l2_policy": f"inputs larger than L2 ({total_bytes / 1e6:.0f} MB/step vs 126 MB), no flush",
        "mb": "1e6 bytes",
    }


# ------------------------------------------------------------------------------------------------
def clocks_sampler(gpu_index: int, stop: threading.Event, out: list):
    q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    try:
        p = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                              "-i", str(gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except OSError:
        return

    def reader():
        for line in p.stdout:
            out.append(line.strip())
    t = threading.Thread(target=reader, daemon=True)
    t.start()
    stop.wait()
    p.terminate()
    try:
        p.wait(timeout=2)
    except Exception:
        p.kill()


def summarize_clocks(lines):
    sm, mx, reasons = [], 0, set()
    for ln in lines:
        f = [x.strip() for x in ln.split(",")]
        if len(f) < 9:
            continue
        try:
            sm.append(float(f[1]))
            mx = max(mx, float(f[2]))
        except ValueError:
            continue
        for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
            if v.lower().startswith("active"):
                reasons.add(name)
    if not sm:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    sm.sort()
    return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_capture(config):
    """Numbers that only a profiler can give (dram bytes, warp instructions per launch), from the
    committed ncu --set full capture of the same command (profiles/traffic.json), keyed by config."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(p) as f:
            t = json.load(f)
        return t.get(config) if isinstance(t.get(config), dict) else None
    except Exception:
        return None


def host_cores():
    """(threads used, what the OS says, what this process may actually use)."""
    logical = os.cpu_count() or 1
    try:
        affinity = len(os.sched_getaffinity(0))
    except Exception:
        affinity = logical
    quota = None
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            a, b = f.read().split()
            if a != "max":
                quota = float(a) / float(b)
    except Exception:
        pass
    eff = affinity if quota is None else max(1, min(affinity, int(quota + 0.5)))
    return eff, {"logical": logical, "affinity": affinity, "cgroup_quota": quota, "effective": eff}


# ------------------------------------------------------------------------------------------------
def cpu_reference_rate(vocab_bytes, data_np, off_np, budget_s=12.0, threads=None):
    """Time oracle/_ref (reference C++ runtime) on `threads` host cores on a bounded prefix of the
    batch.  Returns the rate, the prefix length and the checksum of its ids."""
    from oracle import oracle
    kind = "reference" if oracle.have_reference() else "port"
    eff, cores = host_cores()
    threads = threads or eff
    n = len(off_np) - 1
    if kind == "reference":
        v = oracle.OracleVocab(vocab_bytes, "reference")
        probe = min(n, 2000)
        for _ in range(2):   # warm-ups (page in the tables, spin up the threads)
            v.tokenize_batch_timed(data_np, off_np[:probe + 1], threads)
        t = time.perf_counter()
        v.tokenize_batch_timed(data_np, off_np[:probe + 1], threads)
        dt = max(time.perf_counter() - t, 1e-4)
        rate = float(off_np[probe] - off_np[0]) / dt
        take = int(min(n, max(probe, rate * budget_s / (float(off_np[-1] - off_np[0]) / n))))
        t = time.perf_counter()
        ids, chk = v.tokenize_batch_timed(data_np, off_np[:take + 1], threads)
        dt = time.perf_counter() - t
        nbytes = float(off_np[take] - off_np[0])
        # single thread, pinned to one CPU, >= 2 s after warm-ups (tokenmonster-cpp/tests/bench.cpp:56-80)
        one_rate = None
        try:
            old = os.sched_getaffinity(0)
            os.sched_setaffinity(0, {sorted(old)[0]})
        except Exception:
            old = None
        try:
            one = min(take, 400)
            for _ in range(5):
                v.tokenize_batch_timed(data_np, off_np[:one + 1], 1)
            t1 = time.perf_counter()
            done_bytes, reps = 0.0, 0
            while time.perf_counter() - t1 < 2.0:
                v.tokenize_batch_timed(data_np, off_np[:one + 1], 1)
                done_bytes += float(off_np[one] - off_np[0])
                reps += 1
            one_rate = done_bytes / (time.perf_counter() - t1) / 1e6
        finally:
            if old is not None:
                os.sched_setaffinity(0, old)
        return {"value": nbytes / dt / 1e6, "unit": UNIT, "cores": threads, "cores_detail": cores, "kind": kind,
                "sample": f"first {take} strings ({nbytes / 1e6:.1f} MB) of the same batch, {dt:.1f} s wall, one std::thread "
                          f"per contiguous slice (training/tokenmonsterserver.go:364-377 analogue)",
                "single_thread_value": one_rate, "ids": int(ids), "strings": take, "checksum": int(chk)}
    v = oracle.OracleVocab(vocab_bytes, "port")
    take, nbytes = 0, 0
    t = time.perf_counter()
    while take < n and time.perf_counter() - t < budget_s:
        v.tokenize(data_np[int(off_np[take]):int(off_np[take + 1])].tobytes())
        nbytes += int(off_np[take + 1] - off_np[take])
        take += 1
    dt = time.perf_counter() - t
    return {"value": nbytes / dt / 1e6, "unit": UNIT, "cores": 1, "cores_detail": cores, "kind": "port",
            "sample": f"first {take} strings ({nbytes / 1e6:.1f} MB), scalar C port, 1 thread", "strings": take,
            "checksum": None}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation on the host cores, on the batch the
    GPU arm tokenizes (same generator, same seed: rank 0's shard)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np
    from workloads import configs, corpus
    cfg = configs.CONFIGS[args.config]
    n = strings_per_gpu(args)
    vb = configs.vocab_bytes(args.config)
    base = corpus.base_text(cfg["seed"], 16 << 20, cfg["code"])
    # the same strings as the GPU arm's rank 0 (the generator is deterministic in (n, seed)); each
    # step times a bounded prefix of it
    data, off = corpus.make_batch(base, n, cfg["lo"], cfg["hi"], cfg["seed"] + 1, "cpu")
    data_np, off_np = data.numpy(), off.numpy().astype(np.uint64)
    total_bytes = int(off_np[-1])
    threads, _ = host_cores()
    per_step_budget = max(2.0, min(20.0, 150.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_reference_rate(vb, data_np, off_np, budget_s=min(1.0, per_step_budget), threads=threads)
    vals, last = [], None
    t0 = time.perf_counter()
    for _ in range(args.steps):
        last = cpu_reference_rate(vb, data_np, off_np, budget_s=per_step_budget, threads=threads)
        vals.append(last["value"])
    wall = time.perf_counter() - t0
    value = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": wall * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": config_block(args, n, total_bytes, max(1, args.gpus)),   # the GPU arm's config
        "cpu_baseline": {k: last[k] for k in ("value", "unit", "cores", "cores_detail", "kind", "sample",
                                              "single_thread_value") if k in last},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    line["cpu_baseline"]["value"] = value
    emit_line(line)
    return 0


# ------------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def quiet_stdout():
    """Route everything libraries print on fd 1 (e.g. NCCL's version banner) to stderr; the one JSON
    line goes to the real stdout through emit_line()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit_line(obj):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


def time_host_calls(vocab, h_np, h_off, enc, steps, dev):
    import torch
    r = vocab.tokenize_batch(h_np, h_off, enc)  # warm-up (allocations, pinned pools)
    r = None
    r = vocab.tokenize_batch(h_np, h_off, enc)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(steps):
        r = None  # hand the previous result's pinned buffers back to the pool first
        r = vocab.tokenize_batch(h_np, h_off, enc)
    torch.cuda.synchronize(dev)
    return (time.perf_counter() - t0) / steps, r


def main():
    args = parse()
    quiet_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE {world}", file=sys.stderr)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from tokenmonster_b200 import Vocab
    from tokenmonster_b200.device import DeviceWorkspace
    from workloads import configs, corpus

    cfg = configs.CONFIGS[args.config]
    n = strings_per_gpu(args)
    enc = NAMED[args.config]["enc"]
    if rank == 0:
        vb = configs.vocab_bytes(args.config)  # builds + caches the stand-in once
    if world > 1:
        dist.barrier()
    vb = configs.vocab_bytes(args.config)
    vocab = Vocab(vb, device=local)

    # each rank owns its own shard of the (weakly scaled) batch: same shape, different seed
    base = corpus.base_text(cfg["seed"], 16 << 20, cfg["code"])
    data, off = corpus.make_batch(base, n, cfg["lo"], cfg["hi"], cfg["seed"] + 1 + 1000 * rank, dev)
    total_bytes = int(off[-1].item())
    ws = DeviceWorkspace(vocab, n, total_bytes)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---- kernel-resident timing ---------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        ws.launch_tokenize(data, off, enc, 0)
    sync_all()
    ws.kernel_times_ms()
    stop, samples = threading.Event(), []
    th = threading.Thread(target=clocks_sampler, args=(local, stop, samples), daemon=True)
    th.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for _ in range(args.steps):
        ws.launch_tokenize(data, off, enc, 0)
    e1.record()
    sync_all()
    elapsed_ms = e0.elapsed_time(e1)
    launches = ws.launches() * args.steps
    ktimes = [t for t in ws.kernel_times_ms() if t > 0]
    stages = ws.stage_times_ms()
    res = ws.results()
    n_tokens = res["n_tokens"]
    bad_status = int((res["status"] != 0).sum().item())

    # ---- decode of the ids just produced (device resident) ------------------------------------
    decode_line = None
    if not args.no_decode:
        toks = res["tokens_u8"].clone()
        toff = res["tok_off"].clone()
        for _ in range(2):
            out_b, out_off = ws.decode(toks, toff, enc)
        torch.cuda.synchronize(dev)
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(2, min(args.steps, 5))
        d0.record()
        for _ in range(reps):
            out_b, out_off = ws.decode(toks, toff, enc)
        d1.record()
        torch.cuda.synchronize(dev)
        dec_ms = d0.elapsed_time(d1) / reps
        same = bool(out_b.numel() == data.numel() and torch.equal(out_b, data)) if (cfg["capcode"] == 0 or True) else None
        decode_line = {"ms": dec_ms, "text_bytes": int(out_b.numel()), "ids": int(n_tokens),
                       "value": out_b.numel() / (dec_ms / 1e3) / 1e6, "unit": "MB/s of text",
                       "round_trip_identical": same, "launches": ws.launches()}
        del toks, toff, out_b, out_off

    # ---- end to end through the C-ABI host call ------------------------------------------------
    h_data = torch.empty(total_bytes, dtype=torch.uint8, pin_memory=True)
    h_data.copy_(data)
    h_off = off.cpu().numpy().astype(np.uint64)
    h_np = h_data.numpy()
    sync_all()
    e2e_s, r = time_host_calls(vocab, h_np, h_off, enc, args.e2e_steps, dev)
    d2h = r.raw.nbytes + r.tok_off.nbytes + r.missing.nbytes + r.status.nbytes
    h2d = total_bytes + h_off.nbytes
    stop.set()
    th.join(timeout=3)
    p_np = np.empty(total_bytes, dtype=np.uint8)   # ordinary (pageable) memory: what a cgo caller passes
    p_np[:] = h_np
    page_s, rp = time_host_calls(vocab, p_np, h_off, enc, max(1, args.e2e_steps // 2), dev)
    rp = None
    del p_np

    # ---- N > 1: the single-root path (pipelined NCCL scatter of bytes / gather of ids) ---------
    root_line = None
    if world > 1:
        from tokenmonster_b200.dist import RootPipeline
        pipe = RootPipeline(vocab, ws, enc, dev, n, total_bytes)
        off_host = off.cpu() if rank == 0 else None   # a request handler has the offsets on the host
        for _ in range(2):
            out = pipe.run(data if rank == 0 else None, off if rank == 0 else None, off_host)
        sync_all()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        r0.record()
        for _ in range(reps):
            out = pipe.run(data if rank == 0 else None, off if rank == 0 else None, off_host)
        r1.record()
        sync_all()
        tr = torch.tensor([r0.elapsed_time(r1) / reps], dtype=torch.float64, device=dev)
        dist.all_reduce(tr, op=dist.ReduceOp.MAX)
        if rank == 0:
            same = bool(torch.equal(out[0], res["tokens_u8"]))   # identical bytes to the 1-GPU result
            one_ms = (sum(ktimes) / len(ktimes)) if ktimes else None
            root_line = {"value": total_bytes / (float(tr[0]) / 1e3) / 1e6, "unit": UNIT, "ms": float(tr[0]),
                         "what": "rank 0's batch scattered over all ranks by bytes in chunks, tokenized, ids gathered back "
                                 "(NCCL send/recv, scatter k+1 || kernels k || gather k-1); strong scaling of one batch",
                         "identical_to_1gpu": same, "chunks": pipe.chunks,
                         "nvlink_bound_ms": (total_bytes * (world - 1) / world / 770e9 + n_tokens * enc * (world - 1) / world / 770e9) * 1e3,
                         "strong_scaling_efficiency": (one_ms / (float(tr[0]) * world)) if one_ms else None}

    # ---- reduce over ranks --------------------------------------------------------------------
    t = torch.tensor([elapsed_ms, e2e_s, float(total_bytes), float(n_tokens), page_s], dtype=torch.float64, device=dev)
    if world > 1:
        mx = t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = t.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        elapsed_ms, e2e_s, page_s = float(mx[0]), float(mx[1]), float(mx[4])
        all_bytes, all_tokens = float(sm[2]), float(sm[3])
    else:
        all_bytes, all_tokens = float(total_bytes), float(n_tokens)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    value = all_bytes * args.steps / (elapsed_ms / 1e3) / 1e6
    peak, peak_src = hbm_peak()
    k_ms = sum(ktimes) / len(ktimes) if ktimes else None
    achieved = (total_bytes / (k_ms / 1e3) / 1e9) if k_ms else None
    cap = ncu_capture(args.config)
    clocks = summarize_clocks(samples)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": config_block(args, n, total_bytes, world),
        "e2e": {"value": all_bytes / e2e_s / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "api": "tmgpu_tokenize_batch (C ABI, pinned host buffers)", "steps": args.e2e_steps,
                "pageable": {"value": all_bytes / page_s / 1e6, "unit": UNIT,
                             "what": "same call, input in ordinary pageable host memory (what cgo passes)"}},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None,
                     "traffic": cap.get("dram_bytes_per_launch") if cap else None,
                     "traffic_source": cap.get("source") if cap else None,
                     "kernel": "tokenize pipeline: prep + scan + length sort, table_kernel (capcode encoder + automaton), chain_kernel "
                               "(CUDA events around all of them, on the launching stream)",
                     "kernel_ms": k_ms, "kernel_share_of_step": (k_ms * args.steps / elapsed_ms) if k_ms else None,
                     "stages_ms": dict(zip(("prep_scan_sort", "table", "chain"), stages[:3])),
                     "algorithmic_bytes_per_launch": total_bytes, "peak_source": peak_src},
        "clocks": clocks,
        "tokens_per_step": all_tokens, "chars_per_token": all_bytes / max(all_tokens, 1.0), "strings_failed": bad_status,
    }
    if cap and cap.get("warp_inst_per_launch") and k_ms:
        # issue roofline: warp instructions of one launch (ncu, committed capture) against what the
        # 148 SMs x 4 schedulers can issue at the clock sampled during this run
        mhz = clocks.get("sm_mhz") or 1965.0
        floor_ms = cap["warp_inst_per_launch"] / (148 * 4 * mhz * 1e6) * 1e3
        line["issue_roofline"] = {"warp_inst_per_launch": cap["warp_inst_per_launch"], "issue_floor_ms": floor_ms,
                                  "frac": floor_ms / k_ms, "source": cap.get("source")}
    if decode_line:
        dec_bytes = decode_line["ids"] * enc + decode_line["text_bytes"]
        decode_line["roofline"] = {"bound": "hbm", "achieved": dec_bytes / (decode_line["ms"] / 1e3) / 1e9, "peak": peak,
                                   "unit": "GB/s", "frac": dec_bytes / (decode_line["ms"] / 1e3) / 1e9 / peak,
                                   "algorithmic_bytes": dec_bytes, "what": "ids read + text written, whole decode call"}
        line["decode"] = decode_line
    if root_line:
        line["root_scatter_gather"] = root_line
    if not args.no_cpu_baseline:
        cb = cpu_reference_rate(vb, h_np, h_off, budget_s=args.cpu_budget)
        take, chk = cb.pop("strings"), cb.pop("checksum")
        if chk is not None:
            from oracle import oracle
            mine = oracle.ids_checksum(r.raw, enc, r.tok_off, take)
            ids_mine = int(r.tok_off[take])
            line["parity"] = {"strings": take, "ids": ids_mine, "match": bool(mine == chk and ids_mine == cb["ids"]),
                              "checksum": f"{mine:016x}", "reference_checksum": f"{chk:016x}",
                              "against": "oracle/_ref (reference C++ runtime), per-string FNV-1a of the ids, summed"}
        line["cpu_baseline"] = cb
    emit_line(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
