#!/usr/bin/env python
"""bench.py -- MB/s tokenized (input bytes) on the BASELINE.json configuration.

    python bench.py --gpus N --steps K --warmup W            (this repo's CUDA path)
    python bench.py --impl reference --gpus N --steps K ...  (the reference's own C++ CPU runtime)
    torchrun --nproc-per-node N bench.py --gpus N ...        (one rank per GPU, weak scaling)

A "step" = one pass of the hot path (tokenize kernel + offset scan + id packing) over one batch of
synthetic strings.  N=1 workload = BASELINE.json configs[1]: the englishcode-32000 stand-in,
1,000,000 strings ~U[64,960] bytes (mean 512), resident in HBM before the timed region
(512 MB of input per step >> 126 MB L2, so no L2 flush is needed between steps).
`value`  : whole-job MB/s (10^6 B) with inputs already in HBM, device-timed, max over ranks.
`e2e`    : the same metric through the C-ABI host-buffer call tmgpu_tokenize_batch with pinned HOST
           buffers -- H2D of the strings and D2H of the ids inside the timed region.
`roofline`: the tokenize kernel against the measured HBM copy bandwidth; algorithmic bytes = 1 B
           read per input byte.
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--strings", type=int, default=0, help="strings per GPU (default: the config's 1,000,000)")
    ap.add_argument("--config", default=WORKLOAD)
    ap.add_argument("--e2e-steps", type=int, default=3, help="timed steps of the host-buffer (e2e) leg, >= 1")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--check", action="store_true", help="verify a sample of the GPU ids against the oracle")
    args = ap.parse_args()
    args.e2e_steps = max(1, args.e2e_steps)
    return args


def config_block(args, n_strings, total_bytes, world):
    from workloads import configs
    cfg = configs.CONFIGS[args.config]
    return {
        "workload": f"{cfg['vocab']} (stand-in for BASELINE configs[1] englishcode-32000-consistent-v1), "
                    f"{n_strings} strings/GPU ~U[{cfg['lo']},{cfg['hi']}] B, ASCII prose+code",
        "strings_per_gpu": n_strings,
        "bytes_per_gpu": int(total_bytes),
        "vocab_ids": cfg["ids"],
        "capcode": cfg["capcode"],
        "enc_len": 2,
        "parallelism": f"strings sharded over {world} GPU(s), no data-path collective",
        "l2_policy": "inputs larger than L2 (512 MB/step vs 126 MB), no flush",
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


def ncu_traffic():
    """dram bytes per launch of the tokenize kernel from the committed ncu --set full capture."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(p) as f:
            return json.load(f)
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
def cpu_reference_rate(vocab_bytes, data_np, off_np, budget_s=12.0, threads=None):
    """Time oracle/_ref (reference C++ runtime) on `threads` host cores on a bounded prefix."""
    import numpy as np
    from oracle import oracle
    kind = "reference" if oracle.have_reference() else "port"
    threads = threads or os.cpu_count() or 1
    n = len(off_np) - 1
    if kind == "reference":
        v = oracle.OracleVocab(vocab_bytes, "reference")
        probe = min(n, 2000)
        t = time.perf_counter()
        v.tokenize_batch_timed(data_np, off_np[:probe + 1], threads)
        dt = max(time.perf_counter() - t, 1e-4)
        rate = float(off_np[probe] - off_np[0]) / dt
        take = int(min(n, max(probe, rate * budget_s / (float(off_np[-1] - off_np[0]) / n))))
        t = time.perf_counter()
        ids, chk = v.tokenize_batch_timed(data_np, off_np[:take + 1], threads)
        dt = time.perf_counter() - t
        nbytes = float(off_np[take] - off_np[0])
        t1 = time.perf_counter()
        one = min(take, max(200, take // (4 * threads)))
        v.tokenize_batch_timed(data_np, off_np[:one + 1], 1)
        dt1 = time.perf_counter() - t1
        return {"value": nbytes / dt / 1e6, "unit": UNIT, "cores": threads, "kind": kind,
                "sample": f"first {take} strings ({nbytes / 1e6:.1f} MB) of the same batch, {dt:.1f} s wall, one std::thread "
                          f"per contiguous slice (training/tokenmonsterserver.go:364-377 analogue)",
                "single_thread_value": float(off_np[one] - off_np[0]) / dt1 / 1e6, "ids": int(ids)}
    v = oracle.OracleVocab(vocab_bytes, "port")
    take, nbytes = 0, 0
    t = time.perf_counter()
    while take < n and time.perf_counter() - t < budget_s:
        v.tokenize(data_np[int(off_np[take]):int(off_np[take + 1])].tobytes())
        nbytes += int(off_np[take + 1] - off_np[take])
        take += 1
    dt = time.perf_counter() - t
    return {"value": nbytes / dt / 1e6, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"first {take} strings ({nbytes / 1e6:.1f} MB), scalar C port, 1 thread"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np
    from workloads import configs
    cfg = configs.CONFIGS[args.config]
    n = args.strings or cfg["n"]
    vb = configs.vocab_bytes(args.config)
    sample_n = min(n, 600000)
    data, off = configs.batch(args.config, sample_n, "cpu")
    data_np, off_np = data.numpy(), off.numpy().astype(np.uint64)
    threads = os.cpu_count() or 1
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
        "config": config_block(args, n, float(off_np[-1]) / sample_n * n, max(1, args.gpus)),   # the GPU arm's config
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": last["cores"], "kind": last["kind"],
                         "sample": last["sample"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
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
    n = args.strings or cfg["n"]
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
    enc = 2

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
    res = ws.results()
    n_tokens = res["n_tokens"]
    bad_status = int((res["status"] != 0).sum().item())

    # ---- end to end through the C-ABI host call (pinned host buffers) --------------------------
    h_data = torch.empty(total_bytes, dtype=torch.uint8, pin_memory=True)
    h_data.copy_(data)
    h_off = off.cpu().numpy().astype(np.uint64)
    h_np = h_data.numpy()
    r = vocab.tokenize_batch(h_np, h_off, enc)  # warm-up (allocations, pinned pools)
    d2h = r.raw.nbytes + r.tok_off.nbytes + r.missing.nbytes + r.status.nbytes
    h2d = total_bytes + h_off.nbytes
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        r = None  # hand the previous result's pinned buffers back to the pool first
        r = vocab.tokenize_batch(h_np, h_off, enc)
    torch.cuda.synchronize(dev)
    e2e_s = (time.perf_counter() - t0) / args.e2e_steps
    stop.set()
    th.join(timeout=3)

    # ---- N > 1: the single-root path (one NCCL scatter of bytes, one gather of ids) -----------
    root_line = None
    if world > 1:
        from tokenmonster_b200.dist import gpu_compute, tokenize_from_root
        comp = gpu_compute(ws, enc)
        for _ in range(2):
            out = tokenize_from_root(data if rank == 0 else None, off if rank == 0 else None, comp, 0, dev)
        sync_all()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        out = tokenize_from_root(data if rank == 0 else None, off if rank == 0 else None, comp, 0, dev)
        r1.record()
        sync_all()
        tr = torch.tensor([r0.elapsed_time(r1)], dtype=torch.float64, device=dev)
        dist.all_reduce(tr, op=dist.ReduceOp.MAX)
        if rank == 0:
            same = bool(torch.equal(out[0], res["tokens_u8"]))   # identical bytes to the 1-GPU result
            root_line = {"value": total_bytes / (float(tr[0]) / 1e3) / 1e6, "unit": UNIT, "ms": float(tr[0]),
                         "what": "rank 0's batch scattered over all ranks by bytes, tokenized, ids gathered back "
                                 "(NCCL grouped send/recv); strong scaling of one batch", "identical_to_1gpu": same}

    # ---- reduce over ranks --------------------------------------------------------------------
    t = torch.tensor([elapsed_ms, e2e_s, float(total_bytes), float(n_tokens)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = t.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        elapsed_ms, e2e_s = float(mx[0]), float(mx[1])
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
    traffic = ncu_traffic()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": config_block(args, n, total_bytes, world),
        "e2e": {"value": all_bytes / e2e_s / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "api": "tmgpu_tokenize_batch (C ABI, pinned host buffers)", "steps": args.e2e_steps},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None,
                     "traffic": traffic.get("dram_bytes_per_launch") if traffic else None,
                     "kernel": "tokenize_thread_kernel<1024, capcode, store>", "kernel_ms": k_ms, "kernel_share_of_step": (k_ms * args.steps / elapsed_ms) if k_ms else None,
                     "algorithmic_bytes_per_launch": total_bytes, "peak_source": peak_src},
        "clocks": summarize_clocks(samples),
        "tokens_per_step": all_tokens, "chars_per_token": all_bytes / max(all_tokens, 1.0), "strings_failed": bad_status,
        "kernel_mapping": {"1": "lane-per-string (T2)"}.get(os.environ.get("TMGPU_GROUP", "1"), "lanes-over-branches G=" + os.environ.get("TMGPU_GROUP", "1")),
    }
    if root_line:
        line["root_scatter_gather"] = root_line
    if args.check:
        from oracle import oracle
        chk = oracle.OracleVocab(vb, "port")
        ok = 0
        for s in range(0, min(n, 20000), 97):
            want, _ = chk.tokenize(h_np[int(h_off[s]):int(h_off[s + 1])].tobytes())
            assert np.array_equal(r[s].astype(np.uint32), want), f"string {s} differs from the oracle"
            ok += 1
        line["checked_strings"] = ok
    if not args.no_cpu_baseline:
        sample_n = min(n, 600000)
        cb = cpu_reference_rate(vb, h_np, h_off[:sample_n + 1], budget_s=12.0)
        line["cpu_baseline"] = cb
    emit_line(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
