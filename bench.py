#!/usr/bin/env python
"""Benchmark of the ABnet3 hot path on B200 (contract: see the task brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], "C2"): a synthetic Buckeye-shaped corpus
(40-dim fbank x 7-frame stack = 280-dim float32 rows, tokens of 20-80 frames)
and 1 M same-word token pairs PER GPU; one "step" aligns all of them: cosine
frame distance -> DTW -> traceback -> aligned frame-index pairs (the fused
kernel behind abn_align_pairs).  The pair list shards across ranks with no
data-path collective (weak scaling).

One JSON line on stdout (rank 0):
  value   DTW pairs/s, whole job, inputs resident in HBM, CUDA-event timed
  e2e     the same through the host-buffer call a user makes
          (abnet3_b200.utils.align_pairs_host): pinned feature table + pair
          list copied H2D, aligned, paths copied D2H, all inside the timed region
  roofline       algorithmic bytes / kernel time vs the measured HBM peak
  cpu_baseline   the oracle port (numpy cosine distance + C DTW + gather, the
                 reference's CPU path) on this box's host cores, bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "dtw_pairs_per_sec"
UNIT = "pairs/s"
FEAT_DIM = 280


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=1_000_000, help="pairs per GPU per step")
    ap.add_argument("--tokens", type=int, default=40_000, help="tokens in the corpus")
    ap.add_argument("--cpu-sample", type=int, default=0,
                    help="pairs in the CPU baseline sample (0 = 1024 x cores)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the C3 training-step leg")
    ap.add_argument("--train-steps", type=int, default=40)
    ap.add_argument("--train-batch", type=int, default=8192, help="frame pairs per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ----------------------------------------------------------------- CPU arm ---
_W = {}


def _cpu_init(feat, pairs):
    os.environ["OMP_NUM_THREADS"] = "1"
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(1)
    except Exception:
        pass
    _W["feat"], _W["pairs"] = feat, pairs


def _cpu_work(span):
    """The reference's per-pair CPU path (dataloader.py:183-206): slice both
    tokens, get_dtw_alignment (numpy cosine_distance + C DTW), gather rows."""
    import oracle
    feat, pairs = _W["feat"], _W["pairs"]
    n_ok = 0
    for s1, n1, s2, n2 in pairs[span[0]:span[1]].tolist():
        f1, f2 = feat[s1:s1 + n1], feat[s2:s2 + n2]
        try:
            p1, p2 = oracle.get_dtw_alignment(f1, f2)
        except Exception:
            continue
        x1, x2 = f1[p1, :], f2[p2, :]
        n_ok += int(x1.shape[0] == x2.shape[0])
    return n_ok


def cpu_align_rate(feat, pairs, cores):
    """pairs/s of the oracle port over `cores` processes on a bounded sample."""
    import multiprocessing as mp
    import oracle
    oracle.build_dtw()
    n = len(pairs)
    spans = [(i * n // cores, (i + 1) * n // cores) for i in range(cores)]
    ctx = mp.get_context("fork")
    _cpu_init(feat, pairs)
    with ctx.Pool(cores, initializer=_cpu_init, initargs=(feat, pairs)) as pool:
        pool.map(_cpu_work, [(0, min(8, n))] * cores)      # warm the workers
        t0 = time.perf_counter()
        done = sum(pool.map(_cpu_work, spans))
        dt = time.perf_counter() - t0
    return n / dt, dt, done


# ------------------------------------------------------------------ clocks ---
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.thread = [], None, None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                clk, mxv = float(parts[0]), float(parts[1])
            except ValueError:
                continue
            mx = mxv
            if t0 <= ts <= t1 + 0.1:
                sm.append(clk)
                for nm, val in zip(names, parts[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
        if not sm:
            sm = [float(r[1].split(",")[0]) for r in self.rows[-3:] if r[1]]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


def ncu_traffic():
    """dram bytes per PAIR of the align kernels from the committed ncu captures
    (profiles/traffic.json: {"generic": B, "stacked": B}, written when a --set full
    capture is summarised)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            return json.load(fh)
    except Exception:
        return None


# --------------------------------------------------------------- reference ---
def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path (the
    oracle port; the reference itself is Python + an un-vendored Cython DTW and
    cannot travel to this box), all host cores, bounded sample per step."""
    if rank != 0:
        return
    import torch
    from abnet3_b200 import synth
    cores = os.cpu_count() or 1
    corpus = synth.make_corpus(min(args.tokens, 8000), seed=0, device="cpu")
    feat = corpus.feat.numpy()
    # size a step for ~4 s of CPU work (calibrated on a short run), unless told otherwise
    per_step = args.cpu_sample
    if not per_step:
        probe = synth.make_same_pairs(corpus, 512 * cores, seed=2).numpy()
        rate, _, _ = cpu_align_rate(feat, probe, cores)
        per_step = int(max(2048 * cores, min(rate * 4.0, 4_000_000)))
    pairs = synth.make_same_pairs(corpus, per_step, seed=1).numpy()
    for _ in range(args.warmup):
        cpu_align_rate(feat, pairs[:max(cores * 64, 64)], cores)
    t_total, n_total = 0.0, 0
    for _ in range(args.steps):
        rate, dt, _ = cpu_align_rate(feat, pairs, cores)
        t_total += dt
        n_total += len(pairs)
    value = n_total / t_total
    sample = "%d pairs/step x %d steps of the C2 pair distribution, %d processes" % (
        per_step, args.steps, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64",
        "data": "synthetic",
        "config": {"workload": "C2: same-word token pairs, 20-80 frames, 280-dim stacked fbank, "
                               "cosine distance + DTW + traceback (alignment only)",
                   "pairs_per_step": per_step, "note": "CPU sample of the GPU arm's workload"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------- train leg ---
MLP_FLOP_PER_FRAME_PAIR = 7.72e6     # 280-500-500-500-100, fwd + dgrad + wgrad, 2 rows


def bench_train(args, corpus, pairs, res, world, rank, dev):
    """Config C3: batches of `train_batch` frame pairs (aligned same pairs + 1:1
    diff pairs) through gather -> MLP fwd -> coscos2 -> bwd -> [all-reduce] ->
    Adadelta, the canonical buckeye.yaml setup (abnet3/trainer.py:226-256)."""
    import torch
    import torch.distributed as dist
    from abnet3_b200 import ops, synth
    from abnet3_b200.engine import SiameseTrainStep
    from abnet3_b200.model import SiameseNetwork

    # frame-pair table: same pairs from the alignment above + as many diff pairs
    n_sub = min(pairs.shape[0], 200_000)
    sub = ops.AlignResult(res.idx1, res.idx2, res.path_off[:n_sub + 1].contiguous(),
                          res.path_len[:n_sub].contiguous(), res.cost[:n_sub], res.valid[:n_sub])
    s1, s2, _ = ops.compact_paths(sub)
    dpairs = synth.make_diff_pairs(corpus, n_sub, seed=100 + rank)
    d1, d2, _ = ops.diff_pairs(dpairs, stretch=False)
    idx1 = torch.cat([s1, d1]).contiguous()
    idx2 = torch.cat([s2, d2]).contiguous()
    y = torch.cat([torch.ones(s1.numel(), dtype=torch.int8, device=dev),
                   -torch.ones(d1.numel(), dtype=torch.int8, device=dev)]).contiguous()
    n_fp = idx1.numel()
    B = args.train_batch
    perm = torch.randperm(n_fp, device=dev)
    feat = corpus.feat
    buf = torch.empty((2 * B, feat.shape[1]), dtype=torch.float32, device=dev)
    yb = torch.empty(B, dtype=torch.float32, device=dev)
    out = {}
    stream = torch.cuda.current_stream()
    for prec in ("bf16", "fp32"):
        torch.manual_seed(0)
        net = SiameseNetwork(input_dim=feat.shape[1], num_hidden_layers=2, hidden_dim=500,
                             output_dim=100, p_dropout=0.0, activation_layer="sigmoid",
                             precision=prec).to(dev)
        step = SiameseTrainStep(net, ("coscos2", 0.0, False), "adadelta", lr=0.1, momentum=None)
        steps = args.train_steps if prec == "bf16" else max(5, args.train_steps // 4)

        sx, sy = step.input_buffers(B)

        if prec == "bf16":
            # gather -> bf16 operand -> forward -> loss -> backward -> optimizer, graph replayed
            sel = step.gather_buffers(B)

            def one(i):
                lo = (i * B) % max(n_fp - B, 1)
                sel.copy_(perm[lo:lo + B], non_blocking=True)
                return step.step_gather(feat, idx1, idx2, y, B, graph=True)
        else:
            def one(i):
                lo = (i * B) % max(n_fp - B, 1)
                ops.gather_batch(feat, idx1, idx2, y, perm[lo:lo + B], B, out=(sx[:B], sx[B:], sy))
                return step.step(sx, B, sy, graph=True)

        for i in range(5):
            one(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        beg, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        beg.record(stream)
        for i in range(steps):
            loss = one(5 + i)
        end.record(stream)
        torch.cuda.synchronize()
        ms = beg.elapsed_time(end)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        rate = world * B * steps / (ms * 1e-3)
        out[prec] = {"frame_pairs_per_s": rate, "ms_per_step": ms / steps, "steps": steps,
                     "loss": float(loss.item()),
                     "mlp_tflops": rate * MLP_FLOP_PER_FRAME_PAIR / 1e12 / world}
        if prec == "bf16":
            # e2e: the batch's index pairs and labels come from pinned host memory and the
            # loss is read back every step (the reference's `.data[0]`, trainer.py:242)
            h1 = idx1[perm[:B * 40]].cpu().pin_memory()
            h2 = idx2[perm[:B * 40]].cpu().pin_memory()
            hy = y[perm[:B * 40]].cpu().pin_memory()
            e_steps = 40
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e1 = torch.empty(B, dtype=torch.int32, device=dev)
            e2 = torch.empty(B, dtype=torch.int32, device=dev)
            ey = torch.empty(B, dtype=torch.int8, device=dev)
            sel.copy_(torch.arange(B, device=dev))
            for i in range(3):          # new table pointers: two eager steps, then a fresh graph
                step.step_gather(feat, e1.copy_(h1[:B]), e2.copy_(h2[:B]), ey.copy_(hy[:B]), B)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(e_steps):
                sl = slice(i * B, (i + 1) * B)
                e1.copy_(h1[sl], non_blocking=True)
                e2.copy_(h2[sl], non_blocking=True)
                ey.copy_(hy[sl], non_blocking=True)
                lv = float(step.step_gather(feat, e1, e2, ey, B, graph=True).item())
            dt = time.perf_counter() - t0
            tt = torch.tensor([dt], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            out["e2e"] = {"value": world * B * e_steps / float(tt.item()),
                          "unit": "frame pairs/s", "h2d_bytes_per_step": B * 9,
                          "d2h_bytes_per_step": 4, "steps": e_steps, "last_loss": lv}
    peak = 1378.8
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peak = float(json.load(fh)["bf16_tflops_sustained"])
    except Exception:
        pass
    return {"metric": "trained_frame_pairs_per_sec", "unit": "frame pairs/s",
            "config": "C3: SiameseNetwork 280-500-500-500-100 sigmoid, coscos2(avg=False), Adadelta, "
                      "%d frame pairs per GPU per step (aligned same + 1:1 diff)" % B,
            "value": out["bf16"]["frame_pairs_per_s"], "precision": "bf16 tcgen05 (fp32 accumulate)",
            "ms_per_step": out["bf16"]["ms_per_step"],
            "tensor_util": out["bf16"]["mlp_tflops"] / peak, "tensor_peak_tflops": peak,
            "fp32_simt": out["fp32"], "bf16": out["bf16"], "e2e": out.get("e2e"),
            "frame_pairs_in_table": int(n_fp)}


# -------------------------------------------------------------------- ours ---
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from abnet3_b200 import ops, synth, utils

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    corpus = synth.make_corpus(args.tokens, seed=0, device=dev)       # replicated table
    pairs = synth.make_same_pairs(corpus, args.pairs, seed=1 + rank)  # this rank's shard
    feat = corpus.feat
    P = pairs.shape[0]
    max_frames = int(pairs[:, [1, 3]].max().item())
    torch.cuda.synchronize()

    # the product path picks the stacked fast path when the table is a verified 7x40 stack
    # (FeatureTable does this check at load); both kernel families are timed
    last = torch.zeros(feat.shape[0], dtype=torch.uint8, device=dev)
    last[(corpus.file_off[1:] - 1).long()] = 1
    stack = 7 if ops.stack_violations(feat, 7, last) == 0 else 0
    stream = torch.cuda.current_stream()

    def timed(aligner, steps, sample_clocks):
        for _ in range(max(args.warmup, 3)):
            aligner.align(pairs)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        sampler = ClockSampler(local_rank)
        if sample_clocks and rank == 0:
            sampler.start()
            time.sleep(0.2)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
               for _ in range(steps)]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        beg, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        beg.record(stream)
        for a, b in evs:
            a.record(stream)
            r = aligner.align(pairs)
            b.record(stream)
        end.record(stream)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        if world > 1:
            dist.barrier()
        total_ms = beg.elapsed_time(end)
        kern = float(np.mean([a.elapsed_time(b) for a, b in evs]))
        clk = sampler.stop(t0, t1) if (sample_clocks and rank == 0) else None
        tmax = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        return float(tmax.item()), kern, clk, r

    generic = utils.BatchAligner(feat, max_pairs=P, max_frames=max_frames, stack=0)
    g_total_ms, g_kern_ms, g_clocks, res = timed(generic, args.steps, stack == 0)
    if stack:
        fast = utils.BatchAligner(feat, max_pairs=P, max_frames=max_frames, stack=stack)
        total_ms_max, kern_ms, clocks, res_fast = timed(fast, args.steps, True)
        same_bits = bool(torch.equal(res_fast.cost.view(torch.int64), res.cost.view(torch.int64))
                         and torch.equal(res_fast.path_len, res.path_len))
    else:
        total_ms_max, kern_ms, clocks, same_bits = g_total_ms, g_kern_ms, g_clocks, None
    value = world * P * args.steps / (total_ms_max * 1e-3)
    value_generic = world * P * args.steps / (g_total_ms * 1e-3)

    from abnet3_b200 import _lib
    timed_aligner = fast if stack else generic
    launches_per_call = int(_lib.lib().abn_align_launches(P, max_frames, stack, timed_aligner._ws_bytes))

    # algorithmic bytes of one launch (DESIGN.md "Roofline"):
    #   4*dim*(n1+n2) token rows read once + 8*L index pairs written + 16 B/pair
    n12 = (pairs[:, 1].long() + pairs[:, 3].long()).sum().item()
    L_total = int(res.path_len.long().sum().item())
    n_valid = int(res.valid.long().sum().item())
    alg_bytes = 4 * FEAT_DIM * n12 + 8 * L_total + 16 * P
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    achieved_generic = alg_bytes / (g_kern_ms * 1e-3) / 1e9
    traffic_pp = ncu_traffic()

    def traffic_of(key):
        """DRAM bytes of one launch sequence from the committed ncu capture (per pair x pairs)."""
        v = traffic_pp.get(key) if traffic_pp else None
        return v * P if v else None
    flops = 2.0 * FEAT_DIM * (pairs[:, 1].double() * pairs[:, 3].double()).sum().item()

    # ---- e2e: host buffers in, host paths out, every step ----------------
    e2e = None
    if not args.no_e2e:
        host_feat = torch.empty(feat.shape, dtype=feat.dtype, pin_memory=True)
        host_feat.copy_(feat)
        host_pairs = torch.empty(pairs.shape, dtype=pairs.dtype, pin_memory=True)
        host_pairs.copy_(pairs)
        torch.cuda.synchronize()
        e_steps = max(2, min(args.steps, 3))
        # a verified stack uploads its 40-wide middle blocks only (+ the file-edge flags)
        host_last = None
        if stack:
            host_last = torch.empty(last.shape, dtype=last.dtype, pin_memory=True)
            host_last.copy_(last)
        hres = utils.align_pairs_host(host_feat, host_pairs, max_frames=max_frames, stack=stack,
                                      last_row_of_file=host_last)
        h2d = (host_feat.numel() * 4 // (stack if stack else 1) + host_pairs.numel() * 4 +
               (host_last.numel() if stack else 0))
        d2h = sum(t.numel() * t.element_size() for t in hres)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        tt0 = time.perf_counter()
        for _ in range(e_steps):
            hres = utils.align_pairs_host(host_feat, host_pairs, max_frames=max_frames, stack=stack,
                                          last_row_of_file=host_last)
        torch.cuda.synchronize()
        dt = time.perf_counter() - tt0
        tm = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        e2e = {"value": world * P * e_steps / float(tm.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "steps": e_steps,
               "call": "abnet3_b200.utils.align_pairs_host(feat_host, pair_tok_host, stack=%d%s)"
                       % (stack, ", last_row_of_file" if stack else "")}
        del host_feat

    # ---- C3 leg: siamese training steps on the aligned frame pairs ---------
    train = None
    if not args.no_train:
        train = bench_train(args, corpus, pairs, res, world, rank, dev)

    # ---- CPU baseline (rank 0, N == 1 only) --------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        feat_h = feat.cpu().numpy()
        pairs_h = pairs.cpu().numpy()
        n_s = args.cpu_sample
        if not n_s:         # ~15 s of CPU work, calibrated on a short run
            rate0, _, _ = cpu_align_rate(feat_h, pairs_h[:min(P, 1024 * cores)], cores)
            n_s = int(max(8192 * cores, rate0 * 15.0))
        n_s = min(n_s, 3 * P)
        sp = np.concatenate([pairs_h] * 3)[:n_s] if n_s > P else pairs_h[:n_s]
        rate, dt, _ = cpu_align_rate(feat_h, sp, cores)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d pairs of the step's pair list (cycled), %.1f s on %d processes "
                         "(numpy cosine_distance + C DTW oracle + row gather per pair)"
                         % (n_s, dt, cores)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 distance, f64 DTW accumulate", "data": "synthetic",
            "config": {
                "workload": "C2: 1M same-word token pairs per GPU, tokens 20-80 frames, 280-dim "
                            "(40 fbank x 7 stack) float32, cosine distance + DTW + traceback",
                "pairs_per_gpu_per_step": P, "corpus_frames": int(feat.shape[0]),
                "corpus_bytes": int(feat.numel() * 4),
                "l2": "inputs larger than L2 (feature table %.1f GB, every step re-reads it)"
                      % (feat.numel() * 4 / 1e9),
                "valid_pairs": n_valid, "mean_path_len": L_total / max(n_valid, 1),
                "frame_pairs_per_s": value * L_total / P,
            },
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "peak_source": peak_src,
                         "kernel": ("align_stack_kernel<RA,NCG> (one launch per size class; stacked "
                                    "fast path, bit-identical to the generic kernels)" if stack else
                                    "align_class_kernel<RA,NCG> (one launch per size class)"),
                         "kernel_ms": kern_ms,
                         "algorithmic_bytes_per_launch": int(alg_bytes),
                         "algorithmic_bytes_def": "4*280*(n1+n2) + 8*L + 16 per pair: the stacked rows "
                                                  "the API is handed, read once (SURVEY 8d)",
                         "traffic": traffic_of("stacked" if stack else "generic"),
                         "note": ("the stacked path reads each 40-wide frame once (about 1/6 of the "
                                  "algorithmic bytes), so frac measures work done per second against "
                                  "the stacked-bytes roofline, not DRAM traffic" if stack else None),
                         "generic_kernels": {
                             "value": value_generic, "kernel_ms": g_kern_ms,
                             "achieved": achieved_generic, "frac": achieved_generic / peak,
                             "fp32_tflops": flops / (g_kern_ms * 1e-3) / 1e12,
                             "traffic": traffic_of("generic"),
                             "same_bits_as_fast_path": same_bits}},
            "e2e": e2e,
            "train": train,
            "cpu_baseline": cpu,
            "gpu_launches": args.steps * launches_per_call,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under it
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
               "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
               "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
