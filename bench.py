#!/usr/bin/env python
"""Benchmark of the ABnet3 hot path on B200 (contract: see the task brief).

    python bench.py [--config C3|C2|C4|C5] [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline (default, BASELINE.json configs[2] "C3", the configuration the metric
"aligned+trained frame pairs/sec" is quoted on): a synthetic Buckeye-shaped
corpus (40-dim fbank x 7-frame stack = 280-dim float32 rows, tokens of 20-80
frames), 1 M same-word + 1 M different-word token pairs PER GPU.  One "step" is
the whole path through the reference-facing surface:

    FramesDataLoader.realign()          cosine distance -> DTW -> traceback for every
                                        same pair, diff pairs, global shuffle
    TrainerSiamese.optimize_model()     one epoch of 8192-frame-pair batches
                                        (gather -> MLP fwd -> coscos2 -> bwd -> gradient
                                        exchange -> Adadelta) + the dev sweep

    value = trained frame pairs / (t_align + t_epoch), whole job (weak scaling: every
    rank has its own shard of the pair list; the gradient exchange is inside the step).

One JSON line on stdout (rank 0).  Beside the headline it carries
  align          DTW pairs/s of the alignment stage alone (config C2) + its HBM roofline
  train_step     us per training step, tensor utilisation
  roofline       the dominant kernel of the timed region (forward chain, tensor bound),
                 timed alone with CUDA events
  e2e            the same metric from HOST buffers (feature table + pair lists copied
                 H2D and the epoch's losses read back inside the timed region)
  cpu_baseline   the oracle port of the same path on this box's host cores (bounded sample)
  dp_check       (N > 1) ranks bit-identical after K steps, and the rel. distance to NCCL

--config C2: alignment only (dtw_pairs_per_sec).  --config C4: SiameseMultitaskNetwork.
--config C5: DTW length sweep 50-400 frames, 4 M pairs in all, strong scaling.
"""
import argparse
import contextlib
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

FEAT_DIM = 280
METRICS = {
    "C3": ("aligned_trained_frame_pairs_per_sec", "frame pairs/s"),
    "C4": ("aligned_trained_frame_pairs_per_sec", "frame pairs/s"),
    "C2": ("dtw_pairs_per_sec", "pairs/s"),
    "C5": ("dtw_pairs_per_sec", "pairs/s"),
}
WORKLOADS = {
    "C3": "C3: full siamese training epoch on 1M aligned same-word pairs + 1:1 different pairs per GPU "
          "(tokens 20-80 frames, 280-dim stacked fbank), SiameseNetwork 280-500-500-500-100 sigmoid, "
          "coscos2(avg=False), Adadelta lr 0.1, batches of 8192 frame pairs; alignment (cosine "
          "distance + DTW + traceback) inside every step",
    "C4": "C4: as C3 with SiameseMultitaskNetwork 280-500-(500)x2-{100 spk, 100 phn} sigmoid, "
          "weighted_loss_multi(coscos2, coscos2, 0.5), speaker labels per pair",
    "C2": "C2: 1M same-word token pairs per GPU, tokens 20-80 frames, 280-dim (40 fbank x 7 stack) "
          "float32, cosine distance + DTW + traceback (alignment only)",
    "C5": "C5: DTW length sweep, tokens of 50 / 100 / 200 / 400 frames, 1M pairs per length (4M in "
          "all) sharded over the GPUs, cosine distance + DTW + traceback",
}
# MLP FLOPs per frame pair (2 rows) per training step, unpadded shapes
MLP_FLOP = {"C3": 7.72e6, "C4": 2 * 2 * (3 * (280 * 500 + 2 * 500 * 500 + 500 * 200) - 280 * 500)}
MLP_FLOP_FWD = {"C3": 2.76e6, "C4": 2 * 2 * (280 * 500 + 2 * 500 * 500 + 500 * 200)}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3", choices=["C2", "C3", "C4", "C5"])
    ap.add_argument("--pairs", type=int, default=1_000_000, help="same pairs per GPU per step")
    ap.add_argument("--tokens", type=int, default=40_000, help="tokens in the corpus")
    ap.add_argument("--train-batch", type=int, default=8192, help="frame pairs per GPU per step")
    ap.add_argument("--epoch-batches", type=int, default=0,
                    help="cap the batches per epoch (0 = the whole table, as the config says)")
    ap.add_argument("--align-steps", type=int, default=5, help="timed passes of the alignment leg")
    ap.add_argument("--cpu-sample", type=int, default=0,
                    help="pairs in the CPU baseline sample (0 = calibrated to ~15 s)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-align-leg", action="store_true")
    ap.add_argument("--no-kernel-times", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dp-check", action="store_true")
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
    """The reference's per-pair CPU path (dataloader.py:183-206, :642-653): slice both
    tokens, get_dtw_alignment (numpy cosine_distance + C DTW), gather rows."""
    import oracle
    feat, pairs = _W["feat"], _W["pairs"]
    n_ok, n_fp = 0, 0
    for s1, n1, s2, n2 in pairs[span[0]:span[1]].tolist():
        f1, f2 = feat[s1:s1 + n1], feat[s2:s2 + n2]
        try:
            p1, p2 = oracle.get_dtw_alignment(f1, f2)
        except Exception:
            continue
        x1, x2 = f1[p1, :], f2[p2, :]
        n_ok += int(x1.shape[0] == x2.shape[0])
        n_fp += x1.shape[0]
    return n_ok, n_fp


def cpu_align_rate(feat, pairs, cores):
    """(pairs/s, seconds, valid pairs, frame pairs) of the oracle port over `cores` processes."""
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
        res = pool.map(_cpu_work, spans)
        dt = time.perf_counter() - t0
    return n / dt, dt, sum(r[0] for r in res), sum(r[1] for r in res)


class CpuTrainer(object):
    """The reference's training step on the host cores: abnet3/model.py + loss.py restated by
    oracle/nets.py (pinned to the live reference), torch autograd, torch.optim.Adadelta --
    abnet3/trainer.py:231-243 for one batch."""

    def __init__(self, multitask, cores, seed=0):
        import torch
        torch.set_num_threads(cores)
        g = torch.Generator().manual_seed(seed)
        shapes = [(500, 280), (500, 500), (500, 500)]
        self.multitask = multitask
        names = ["input_emb.0", "hidden_layers%s.0" % ("_shared" if multitask else ""),
                 "hidden_layers%s.3" % ("_shared" if multitask else "")]
        heads = ["output_layer_spk.0", "output_layer_phn.0"] if multitask else ["output_layer.0"]
        self.sd = {}
        for nm, (o, i) in list(zip(names, shapes)) + [(h, (100, 500)) for h in heads]:
            bound = float(np.sqrt(6.0 / (o + i)))
            self.sd[nm + ".weight"] = ((torch.rand(o, i, generator=g) * 2 - 1) * bound).requires_grad_()
            self.sd[nm + ".bias"] = torch.zeros(o, requires_grad=True)
        self.opt = torch.optim.Adadelta(list(self.sd.values()), lr=0.1)

    def step(self, x1, x2, y, y_spk=None):
        import torch
        from oracle import nets
        n = x1.shape[0]
        x = torch.cat([x1, x2])
        if self.multitask:
            spk, phn = nets.multitask_forward_once(self.sd, x)
            loss = 0.5 * nets.coscos2(spk[:n], spk[n:], y_spk, avg=False) + \
                0.5 * nets.coscos2(phn[:n], phn[n:], y, avg=False)
        else:
            e = nets.siamese_forward_once(self.sd, x)
            loss = nets.coscos2(e[:n], e[n:], y, avg=False)
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        return float(loss.detach())


def cpu_path_rate(feat, same, diff, cores, batch, multitask=False, max_batches=None):
    """aligned+trained frame pairs/s of the CPU port on a sample: align `same` on all cores,
    build the frame-pair table (+ truncated diff pairs), shuffle, train batches of `batch`."""
    import torch
    _, t_align, _, _ = cpu_align_rate(feat, same, cores)
    # the table itself (single process, not timed twice: alignment was timed above)
    import oracle
    oracle.build_dtw()
    rows1, rows2, ys = [], [], []
    for s1, n1, s2, n2 in same.tolist():
        try:
            p1, p2 = oracle.get_dtw_alignment(feat[s1:s1 + n1], feat[s2:s2 + n2])
        except Exception:
            continue
        rows1.append(s1 + np.asarray(p1))
        rows2.append(s2 + np.asarray(p2))
        ys.append(np.ones(len(p1), np.float32))
    for s1, n1, s2, n2 in diff.tolist():
        m = min(n1, n2)
        rows1.append(s1 + np.arange(m))
        rows2.append(s2 + np.arange(m))
        ys.append(-np.ones(m, np.float32))
    r1, r2, y = np.concatenate(rows1), np.concatenate(rows2), np.concatenate(ys)
    perm = np.random.default_rng(0).permutation(len(y))
    r1, r2, y = r1[perm], r2[perm], y[perm]
    n_b = len(y) // batch
    if max_batches:
        n_b = min(n_b, max_batches)
    tr = CpuTrainer(multitask, cores)
    ft = torch.from_numpy(feat)
    yt = torch.from_numpy(y)
    ysp = torch.from_numpy(np.where(np.arange(len(y)) % 2 == 0, 1.0, -1.0).astype(np.float32))
    t0 = time.perf_counter()
    for b in range(n_b):
        sl = slice(b * batch, (b + 1) * batch)
        tr.step(ft[r1[sl]], ft[r2[sl]], yt[sl], ysp[sl] if multitask else None)
    t_train = time.perf_counter() - t0
    # the sample trains n_b batches out of len(y)/batch: scale the alignment time to the
    # same share of the table so that both stages cover the same frame pairs
    share = n_b * batch / max(len(y), 1)
    fp = n_b * batch
    return fp / (t_align * share + t_train), t_align * share, t_train, fp


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
        # nvidia-smi takes up to a second to come up, and its start-up (NVML initialisation)
        # stalls kernel submission: the timed region must not begin before the first sample
        t_end = time.perf_counter() + 8.0
        while not self.rows and time.perf_counter() < t_end and self.proc.poll() is None:
            time.sleep(0.02)

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


def peaks():
    out = {"hbm_gbs": 6650.0, "bf16_tflops": 1600.0, "bf16_tflops_sustained": 1600.0,
           "source": "B200_PROFILING.md fallback"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            d = json.load(fh)
        out.update({k: float(d[k]) for k in ("hbm_gbs", "bf16_tflops", "bf16_tflops_sustained")
                    if k in d})
        out["source"] = "MEASURED_PEAKS.json"
    except Exception:
        pass
    return out


def ncu_traffic():
    """DRAM bytes per unit of the dominant kernels from the committed ncu --set full captures
    (profiles/traffic.json, written when a capture is summarised)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            return json.load(fh)
    except Exception:
        return {}


def max_over_ranks(x, dev, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x, dev, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


# --------------------------------------------------------------- reference ---
def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path (the oracle port; the
    reference itself is Python + an un-vendored Cython DTW and cannot travel to this box), all
    host cores, a bounded sample of the GPU arm's workload per step."""
    if rank != 0:
        return
    from abnet3_b200 import synth
    cfg = args.config
    metric, unit = METRICS[cfg]
    cores = os.cpu_count() or 1
    lens = (20, 80)
    corpus = synth.make_corpus(min(args.tokens, 8000), seed=0, device="cpu", len_range=lens)
    feat = corpus.feat.numpy()
    if cfg in ("C2", "C5"):
        if cfg == "C5":
            corpus = synth.make_corpus(2000, seed=0, device="cpu", len_range=(200, 200))
            feat = corpus.feat.numpy()
        per_step = args.cpu_sample
        if not per_step:
            probe = synth.make_same_pairs(corpus, 64 * cores, seed=2).numpy()
            rate, _, _, _ = cpu_align_rate(feat, probe, cores)
            per_step = int(max(256 * cores, min(rate * 4.0, 4_000_000)))
        pairs = synth.make_same_pairs(corpus, per_step, seed=1).numpy()
        for _ in range(args.warmup):
            cpu_align_rate(feat, pairs[:max(cores * 16, 64)], cores)
        t_total, n_total = 0.0, 0
        for _ in range(args.steps):
            _, dt, _, _ = cpu_align_rate(feat, pairs, cores)
            t_total += dt
            n_total += len(pairs)
        value = n_total / t_total
        sample = "%d pairs/step x %d steps of the %s pair distribution%s, %d processes" % (
            per_step, args.steps, cfg, " at 200 frames/token" if cfg == "C5" else "", cores)
    else:
        per_step = args.cpu_sample or 1500
        same = synth.make_same_pairs(corpus, per_step, seed=1).numpy()
        diff = synth.make_diff_pairs(corpus, per_step, seed=2).numpy()
        for _ in range(min(args.warmup, 1)):
            cpu_path_rate(feat, same[:200], diff[:200], cores, args.train_batch, cfg == "C4", 2)
        t_total, n_total = 0.0, 0
        for _ in range(args.steps):
            _, ta, tt, fp = cpu_path_rate(feat, same, diff, cores, args.train_batch, cfg == "C4")
            t_total += ta + tt
            n_total += fp
        value = n_total / t_total
        sample = ("%d same + %d diff pairs/step x %d steps: oracle alignment on %d processes, then the "
                  "reference MLP + coscos2 + Adadelta (torch CPU, %d threads) over the resulting "
                  "frame pairs in batches of %d" % (per_step, per_step, args.steps, cores, cores,
                                                    args.train_batch))
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True, "scaling": "strong" if cfg == "C5" else "weak", "vs_baseline": None,
        "dtype": "f32 (f64 DTW accumulate)", "data": "synthetic",
        "config": {"workload": WORKLOADS[cfg], "pairs_per_step": per_step,
                   "note": "CPU sample of the GPU arm's workload"},
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------ alignment leg ---
def align_leg(args, corpus, pairs, world, rank, local_rank, dev, steps, sample_clocks=True,
              with_e2e=True):
    """Config C2: device-resident alignment of the pair list, both kernel families, plus the
    host-buffer call.  Returns (record, AlignResult of the last pass)."""
    import torch
    import torch.distributed as dist
    from abnet3_b200 import ops, utils, _lib

    feat = corpus.feat
    P = pairs.shape[0]
    max_frames = int(pairs[:, [1, 3]].max().item())
    last = torch.zeros(feat.shape[0], dtype=torch.uint8, device=dev)
    last[(corpus.file_off[1:] - 1).long()] = 1
    stack = 7 if ops.stack_violations(feat, 7, last) == 0 else 0
    stream = torch.cuda.current_stream()

    def timed(aligner, steps, clocks_on):
        for _ in range(max(args.warmup, 3)):
            aligner.align(pairs)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        sampler = ClockSampler(local_rank)
        if clocks_on and rank == 0:
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
        total_ms = max_over_ranks(beg.elapsed_time(end), dev, world)
        kern = float(np.mean([a.elapsed_time(b) for a, b in evs]))
        clk = sampler.stop(t0, t1) if (clocks_on and rank == 0) else None
        return total_ms, kern, clk, r

    generic = utils.BatchAligner(feat, max_pairs=P, max_frames=max_frames, stack=0)
    g_total_ms, g_kern_ms, g_clocks, res = timed(generic, steps, sample_clocks and stack == 0)
    if stack:
        fast = utils.BatchAligner(feat, max_pairs=P, max_frames=max_frames, stack=stack)
        total_ms, kern_ms, clocks, res_fast = timed(fast, steps, sample_clocks)
        same_bits = bool(torch.equal(res_fast.cost.view(torch.int64), res.cost.view(torch.int64))
                         and torch.equal(res_fast.path_len, res.path_len))
    else:
        total_ms, kern_ms, clocks, same_bits = g_total_ms, g_kern_ms, g_clocks, None
    value = world * P * steps / (total_ms * 1e-3)
    value_generic = world * P * steps / (g_total_ms * 1e-3)
    timed_aligner = fast if stack else generic
    launches = int(_lib.lib().abn_align_launches(P, max_frames, stack, timed_aligner._ws_bytes))

    # algorithmic bytes of one launch sequence (DESIGN.md "Roofline"):
    #   4*dim*(n1+n2) token rows read once + 8*L index pairs written + 16 B/pair
    n12 = (pairs[:, 1].long() + pairs[:, 3].long()).sum().item()
    L_total = int(res.path_len.long().sum().item())
    n_valid = int(res.valid.long().sum().item())
    alg_bytes = 4 * FEAT_DIM * n12 + 8 * L_total + 16 * P
    pk = peaks()
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    achieved_generic = alg_bytes / (g_kern_ms * 1e-3) / 1e9
    tr = ncu_traffic()
    flops = 2.0 * FEAT_DIM * (pairs[:, 1].double() * pairs[:, 3].double()).sum().item()

    def traffic_of(key):
        v = tr.get(key)
        return v * P if v else None

    rec = {
        "metric": "dtw_pairs_per_sec", "value": value, "unit": "pairs/s",
        "ms_per_pass": total_ms / steps, "passes": steps, "pairs_per_gpu": P,
        "valid_pairs": n_valid, "mean_path_len": L_total / max(n_valid, 1),
        "aligned_frame_pairs_per_s": value * L_total / P,
        "launches_per_pass": launches,
        "roofline": {
            "bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
            "frac": achieved / pk["hbm_gbs"], "peak_source": pk["source"] + " hbm_gbs",
            "kernel": ("align_stack_kernel<RA,NCG> + dtw_skew_kernel<G> (one launch per size class; "
                       "stacked fast path, bit-identical to the generic kernels)" if stack else
                       "align_class_kernel<RA,NCG> + dtw_skew_kernel<G> (one launch per size class)"),
            "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": int(alg_bytes),
            "algorithmic_bytes_def": "4*280*(n1+n2) + 8*L + 16 per pair: the stacked rows the API is "
                                     "handed, read once (SURVEY 8d)",
            "traffic": traffic_of("stacked" if stack else "generic"),
            "note": ("the stacked path reads each 40-wide frame once (about 1/6 of the algorithmic "
                     "bytes): frac is work per second against the stacked-bytes roofline, not DRAM "
                     "utilisation" if stack else None),
            "generic_kernels": {
                "value": value_generic, "kernel_ms": g_kern_ms, "achieved": achieved_generic,
                "frac": achieved_generic / pk["hbm_gbs"],
                "fp32_tflops": flops / (g_kern_ms * 1e-3) / 1e12,
                "traffic": traffic_of("generic"), "same_bits_as_fast_path": same_bits}},
        "clocks": clocks,
    }

    # ---- e2e: host buffers in, host paths out, every pass ----------------
    if with_e2e and not args.no_e2e:
        host_feat = torch.empty(feat.shape, dtype=feat.dtype, pin_memory=True)
        host_feat.copy_(feat)
        host_pairs = torch.empty(pairs.shape, dtype=pairs.dtype, pin_memory=True)
        host_pairs.copy_(pairs)
        torch.cuda.synchronize()
        e_steps = 2

        def run_e2e(**kw):
            hres = utils.align_pairs_host(host_feat, host_pairs, max_frames=max_frames, **kw)
            d2h = sum(t.numel() * t.element_size() for t in hres)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            tt0 = time.perf_counter()
            for _ in range(e_steps):
                hres = utils.align_pairs_host(host_feat, host_pairs, max_frames=max_frames, **kw)
            torch.cuda.synchronize()
            dt = max_over_ranks(time.perf_counter() - tt0, dev, world)
            return world * P * e_steps / dt, int(d2h)

        host_last0 = torch.empty(last.shape, dtype=last.dtype, pin_memory=True)
        host_last0.copy_(last)
        v_full, d2h = run_e2e(stack=0, last_row_of_file=host_last0)
        rec["e2e"] = {"value": v_full, "unit": "pairs/s",
                      "h2d_bytes_per_step": int(host_feat.numel() * 4 + host_pairs.numel() * 4),
                      "d2h_bytes_per_step": d2h, "steps": e_steps,
                      "call": "abnet3_b200.utils.align_pairs_host(feat_host [N,280], pair_tok_host, "
                              "last_row_of_file): the whole table it is handed is uploaded, every pass "
                              "(its stack structure is checked on the device); paths come back as two "
                              "int32 per step"}
        if stack:
            # first-class un-stacked input: the [N, 40] frames + file-edge flags are all there is
            # to upload (nothing vouched for); paths come back as 2-bit directions
            f = FEAT_DIM // 7
            host_frames = torch.empty((feat.shape[0], f), dtype=feat.dtype, pin_memory=True)
            host_frames.copy_(feat[:, 3 * f:4 * f])
            host_last = torch.empty(last.shape, dtype=last.dtype, pin_memory=True)
            host_last.copy_(last)
            host_feat = None
            torch.cuda.synchronize()

            def run_e2e_frames():
                kw = dict(max_frames=max_frames, frames_host=host_frames, last_row_of_file=host_last,
                          paths="directions")
                hres = utils.align_pairs_host(None, host_pairs, **kw)
                d2h_b = sum(t.numel() * t.element_size() for t in hres)
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                tt0 = time.perf_counter()
                for _ in range(e_steps):
                    hres = utils.align_pairs_host(None, host_pairs, **kw)
                torch.cuda.synchronize()
                dt = max_over_ranks(time.perf_counter() - tt0, dev, world)
                return world * P * e_steps / dt, int(d2h_b)

            v_un, d2h_un = run_e2e_frames()
            rec["e2e_unstacked"] = {
                "value": v_un, "unit": "pairs/s",
                "h2d_bytes_per_step": int(host_frames.numel() * 4 + host_last.numel() + host_pairs.numel() * 4),
                "d2h_bytes_per_step": d2h_un, "steps": e_steps,
                "call": "abnet3_b200.utils.align_pairs_host(frames_host=[N,40] frames, last_row_of_file, "
                        "pair_tok_host, paths='directions'): un-stacked input stacked on the device, paths "
                        "back as 2-bit step directions (utils.decode_directions restores the indices)"}
        del host_feat
    return rec, res


# ------------------------------------------------------------- kernel times ---
def kernel_times(engine, feat, table, B, iters=30):
    """Each kernel of the training step timed ALONE (eager launches, CUDA events on the
    launching stream): the per-kernel rooflines beside the graph-replayed step time."""
    import torch
    from abnet3_b200 import ops
    stream = torch.cuda.current_stream()
    sel = engine.gather_buffers(B)
    sel.copy_(torch.arange(B, device=feat.device))
    engine._table_step(feat, table, B, sel, True, graph=False)        # buffers hold a real batch
    torch.cuda.synchronize()

    def t(fn):
        """us per launch: `iters` back-to-back launches captured in one CUDA graph (no host
        launch latency between them), replayed between two events."""
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(iters):
                fn()
        g.replay()
        torch.cuda.synchronize()
        stream = torch.cuda.current_stream()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        g.replay()
        b.record(stream)
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters * 1e3

    out = {}
    ys = table[2:]
    out["gather"] = t(lambda: ops.gather_batch_bf16(
        feat, table[0], table[1], ys[0], sel, B, engine.xb, y_out=engine._gy[0], zero=engine._zbuf,
        y2=ys[1] if len(ys) > 1 else None, y2_out=engine._gy[1] if len(ys) > 1 else None))
    if engine._fwd_fused is not None and engine._fuse_loss():
        # the step's own launch: forward chain with the pair loss in its last epilogue
        out["forward"] = t(lambda: engine._fwd_loss(B))
        out["loss"] = 0.0
    else:
        if engine._fwd_fused is not None:
            out["forward"] = t(lambda: ops.mlp_forward_fused(engine.xb, engine._fwd_rows, engine._fwd_fused))
        else:
            out["forward"] = t(lambda: engine._forward_bf16(None))
        o = engine.out_last
        outs = [o[:, :engine.head_dim], o[:, engine.head_dim:]] if engine.heads else o
        out["loss"] = t(lambda: engine._loss_and_seed_bf16(outs, B, engine._gy))
    if engine._dgrad_fused is not None:
        out["dgrad"] = t(lambda: ops.mlp_dgrad_fused(engine.dzb[-1], engine._fwd_rows, engine._dgrad_fused))

    def wgrad():
        for grp in engine._backward_groups:
            ops.gemm_group(grp)
    out["wgrad" if engine._dgrad_fused is not None else "backward"] = t(wgrad)
    if engine.world == 1:
        keep = [x.clone() for x in (engine.bucket.param, engine.state0, engine.state1) if x is not None]
        out["optimizer"] = t(lambda: engine._optimizer(1.0, 1))
        for dst, src in zip([x for x in (engine.bucket.param, engine.state0, engine.state1)
                             if x is not None], keep):
            dst.copy_(src)
        engine.refresh_bf16_weights()
    engine.bucket.trained_grad.zero_()
    engine._grads_clean = True
    return out


# ------------------------------------------------------------------ dp check ---
def dp_check(args, corpus, table_src, world, rank, dev, multitask, steps=12):
    """Outside the timed region (N > 1): K steps from rank 0's initialisation through the
    default exchange, then the same K steps through the NCCL all-reduce; ranks must hold
    bit-identical weights, and the two exchanges must agree to fp32 rounding."""
    import torch
    import torch.distributed as dist
    feat, table = table_src
    B = args.train_batch

    def run(p2p):
        old = os.environ.get("ABN_DP_P2P")
        if p2p is None:
            os.environ.pop("ABN_DP_P2P", None)
        else:
            os.environ["ABN_DP_P2P"] = p2p
        try:
            torch.manual_seed(1234)          # identical init on every rank, for both runs
            net, loss = make_network(multitask, dev)
            from abnet3_b200.trainer import _loss_spec
            from abnet3_b200.engine import SiameseTrainStep
            # plain SGD: the comparison is about the summed gradient (Adadelta's ratio of running
            # averages amplifies the last-bit differences of the two summation orders)
            eng = SiameseTrainStep(net, _loss_spec(loss), "sgd", lr=1e-3, momentum=0.0)
            mode = {0: "push2", 1: "push1", 2: "ll"}[int(eng._dp_push.one_shot)] if eng._dp_push is not None else (
                "reads" if eng._dp is not None else "nccl")
            eng.sweep_table(feat, table, B, steps, start=0, do_training=True)
            torch.cuda.synchronize()
            return eng.bucket.trained_param.clone(), mode
        finally:
            if old is None:
                os.environ.pop("ABN_DP_P2P", None)
            else:
                os.environ["ABN_DP_P2P"] = old

    p_def, mode = run(None)
    p_nccl, _ = run("0")
    bits = p_def.view(torch.int32).to(torch.int64)
    digest = torch.stack([bits.sum(), (bits * torch.arange(1, bits.numel() + 1, device=dev)).sum()])
    all_d = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(all_d, digest)
    identical = all(bool(torch.equal(all_d[0], d)) for d in all_d)
    rel = float(((p_def - p_nccl).double().norm() / p_nccl.double().norm()).item())
    return {"ranks_identical": identical, "rel_vs_nccl": rel, "steps": steps, "exchange": mode,
            "optimizer": "sgd lr 1e-3 (the timed run uses Adadelta through the same exchange kernel)"}


def make_network(multitask, dev):
    from abnet3_b200.model import SiameseNetwork, SiameseMultitaskNetwork
    from abnet3_b200.loss import coscos2, weighted_loss_multi
    if multitask:
        net = SiameseMultitaskNetwork(input_dim=FEAT_DIM, num_hidden_layers_shared=2,
                                      num_hidden_layers_spk=1, num_hidden_layers_phn=1, hidden_dim=500,
                                      output_dim=100, p_dropout=0.0, activation_layer="sigmoid").to(dev)
        loss = weighted_loss_multi(avg=False, loss_phn=coscos2(avg=False), loss_spk=coscos2(avg=False),
                                   weight=0.5)
    else:
        net = SiameseNetwork(input_dim=FEAT_DIM, num_hidden_layers=2, hidden_dim=500, output_dim=100,
                             p_dropout=0.0, activation_layer="sigmoid").to(dev)
        loss = coscos2(avg=False)
    return net, loss


# ---------------------------------------------------------- C3 / C4: the path ---
def run_path(args, rank, world, local_rank, dev):
    import torch
    import torch.distributed as dist
    from abnet3_b200 import synth, utils
    from abnet3_b200.dataloader import FramesDataLoader, MultiTaskFramesDataLoader
    from abnet3_b200.trainer import TrainerSiamese, TrainerSiameseMultitask

    cfg = args.config
    multitask = cfg == "C4"
    metric, unit = METRICS[cfg]
    B = args.train_batch
    P = args.pairs
    corpus = synth.make_corpus(args.tokens, seed=0, device=dev)          # replicated table
    same = synth.make_same_pairs(corpus, P, seed=1 + rank)               # this rank's shard
    diff = synth.make_diff_pairs(corpus, P, seed=100 + rank)
    P_dev = max(P // 100, 64)
    same_dev = synth.make_same_pairs(corpus, P_dev, seed=500 + rank)
    diff_dev = synth.make_diff_pairs(corpus, P_dev, seed=600 + rank)

    def spk_labels(tok):
        """+1 when both tokens come from the same 'speaker' (groups of files), else -1."""
        fo = corpus.file_off
        f1 = torch.searchsorted(fo, tok[:, 0].long(), right=True) - 1
        f2 = torch.searchsorted(fo, tok[:, 2].long(), right=True) - 1
        return torch.where((f1 // 4) == (f2 // 4), 1, -1).to(torch.int8)

    def tokens_of(s, d):
        if multitask:
            return (s, d, [(spk_labels(s), spk_labels(d))])
        return (s, d)

    tokens = {"train": tokens_of(same, diff), "dev": tokens_of(same_dev, diff_dev)}
    torch.manual_seed(0)
    net, loss = make_network(multitask, dev)
    ftable = utils.FeatureTable.from_device(corpus.feat, corpus.file_off)
    LoaderCls = MultiTaskFramesDataLoader if multitask else FramesDataLoader
    TrainerCls = TrainerSiameseMultitask if multitask else TrainerSiamese

    def build(table):
        loader = LoaderCls.from_tokens(table, tokens, batch_size=B, randomize_dataset=True,
                                       max_batches_per_epoch=None, exact_numpy_shuffle=False)
        return loader

    loader = build(ftable)
    trainer = TrainerCls(network=net, loss=loss, optimizer_type="adadelta", lr=0.1, momentum=None,
                         cuda=True, dataloader=loader, log_dir="/tmp/abn_bench_runs")
    engine = trainer.engine
    cap = args.epoch_batches

    if cap:                       # --epoch-batches: a shorter epoch for quick runs (not the headline)
        orig = loader.epoch_table

        def capped(train_mode=True):
            f, t, bs, first, n = orig(train_mode)
            return f, t, bs, first, min(n, cap)
        loader.epoch_table = capped

    stream = torch.cuda.current_stream()
    stats = {}

    def one_step():
        """align every pair + one epoch (train sweep + dev sweep) -> trained frame pairs"""
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record(stream)
        loader.realign("train")
        t1.record(stream)
        trainer.optimize_model(do_training=True)
        stats["align_ev"] = (t0, t1)
        n_b = trainer.last_sweep["train_batches"]
        return n_b * min(B, loader.frame_pairs["train"][0].numel()), n_b

    with contextlib.redirect_stdout(sys.stderr):
        loader.load_data()
        for _ in range(max(args.warmup, 3)):
            one_step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
            time.sleep(0.2)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()                 # (rank 0 just slept: nobody may start the clock early)
            torch.cuda.synchronize()
        w0 = time.perf_counter()
        beg, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        beg.record(stream)
        fp_total, nb_total, align_ms = 0, 0, 0.0
        evs = []
        for _ in range(args.steps):
            fp, n_b = one_step()
            fp_total += fp
            nb_total += n_b
            evs.append(stats["align_ev"])
        end.record(stream)
        torch.cuda.synchronize()
        w1 = time.perf_counter()
        if world > 1:
            dist.barrier()
        clocks = sampler.stop(w0, w1) if rank == 0 else None
    total_ms = max_over_ranks(beg.elapsed_time(end), dev, world)
    align_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    fp_all = sum_over_ranks(fp_total, dev, world)
    value = fp_all / (total_ms * 1e-3)
    n_table = int(loader.frame_pairs["train"][0].numel())
    train_loss, dev_loss = trainer.train_losses[-1], trainer.dev_losses[-1]

    # training sweep alone (graph replays, no alignment / shuffle / dev sweep): us per step
    feat_t, table, bs, first, n_b = loader.epoch_table(True)
    n_b = trainer._agreed_batches(min(n_b, 2000)) if world > 1 else min(n_b, 2000)
    engine.sweep_table(feat_t, table, bs, 5, start=first, do_training=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    engine.sweep_table(feat_t, table, bs, n_b, start=first, do_training=True)
    b.record(stream)
    torch.cuda.synchronize()
    step_us = max_over_ranks(a.elapsed_time(b), dev, world) / n_b * 1e3
    pk = peaks()
    step_tflops = bs * MLP_FLOP[cfg] / (step_us * 1e-6) / 1e12

    ktimes = None
    if not args.no_kernel_times:
        ktimes = kernel_times(engine, feat_t, table, bs)
    tr = ncu_traffic()
    fwd_us = ktimes["forward"] if ktimes else None
    roofline = None
    if fwd_us:
        ach = bs * MLP_FLOP_FWD[cfg] / (fwd_us * 1e-6) / 1e12
        roofline = {
            "bound": "tensor", "achieved": ach, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
            "frac": ach / pk["bf16_tflops"], "peak_source": pk["source"] + " bf16_tflops (burst: the "
            "kernel is timed alone)",
            "kernel": "mlp_chain_kernel<0> (forward chain: every layer in one launch, tcgen05 "
                      "cta_group::2, activations resident in shared memory%s)"
                      % (", pair loss + output-layer dz in the last epilogue" if engine._fuse_loss() else ""),
            "kernel_us": fwd_us, "algorithmic_flops_per_launch": bs * MLP_FLOP_FWD[cfg],
            "algorithmic_flops_def": "2 rows x 2 x sum(n_in*n_out) per frame pair, unpadded shapes "
                                     "(SURVEY 8d)",
            "traffic": tr.get("mlp_forward"),
            "step": {"us": step_us, "achieved": step_tflops, "peak": pk["bf16_tflops_sustained"],
                     "frac": step_tflops / pk["bf16_tflops_sustained"],
                     "peak_source": pk["source"] + " bf16_tflops_sustained (inside a long sweep)",
                     "flops_per_step": bs * MLP_FLOP[cfg]},
            "kernels_us_alone": ktimes,
        }

    # ---- e2e: the same step from HOST buffers ---------------------------------
    e2e = None
    if not args.no_e2e:
        with contextlib.redirect_stdout(sys.stderr):
            feat = corpus.feat
            host_feat = torch.empty(feat.shape, dtype=feat.dtype, pin_memory=True)
            host_feat.copy_(feat)
            flat = []
            for mode in ("train", "dev"):
                for tkn in tokens[mode][:2]:
                    h = torch.empty(tkn.shape, dtype=tkn.dtype, pin_memory=True)
                    h.copy_(tkn)
                    flat.append(h)
            hl = [[torch.empty(c.shape, dtype=c.dtype, pin_memory=True).copy_(c) for c in pl]
                  for mode in ("train", "dev") for pl in (tokens[mode][2] if multitask else [])]
            file_off = corpus.file_off.tolist()
            torch.cuda.synchronize()
            h2d = host_feat.numel() * 4 + sum(h.numel() * 4 for h in flat) + \
                sum(c.numel() for pl in hl for c in pl)

            def e2e_step():
                tab = utils.FeatureTable.from_host(host_feat, file_off, device=dev)
                dt = [h.to(dev, non_blocking=True) for h in flat]
                if multitask:
                    dl = [[c.to(dev, non_blocking=True) for c in pl] for pl in hl]
                    tk = {"train": (dt[0], dt[1], [tuple(dl[0])]), "dev": (dt[2], dt[3], [tuple(dl[1])])}
                else:
                    tk = {"train": (dt[0], dt[1]), "dev": (dt[2], dt[3])}
                ld = LoaderCls.from_tokens(tab, tk, batch_size=B, randomize_dataset=True,
                                           exact_numpy_shuffle=False)
                if cap:
                    o2 = ld.epoch_table
                    ld.epoch_table = lambda train_mode=True: (lambda r: r[:4] + (min(r[4], cap),))(o2(train_mode))
                trainer.dataloader = ld
                ld.load_data()                       # alignment of train + dev pairs
                trainer.optimize_model(do_training=True)     # losses are read back (2 x float64)
                return trainer.last_sweep["train_batches"] * min(B, ld.frame_pairs["train"][0].numel())

            e2e_step()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e_steps = 2
            t0 = time.perf_counter()
            fp_e = 0
            for _ in range(e_steps):
                fp_e += e2e_step()
            torch.cuda.synchronize()
            dt_e = max_over_ranks(time.perf_counter() - t0, dev, world)
            trainer.dataloader = loader
        e2e = {"value": sum_over_ranks(fp_e, dev, world) / dt_e, "unit": unit,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 16, "steps": e_steps,
               "call": "FeatureTable.from_host(pinned [N,280] table) -> %s.from_tokens(pair lists "
                       "from pinned host memory).load_data() -> %s.optimize_model(): whole table + "
                       "pair lists uploaded, epoch losses read back, every step"
                       % (LoaderCls.__name__, TrainerCls.__name__)}
        del host_feat

    # ---- alignment leg (C2 sub-record) ------------------------------------------
    align = None
    if not args.no_align_leg:
        del loader.frame_pairs["train"]
        loader.frame_pairs["train"] = None
        torch.cuda.empty_cache()
        align, _ = align_leg(args, corpus, same, world, rank, local_rank, dev, args.align_steps,
                             sample_clocks=False, with_e2e=(world == 1))
        with contextlib.redirect_stdout(sys.stderr):
            loader.realign("train")

    # ---- data-parallel correctness (outside the timed region) -------------------
    dpc = None
    if world > 1 and not args.no_dp_check:
        with contextlib.redirect_stdout(sys.stderr):
            dpc = dp_check(args, corpus, (corpus.feat, loader.frame_pairs["train"]), world, rank, dev,
                           multitask)

    # ---- CPU baseline (rank 0, N == 1 only) --------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        small = synth.make_corpus(8000, seed=0, device="cpu")
        n_s = args.cpu_sample or 3000
        s_h = synth.make_same_pairs(small, n_s, seed=1).numpy()
        d_h = synth.make_diff_pairs(small, n_s, seed=2).numpy()
        rate, ta, tt, fp = cpu_path_rate(small.feat.numpy(), s_h, d_h, cores, B, multitask)
        cpu = {"value": rate, "unit": unit, "cores": cores, "kind": "port",
               "sample": "%d same + %d diff pairs of the step's pair distribution: oracle alignment "
                         "(numpy cosine_distance + C DTW) on %d processes %.2f s, then %d frame pairs "
                         "through the reference MLP + coscos2 + Adadelta on torch CPU (%d threads) "
                         "%.1f s" % (n_s, n_s, cores, ta, fp, cores, tt)}

    if rank == 0:
        n_train_b = nb_total // args.steps
        n_dev_b = max(int(loader.frame_pairs["dev"][0].numel()) // B, 1)
        kernels_per_step = 6 if engine._dgrad_fused is not None else 5
        if multitask:
            kernels_per_step += 1
        elif engine._fuse_loss():
            kernels_per_step -= 1          # the loss is computed inside the forward chain launch
        launches = args.steps * ((align or {}).get("launches_per_pass", 0) + 4 +
                                 n_train_b * kernels_per_step + n_dev_b * 3)
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16 tcgen05 (fp32 accumulate, fp32 master weights); f32 distance, f64 DTW",
            "data": "synthetic",
            "config": {
                "workload": WORKLOADS[cfg], "same_pairs_per_gpu": P, "diff_pairs_per_gpu": P,
                "frame_pairs_in_table_per_gpu": n_table, "batch": B,
                "train_batches_per_epoch": n_train_b, "dev_pairs_per_gpu": 2 * P_dev,
                "corpus_frames": int(corpus.feat.shape[0]), "corpus_bytes": int(corpus.feat.numel() * 4),
                "surface": "%s.realign() + %s.optimize_model()" % (LoaderCls.__name__, TrainerCls.__name__),
                "l2": "inputs larger than L2 (feature table %.1f GB; every batch gathers 16384 random "
                      "rows of it)" % (corpus.feat.numel() * 4 / 1e9),
                "epoch_batches_cap": cap or None,
            },
            "phases_ms": {"align_and_shuffle": align_ms, "epoch": total_ms / args.steps - align_ms},
            "losses": {"train": train_loss, "dev": dev_loss},
            "train_step": {"us": step_us, "frame_pairs_per_s": world * bs / (step_us * 1e-6),
                           "tflops": step_tflops,
                           "tensor_util_sustained": step_tflops / pk["bf16_tflops_sustained"],
                           "tensor_util_burst": step_tflops / pk["bf16_tflops"]},
            "roofline": roofline,
            "align": align,
            "e2e": e2e,
            "cpu_baseline": cpu,
            "dp_check": dpc,
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)


# -------------------------------------------------------------- C2: align only ---
def run_align_only(args, rank, world, local_rank, dev):
    import torch
    from abnet3_b200 import synth
    corpus = synth.make_corpus(args.tokens, seed=0, device=dev)
    pairs = synth.make_same_pairs(corpus, args.pairs, seed=1 + rank)
    rec, _ = align_leg(args, corpus, pairs, world, rank, local_rank, dev, args.steps)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        feat_h, pairs_h = corpus.feat.cpu().numpy(), pairs.cpu().numpy()
        n_s = args.cpu_sample
        if not n_s:
            rate0, _, _, _ = cpu_align_rate(feat_h, pairs_h[:1024 * cores], cores)
            n_s = int(max(8192 * cores, rate0 * 15.0))
        n_s = min(n_s, len(pairs_h))
        rate, dt, _, _ = cpu_align_rate(feat_h, pairs_h[:n_s], cores)
        cpu = {"value": rate, "unit": "pairs/s", "cores": cores, "kind": "port",
               "sample": "%d pairs of the step's pair list, %.1f s on %d processes (numpy "
                         "cosine_distance + C DTW oracle + row gather per pair)" % (n_s, dt, cores)}
    if rank == 0:
        line = {
            "metric": "dtw_pairs_per_sec", "value": rec["value"], "unit": "pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": rec["ms_per_pass"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 distance, f64 DTW accumulate", "data": "synthetic",
            "config": {"workload": WORKLOADS["C2"], "pairs_per_gpu_per_step": args.pairs,
                       "corpus_frames": int(corpus.feat.shape[0]),
                       "l2": "inputs larger than L2 (feature table %.1f GB, every step re-reads it)"
                             % (corpus.feat.numel() * 4 / 1e9)},
            "roofline": rec["roofline"], "e2e": rec.get("e2e"),
            "e2e_unstacked": rec.get("e2e_unstacked"), "cpu_baseline": cpu,
            "gpu_launches": args.steps * rec["launches_per_pass"], "clocks": rec["clocks"],
            "align": {k: rec[k] for k in ("valid_pairs", "mean_path_len", "aligned_frame_pairs_per_s")},
        }
        print(json.dumps(line), flush=True)


# ------------------------------------------------------------ C5: length sweep ---
def run_sweep(args, rank, world, local_rank, dev):
    import torch
    import torch.distributed as dist
    from abnet3_b200 import synth, utils, ops, _lib
    pk = peaks()
    stream = torch.cuda.current_stream()
    lengths = (50, 100, 200, 400)
    total_pairs = args.pairs                    # per length, over ALL ranks (strong scaling)
    per_len, t_sum, p_sum, launches = [], 0.0, 0, 0
    clocks = None
    for n in lengths:
        n_tok = max(2000, min(args.tokens, 4_000_000 // n))
        corpus = synth.make_corpus(n_tok, seed=n, device=dev, len_range=(n, n),
                                   tokens_per_file=max(50, 100_000 // n))
        allp = synth.make_same_pairs(corpus, total_pairs, seed=1)
        lo, hi = rank * total_pairs // world, (rank + 1) * total_pairs // world
        pairs = allp[lo:hi].contiguous()
        feat = corpus.feat
        last = torch.zeros(feat.shape[0], dtype=torch.uint8, device=dev)
        last[(corpus.file_off[1:] - 1).long()] = 1
        stack = 7 if ops.stack_violations(feat, 7, last) == 0 else 0
        al = utils.BatchAligner(feat, max_pairs=pairs.shape[0], max_frames=n, stack=stack)
        for _ in range(max(args.warmup, 3)):
            res = al.align(pairs)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0 and n == lengths[-1]:
            sampler.start()
            time.sleep(0.2)
        w0 = time.perf_counter()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(args.steps):
            res = al.align(pairs)
        b.record(stream)
        torch.cuda.synchronize()
        w1 = time.perf_counter()
        if rank == 0 and n == lengths[-1]:
            clocks = sampler.stop(w0, w1)
        ms = max_over_ranks(a.elapsed_time(b), dev, world) / args.steps
        L_total = sum_over_ranks(float(res.path_len.long().sum().item()), dev, world)
        alg = 4 * FEAT_DIM * 2 * n * total_pairs + 8 * L_total + 16 * total_pairs
        flops = 2.0 * FEAT_DIM * n * n * total_pairs
        rate = total_pairs / (ms * 1e-3)
        launches += args.steps * int(_lib.lib().abn_align_launches(pairs.shape[0], n, stack, al._ws_bytes))
        per_len.append({"frames_per_token": n, "pairs": total_pairs, "ms": ms, "pairs_per_s": rate,
                        "roofline_frac": alg / (ms * 1e-3) / 1e9 / (pk["hbm_gbs"] * world),
                        "achieved_gbs": alg / (ms * 1e-3) / 1e9,
                        "fp32_tflops_280deep": flops / (ms * 1e-3) / 1e12, "stacked_fast_path": bool(stack)})
        t_sum += ms
        p_sum += total_pairs
        del al, corpus, feat
        torch.cuda.empty_cache()
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        cpu_l = []
        for n in lengths:
            c = synth.make_corpus(max(400, 40_000 // n), seed=n, device="cpu", len_range=(n, n),
                                  tokens_per_file=200)
            ps = synth.make_same_pairs(c, max(cores * 8, int(3.0e9 / (n * n))), seed=1).numpy()
            r, dt, _, _ = cpu_align_rate(c.feat.numpy(), ps, cores)
            cpu_l.append({"frames_per_token": n, "pairs_per_s": r, "sample_pairs": len(ps), "seconds": dt})
        hm = len(cpu_l) / sum(1.0 / x["pairs_per_s"] for x in cpu_l)
        cpu = {"value": hm, "unit": "pairs/s", "cores": cores, "kind": "port", "per_length": cpu_l,
               "sample": "per length a sample of about 3e9 / n^2 pairs on %d processes; value = pairs/s "
                         "of equal pair counts at the four lengths (harmonic mean)" % cores}
    if rank == 0:
        worst = min(per_len, key=lambda r: r["roofline_frac"])
        line = {
            "metric": "dtw_pairs_per_sec", "value": p_sum / (t_sum * 1e-3), "unit": "pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": t_sum,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32 distance, f64 DTW accumulate", "data": "synthetic",
            "config": {"workload": WORKLOADS["C5"], "pairs_per_length": total_pairs,
                       "lengths": list(lengths), "l2": "inputs larger than L2 (feature tables 1-4 GB)"},
            "per_length": per_len,
            "roofline": {"bound": "hbm", "achieved": worst["achieved_gbs"], "peak": pk["hbm_gbs"] * world,
                         "unit": "GB/s", "frac": worst["roofline_frac"],
                         "kernel": "alignment kernels at %d frames/token (the length furthest from the "
                                   "roofline)" % worst["frames_per_token"],
                         "peak_source": pk["source"] + " hbm_gbs x n_gpus", "traffic": None},
            "cpu_baseline": cpu, "e2e": None, "gpu_launches": launches, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    try:
        if args.config in ("C3", "C4"):
            run_path(args, rank, world, local_rank, dev)
        elif args.config == "C2":
            run_align_only(args, rank, world, local_rank, dev)
        else:
            run_sweep(args, rank, world, local_rank, dev)
    finally:
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
