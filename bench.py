#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: hybrid top-10 queries/sec at 10M x 1536 chunks on 1/2/4/8 B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config 3|2|4|pairwise] [--batch B]
    torchrun ... bench.py --gpus N ...        (one rank per GPU, corpus row-sharded)

Default (--config 3, the configuration BASELINE's metric is quoted on, sized for N GPUs): a "step" is one hybrid
retrieval (exact cosine top-10 + BM25 top-10 + RRF) of a batch of 256 queries against the whole synthetic corpus
(10M chunks x 1536-d fp32 + Zipf token corpus, V=50k).  `value` = whole-job queries/s with inputs resident in HBM;
`e2e` = the same through the public search call with HOST query buffers (H2D + D2H inside the timed region).
The other BASELINE configs print the same contract line for their own workload:
    --config 2         1M x 1536 exact cosine top-10, batch 256 (one GPU)
    --config 4         BM25-only over the 10M-doc Zipf corpus, batch 1024
    --config pairwise  consistency-checker claim pairs, 65536 x 1536, threshold 0.85 (config 5, one GPU)
`--batch B` changes the query batch (SURVEY.md §8d sweep B in {1, 16, 64, 256, 1024}).  One JSON line on rank 0.

Outside the timed regions the results of the last step are compared with the CPU oracle at FULL corpus size on a
query subset (`verified_against_oracle`); the time the oracle takes for that is the reported `cpu_baseline`.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

UNIT = "queries/s"
DIM = 1536
VOCAB = 50000
TOPK = 10
LMIN, LMAX = 100, 300

CONFIGS = {
    "3": {"metric": "hybrid_top10_queries_per_sec_10Mx1536", "rows": 10_000_000, "batch": 256},
    "2": {"metric": "cosine_top10_queries_per_sec_1Mx1536", "rows": 1_000_000, "batch": 256},
    "4": {"metric": "bm25_top10_queries_per_sec_10M_docs", "rows": 10_000_000, "batch": 1024},
    "pairwise": {"metric": "consistency_claim_pairs_per_sec_64kx1536", "rows": 65536, "batch": 0},
}


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default="3", choices=list(CONFIGS))
    ap.add_argument("--rows", type=int, default=None)
    ap.add_argument("--batch", "--queries", dest="batch", type=int, default=None)
    ap.add_argument("--mode", default=os.environ.get("ORAG_BENCH_MODE", "f16"), choices=["tf32", "bf16", "f16"])
    ap.add_argument("--tile-docs", type=int, default=2048)
    ap.add_argument("--verify-queries", type=int, default=None,
                    help="queries compared with the full-size CPU oracle after the timed loops (0 = skip); default 32 "
                         "on one GPU, 8 under torchrun")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="alias of --verify-queries 0")
    ap.add_argument("--timeline", default=None, metavar="FILE",
                    help="after the timed loop, record start/end of every tagged launch of 8 more submitted steps "
                         "(orag_timeline_*) and write them to FILE (.rankN appended under torchrun)")
    ap.add_argument("--ref-sample-rows", type=int, default=None)
    ap.add_argument("--ref-sample-queries", type=int, default=16)
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.rows is None:
        args.rows = int(os.environ.get("ORAG_BENCH_ROWS", cfg["rows"]))
    if args.batch is None:
        args.batch = cfg["batch"]
    args.metric = cfg["metric"]
    if args.no_cpu_baseline:
        args.verify_queries = 0
    return args


def workload_config(args, extra=None):
    if args.config == "3":
        what = (f"{args.rows} chunks x {DIM}-d fp32 hybrid (exact cosine + BM25 + RRF) top-{TOPK}, query batch "
                f"{args.batch}, Zipf(s=1) token corpus V={VOCAB} L~U[{LMIN},{LMAX}]")
    elif args.config == "2":
        what = f"{args.rows} chunks x {DIM}-d fp32 exact cosine top-{TOPK}, query batch {args.batch}"
    elif args.config == "4":
        what = (f"BM25-only top-{TOPK} over {args.rows} chunks, Zipf(s=1) token corpus V={VOCAB} L~U[{LMIN},{LMAX}], "
                f"query batch {args.batch}")
    else:
        what = (f"consistency-checker claim pairs: all i<j of {args.rows} claims x {DIM}-d with doc_idx = i // 16 and "
                f"float64 cosine >= 0.85 (rag/consistency_checker.py:169-189)")
    cfg = {"workload": what, "baseline_config": args.config, "rows": args.rows, "dim": DIM, "query_batch": args.batch,
           "top_k": TOPK, "vocab": VOCAB,
           "l2": "inputs larger than L2 (corpus / postings streamed from HBM every step)"}
    if extra:
        cfg.update(extra)
    return cfg


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """The reference's CPU path (oracle port of rag/retrieval.py:362-371, 324-347 and rag/reranker.py:224-271; the
    Python reference itself cannot travel to the GPU box) on all host cores.  Each step is a bounded sample of the
    workload -- `--ref-sample-queries` queries against the first rows/8 rows and docs, regenerated from the seeds --
    scaled linearly in N (every piece is O(N) per query)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    from oracle import scale_check
    from optimized_rag_b200 import synthetic as syn
    cores = os.cpu_count() or 1
    oracle.build()
    oracle.set_threads(cores)  # torchrun exports OMP_NUM_THREADS=1: the baseline is "all host cores"
    S = args.ref_sample_rows or max(args.rows // 8, 1000)
    S = min(S, args.rows)
    Bs = args.ref_sample_queries
    thr = syn.zipf_thresholds(VOCAB)
    if args.config == "pairwise":
        emb = pairwise_claims_host(S)
        doc = (np.arange(S) // 16).astype(np.int32)

        def step():
            t0 = time.perf_counter()
            oracle.pairwise_candidates_parallel(emb, doc, 0.85)
            return time.perf_counter() - t0
        unit, per_step = "pairs/s", S * (S - 1) / 2
        scale = (args.rows * (args.rows - 1) / 2) / per_step
        sample = f"all pairs of the first {S} claims per step (O(M^2): scaled x{scale:.0f} in time to {args.rows} claims)"
        value_of = lambda t_total, steps: (args.rows * (args.rows - 1) / 2) / (t_total / steps * scale)
    else:
        nq_total = max(args.batch, Bs)
        q = syn.query_embeddings(nq_total, args.rows, DIM)[:Bs] if args.config in ("3", "2") else None
        qt, ql = syn.keyword_queries(nq_total, VOCAB, thresholds=thr)
        qt, ql = qt[:Bs], ql[:Bs]
        bm25 = oracle.StreamedBM25(syn.SEED_TOKENS, S, VOCAB, LMIN, LMAX, thr) if args.config in ("3", "4") else None

        def step():
            t0 = time.perf_counter()
            scale_check.reference_lists(S, DIM, q, qt, ql, TOPK, seed_corpus=syn.SEED_CORPUS, bm25=bm25,
                                        want_cosine=args.config in ("3", "2"))
            return time.perf_counter() - t0
        unit = UNIT
        scale = args.rows / S
        sample = (f"{Bs} queries x first {S} rows/docs per step (oracle: cosine fp64 Neumaier + BM25 get_scores + RRF "
                  f"over inputs regenerated from the seeds), scaled x{scale:.0f} to {args.rows} rows (O(N) per query); "
                  f"BM25Okapi rebuild per call (rag/retrieval.py:338) NOT charged")
        value_of = lambda t_total, steps: (Bs * steps) / (t_total * scale)
    for _ in range(args.warmup):
        step()
    t = sum(step() for _ in range(args.steps))
    value = value_of(t, args.steps)
    line = {"impl": "reference", "metric": args.metric, "value": value, "unit": unit, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / max(args.steps, 1),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def pairwise_claims_host(m: int) -> np.ndarray:
    """Config-5 claims on the host (numpy twin of `pairwise_claims`): synthetic rows, every 64th one a noisy copy of
    another (cosine ~0.85-0.97 with its source)."""
    from optimized_rag_b200 import synthetic as syn
    emb = syn.embeddings(syn.SEED_CORPUS, 0, m, DIM)
    dst = np.arange(0, m, 64)
    src = (dst * 7919 + 13) % m
    w = np.float32(0.25) + np.float32(0.4) * ((dst * 2654435761 % 1000).astype(np.float32) / np.float32(1000))
    emb[dst] = emb[src] + w[:, None] * emb[dst]
    return emb


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi polled every 25 ms (rank 0 only) from BEFORE the warm-up (its start-up alone takes ~0.1 s); only the samples whose
    timestamps fall inside the device-timed loop count (`window(t0, t1)` marks it; the GPU is continuously busy there,
    whereas the end-to-end loop idles between steps and would show ramped-down clocks), so that a 30 ms multi-GPU run still
    gets its clocks and idle set-up time never dilutes them."""
    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,power.draw.instant,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap,power.limit")

    def __init__(self, uuid: str):
        self.uuid = uuid
        self.proc = None
        self.windows = []
        self.path = ROOT / "gpurun_out" / f"clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.path.parent.mkdir(exist_ok=True)
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", self.uuid, f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "25"], stdout=self.fh,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def window(self, t0, t1):
        self.windows.append((t0, t1))

    def stop(self):
        import datetime as dt
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for ln in self.path.read_text().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                ts = dt.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[1]), float(f[2]), [n for n, v in zip(names, f[4:8]) if v.lower().startswith("active")],
                             _num(f[3]), _num(f[8]) if len(f) > 8 else None))
            except ValueError:
                continue
        try:
            self.path.unlink()
        except OSError:
            pass
        inside = [r for r in rows if any(a <= r[0] <= b for a, b in self.windows)]
        note = "inside the timed regions"
        if not inside and rows and self.windows:  # run shorter than the polling period: nearest samples under the same load
            a, b = min(w[0] for w in self.windows), max(w[1] for w in self.windows)
            inside = [r for r in rows if a - 0.25 <= r[0] <= b + 0.25]
            note = "nearest to the timed regions (+-0.25 s, same workload: warm-up loop)"
        if inside:
            out.update(sm_mhz=statistics.median(r[1] for r in inside), sm_max_mhz=max(r[2] for r in inside),
                       reasons=sorted({n for r in inside for n in r[3]}), samples=len(inside), window=note)
            watts = [r[4] for r in inside if r[4] is not None]
            limit = [r[5] for r in inside if r[5] is not None]
            if watts:   # the hybrid step runs at the board's power limit: the SM clock is what the cap leaves
                out.update(power_w=statistics.median(watts), power_limit_w=max(limit) if limit else None)
        return out


def _num(text):
    try:
        return float(text)
    except ValueError:
        return None


# ------------------------------------------------------------------------------------------------ native arm
class Harness:
    """Process-group set-up, barriers, the device-timed loop, the end-to-end loop, clocks: shared by all configs."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from optimized_rag_b200 import _ffi
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py (native arm) needs a GPU: there is no CPU fallback for the retrieval hot path")
        torch.cuda.set_device(local_rank)
        self.dev = torch.device("cuda", local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.L = _ffi.lib()
        uuid = str(torch.cuda.get_device_properties(self.dev).uuid)
        self.sampler = ClockSampler(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
        self.warmup = max(args.warmup, 3)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, ms: float) -> float:
        if self.world == 1:
            return ms
        t = self.torch.tensor([ms], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def device_timed(self, step, drain=None):
        """W warm-up steps, then exactly K steps between two events (barrier + synchronize on both sides, max over
        ranks).  `step()` enqueues one step without synchronising with the host; `drain()` (optional) makes the
        current stream wait for everything the steps submitted and is called before every closing event / barrier.
        Returns (total ms, launches, brackets[, what the timed loop's drain() returned]) where brackets = per-step
        CUDA-event durations of the two dominant kernels (library hooks)."""
        drain = drain or (lambda: None)
        torch, L, args = self.torch, self.L, self.args
        if self.rank == 0:  # one poller per box: nvidia-smi queries take driver locks that kernel launches also need
            self.sampler.start()
        for _ in range(self.warmup):
            step()
        drain()
        self.barrier()
        L.orag_profile_enable(1)
        launches0 = int(L.orag_launch_count())
        self.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.time()
        ev0.record()
        for _ in range(args.steps):
            step()
        drained = drain()
        ev1.record()
        self.barrier()
        w1 = time.time()
        self.sampler.window(w0, w1)
        dev_ms = self.max_over_ranks(ev0.elapsed_time(ev1))
        launches = int(L.orag_launch_count()) - launches0
        brackets = []
        buf = (ctypes.c_float * 256)()
        for slot in (0, 1):
            n = int(L.orag_profile_read_all(slot, buf, 256))
            brackets.append([float(buf[i]) for i in range(max(n, 0))])
        L.orag_profile_enable(0)
        if dev_ms < 150.0:
            # the timed loop is shorter than a few polling periods (small shards): keep the same back-to-back load
            # running, untimed, for ~0.2 s so that the poller sees it.  The step count derives from dev_ms, which is
            # identical on every rank (max over ranks), so all ranks enter the same number of collectives.
            n_probe = int(200.0 / max(dev_ms / args.steps, 1e-3)) + 1
            p0 = time.time()
            for _ in range(n_probe):
                step()
            drain()
            torch.cuda.synchronize()
            self.sampler.window(p0, time.time())
            self.barrier()
        # the clocks line describes the device-timed loop: stop polling before the end-to-end loop, where every step
        # synchronises with the host and a poller taking driver locks would be measured with it
        self.clocks = self.sampler.stop()
        return (dev_ms, launches, brackets) if drained is None else (dev_ms, launches, brackets, drained)

    def kernel_brackets(self, plain_step):
        """Per-launch CUDA-event durations of the two dominant kernels over K PLAIN steps (one batch at a time, nothing
        else in flight): the figures the rooflines are computed from.  (Inside the submitted loop above two batches
        overlap, so a bracket there also measures the other batch's kernels competing for the SMs.)"""
        L, torch = self.L, self.torch
        for _ in range(3):
            plain_step()
        torch.cuda.synchronize()
        L.orag_profile_enable(1)
        for _ in range(self.args.steps):
            plain_step()
        torch.cuda.synchronize()
        buf = (ctypes.c_float * 256)()
        out = []
        for slot in (0, 1):
            n = int(L.orag_profile_read_all(slot, buf, 256))
            out.append([float(buf[i]) for i in range(max(n, 0))])
        L.orag_profile_enable(0)
        self.barrier()
        return out

    def e2e_timed(self, step_with_copies, drain=None):
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(self.args.steps):
            step_with_copies()
        if drain:
            drain()
        e1.record()
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1))

    def finish(self, line, closer=None):
        if self.rank == 0:
            print(json.dumps(line), flush=True)
        if self.world > 1:
            if closer:
                closer()  # collective: nobody frees a buffer a peer still has mapped
            self.dist.destroy_process_group()


TIMELINE_TAGS = ["", "query_prep", "seed_scan", "seed_finalize", "main_scan", "prefilter", "rescore", "select",
                 "bm25_prepare", "bm25_first_pass", "bm25_finalize", "rrf", "push", "merge", "query_sq", "wait"]


def write_timeline(h, step, drain, path, n_steps=8):
    """Diagnostic: the per-launch picture of `n_steps` submitted steps (how the batches in flight overlap)."""
    L, torch = h.L, h.torch
    for _ in range(4):
        step()
    drain()
    h.barrier()
    L.orag_timeline_enable(1)
    for _ in range(n_steps):
        step()
    drain()
    torch.cuda.synchronize()
    cap = 4096
    tags, t0, t1 = (ctypes.c_int * cap)(), (ctypes.c_float * cap)(), (ctypes.c_float * cap)()
    n = int(L.orag_timeline_read(tags, t0, t1, cap))
    L.orag_timeline_enable(0)
    h.barrier()
    rows = sorted((float(t0[i]), float(t1[i]), TIMELINE_TAGS[tags[i]], i) for i in range(max(n, 0)))
    with open(path, "w") as f:
        f.write("# start_ms end_ms dur_ms tag (events on the launching stream: `start` = the stream reached the launch, "
                "`end` = the kernel(s) finished)\n")
        for a, b, tag, _ in rows:
            f.write(f"{a:9.4f} {b:9.4f} {b - a:8.4f} {tag}\n")
        scans = [(a, b) for a, b, tag, _ in rows if tag == "main_scan"]
        if len(scans) > 1:
            gaps = [scans[i + 1][0] - scans[i][1] for i in range(len(scans) - 1)]
            per = [scans[i + 1][1] - scans[i][1] for i in range(len(scans) - 1)]
            f.write("# main scan: durations " + " ".join(f"{b - a:.4f}" for a, b in scans) + "\n")
            f.write("# gap between the end of one main scan and the start event of the next: " +
                    " ".join(f"{g:.4f}" for g in gaps) + "\n")
            f.write("# end-to-end period (scan end to scan end): " + " ".join(f"{g:.4f}" for g in per) + "\n")


def bracket_stats(ms_list, per_step):
    """Per-step durations of a kernel that launches `per_step` times per step -> (mean, min, n_steps) in ms."""
    xs = [x for x in ms_list if x >= 0]
    if not xs:
        return None, None, 0
    if per_step > 1:
        xs = [sum(xs[i:i + per_step]) for i in range(0, len(xs) - per_step + 1, per_step)]
    return statistics.mean(xs), min(xs), len(xs)


TIMING_NOTE = ("CUDA events on the launching stream around every launch of K plain steps run right after the timed loop "
               "(mean; min alongside); `launch_ms_in_timed_loop` = the same bracket inside the timed loop, where two "
               "submitted batches overlap")


def tensor_roofline(flops_per_step, t_mean_ms, t_min_ms, n, pk, kernel, launches_per_step, traffic=None, extra=None):
    ach = flops_per_step / (t_mean_ms * 1e-3) / 1e12
    r = {"bound": "tensor", "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"],
         "traffic": traffic, "frac_sustained": ach / pk["tf_sustained"], "frac_burst": ach / pk["tf_burst"],
         "peak_burst": pk["tf_burst"],
         "peak_kind": "bf16 dense, measured: sustained (back-to-back 4 s) is `peak`, best-of-10 burst is `peak_burst`",
         "kernel": kernel, "peak_source": pk["source"], "launch_ms": t_mean_ms / launches_per_step,
         "launch_ms_min": t_min_ms / launches_per_step, "launches_per_step": launches_per_step, "samples": n,
         "timing": TIMING_NOTE}
    if extra:
        r.update(extra)
    return r


def hbm_roofline(bytes_per_step, t_mean_ms, t_min_ms, n, pk, kernel, launches_per_step, what, traffic=None):
    ach = bytes_per_step / (t_mean_ms * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
            "traffic": traffic, "bytes": what, "algorithmic_bytes": bytes_per_step, "kernel": kernel,
            "peak_source": pk["source"], "launch_ms": t_mean_ms / launches_per_step,
            "launch_ms_min": t_min_ms / launches_per_step, "launches_per_step": launches_per_step, "samples": n,
            "timing": TIMING_NOTE}


def traffic_for(key, rows_per_gpu, batch):
    """ncu-measured DRAM bytes per launch of a dominant kernel, only if the capture was of this launch shape."""
    tp = ROOT / "profiles" / "traffic.json"
    tj = json.loads(tp.read_text()) if tp.exists() else {}
    cap = tj.get("captured_at", {})
    if cap.get("rows_per_gpu") == rows_per_gpu and cap.get("queries") == batch:
        return tj.get(key)
    return None


def scan_rooflines(args, n_local, brackets, pk):
    """Roofline of the cosine main scan from the per-launch brackets (slot 0)."""
    Bq = args.batch
    groups = (Bq + 255) // 256
    mean, mn, n = bracket_stats(brackets[0], groups)
    if mean is None:
        return None
    n_scan_rows = max(n_local - 2048, 0)  # the first 2048 rows are the dense seed pass
    half = args.mode in ("bf16", "f16")
    streamed = n_scan_rows * DIM * (2 if half else 4) * groups
    flops = 2.0 * Bq * n_scan_rows * DIM
    traffic = traffic_for(f"cosine_scan_{args.mode}", n_local, Bq)
    kernel = f"cosine_scan_kernel<{args.mode}> (main scan, {n_scan_rows} rows x {min(Bq, 256)} queries x {groups} group(s))"
    hbm = hbm_roofline(streamed, mean, mn, n, pk, kernel, groups,
                       f"{args.mode} shadow copy actually streamed (N*D*2)" if half else "fp32 corpus (N*D*4)", traffic)
    tens = tensor_roofline(flops, mean, mn, n, pk, kernel, groups, traffic,
                           None if half else {"note": "tf32 runs at half the bf16 rate: x2 for the tf32 ceiling"})
    # which resource bounds the kernel: 16-bit operands at B = 256 have 256 flop/B of streamed data > the ~207 flop/B ridge
    primary, other = (tens, hbm) if (half and min(Bq, 256) >= 208) else (hbm, tens)
    out = dict(primary)
    fp32_bytes = n_scan_rows * DIM * 4 * groups
    out.update({"fp32_equivalent_gbs": fp32_bytes / (mean * 1e-3) / 1e9,
                "fp32_equivalent_frac_of_hbm_peak": fp32_bytes / (mean * 1e-3) / 1e9 / pk["hbm_gbs"],
                "other_view": {k: other[k] for k in ("bound", "achieved", "peak", "unit", "frac")}})
    return out


def bm25_roofline(bm25, q_tok, q_len, brackets, pk, n_local, Bq):
    mean, mn, n = bracket_stats(brackets[1], 1)
    if mean is None:
        return None
    post_bytes = bm25.posting_bytes(q_tok, q_len)
    return hbm_roofline(post_bytes, mean, mn, n, pk,
                        "bm25_ms_kernel (fp32 MaxScore first pass over the fp16-r posting view)", 1,
                        "6 B per posting of every query term (SURVEY.md §8d): sum_q sum_t df(t) * 6",
                        traffic_for("bm25_ms", n_local, Bq))


def oracle_check(args, h, got, q_emb_np, qt_np, ql_np, want_cosine, want_bm25, thr):
    """Full-size comparison with the streamed CPU oracle on a query subset (rank 0; outside every timed region).
    Returns (verified dict, cpu_baseline dict)."""
    import oracle
    from oracle import scale_check
    from optimized_rag_b200 import synthetic as syn
    nv = args.verify_queries if args.verify_queries is not None else (32 if h.world == 1 else 8)
    if nv <= 0:
        return None, None
    cores = os.cpu_count() or 1
    oracle.build()
    oracle.set_threads(cores)
    Bq = args.batch
    rows = scale_check.pick_queries(qt_np, ql_np, Bq, min(nv, Bq)) if qt_np is not None else \
        np.unique(np.linspace(0, Bq - 1, min(nv, Bq)).astype(np.int64))
    t0 = time.perf_counter()
    bm = oracle.StreamedBM25(syn.SEED_TOKENS, args.rows, VOCAB, LMIN, LMAX, thr) if want_bm25 else None
    t_stats = time.perf_counter() - t0
    want, secs = scale_check.reference_lists(args.rows, DIM, q_emb_np[rows] if want_cosine else None,
                                             qt_np[rows] if want_bm25 else None, ql_np[rows] if want_bm25 else None,
                                             TOPK, seed_corpus=syn.SEED_CORPUS, bm25=bm, want_cosine=want_cosine)
    bad = scale_check.compare(got, want, rows)
    if bad:
        raise SystemExit("bench: the CUDA results DIFFER from the CPU oracle at full size: " + "; ".join(bad))
    t_cpu = sum(secs.values())
    verified = {"queries": int(len(rows)), "rows": args.rows, "arrays": sorted(want), "bitwise": True,
                "oracle": "oracle.c streamed restatement (inputs regenerated from the seeds; rag/retrieval.py:362-371, "
                          "324-347, 320; rag/reranker.py:224-271)"}
    cpu = {"value": len(rows) / t_cpu, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": f"{len(rows)} queries of the batch against ALL {args.rows} rows/docs (no extrapolation) through the "
                     f"oracle (C port of the reference arithmetic, OpenMP, queries in SIMD lanes): "
                     + ", ".join(f"{k} {v:.1f} s" for k, v in secs.items())
                     + (f"; BM25 statistics pass {t_stats:.1f} s (BM25Okapi rebuild per call, rag/retrieval.py:338) NOT charged"
                        if want_bm25 else "")}
    return verified, cpu


def run_hybrid_like(args):
    """Configs 3 (hybrid), 2 (cosine only) and 4 (BM25 only): row-sharded corpus, one search call per step."""
    h = Harness(args)
    torch = h.torch
    from optimized_rag_b200 import engine, synthetic as syn
    from optimized_rag_b200.bm25_index import Bm25Index
    from optimized_rag_b200.dist import ShardedBm25, ShardedCosine, ShardedHybrid, shard_range, sharded_plan
    dev, world, rank = h.dev, h.world, h.rank
    N, Bq, k = args.rows, args.batch, TOPK
    want_cos, want_bm = args.config in ("3", "2"), args.config in ("3", "4")
    lo, hi = shard_range(N, rank, world)
    n_local = hi - lo

    # ---- build the shard: embeddings (+ inverse norms, + 16-bit shadow), token corpus, inverted index
    t_setup = time.perf_counter()
    thr = syn.zipf_thresholds(VOCAB)
    cos = bm25 = None
    build_s = None
    if want_cos:
        corpus = engine.gen_embeddings(n_local, DIM, lo, syn.SEED_CORPUS, 0, device=dev)
        cos = engine.CosineIndex(corpus, row_id_base=lo, mode=args.mode)
    if want_bm:
        doc_off, tokens = engine.gen_token_corpus(n_local, lo, syn.SEED_TOKENS, thr, VOCAB, LMIN, LMAX, device=dev)
        torch.cuda.synchronize()
        t_b = time.perf_counter()
        plan, stats = sharded_plan(doc_off, tokens, VOCAB, tile_docs=args.tile_docs)
        bm25 = Bm25Index.from_plan(plan, stats, doc_id_base=lo)
        torch.cuda.synchronize()
        build_s = time.perf_counter() - t_b
        del tokens, plan
        torch.cuda.empty_cache()
    if args.config == "3":
        sh = ShardedHybrid(engine.HybridShard(cos, bm25))
    elif args.config == "2":
        sh = ShardedCosine(cos)
    else:
        sh = ShardedBm25(bm25)
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t_setup

    # ---- queries: generated on the host, staged in pinned memory (replicated on every rank)
    q_emb_np = syn.query_embeddings(Bq, N, DIM) if want_cos else None
    qt_np, ql_np = syn.keyword_queries(Bq, VOCAB, thresholds=thr) if want_bm else (None, None)
    host, devt = [], []
    if want_cos:
        host.append(torch.from_numpy(q_emb_np).pin_memory())
    if want_bm:
        host += [torch.from_numpy(qt_np).pin_memory(), torch.from_numpy(ql_np).pin_memory()]
    devt = [t.to(dev) for t in host]
    out_ids_h = torch.empty((Bq, k), dtype=torch.int64).pin_memory()
    out_sc_h = torch.empty((Bq, k), dtype=torch.float64).pin_memory()
    status_h = torch.empty(Bq, dtype=torch.int32).pin_memory()
    h2d = sum(t.numel() * t.element_size() for t in host)
    d2h = out_ids_h.numel() * 8 + out_sc_h.numel() * 8 + Bq * 4

    # back-to-back batches, inputs resident: nothing synchronises with the host inside the timed region.  Batches are
    # SUBMITTED (two in flight: the latency-bound tail of batch i runs under the scan of batch i+1) and all of them are
    # waited for before the closing event; the per-query overflow flags of every step are kept and checked after it --
    # a flagged step would invalidate the run.
    tickets, last = [], {}

    def step():
        tickets.append(sh.submit(*devt, k))

    def drain():
        res = None
        flags = []
        for t in tickets:
            res = t.wait()          # orders the current stream after that search; no host synchronisation
            flags.append(res["status"])
        tickets.clear()
        last["res"] = res
        return flags

    dev_ms, launches, brackets_overlapped, flags = h.device_timed(step, drain)
    if bool(torch.stack(flags).any()):
        raise SystemExit("bench: a candidate buffer overflowed inside the timed region; results would need the repair path")
    if args.timeline:
        write_timeline(h, step, drain, args.timeline + (f".rank{h.rank}" if world > 1 else ""))
    # (plain steps run the BM25 first pass BEFORE the scan, both with the whole GPU: with co-scheduling it is sized to
    # hide behind the scan on leftover SM resources, and its own duration says nothing about the kernel)
    shard_obj = getattr(sh, "shard", None)
    cosched = bool(shard_obj is not None and shard_obj.coschedule)
    if shard_obj is not None:
        shard_obj.coschedule, shard_obj.serial = False, True
    brackets = h.kernel_brackets(lambda: sh.search(*devt, k, check_overflow=False))
    if shard_obj is not None:
        shard_obj.coschedule, shard_obj.serial = cosched, False
    if os.environ.get("ORAG_BENCH_DEBUG") and rank == 0:
        print("plain-step brackets (ms): scan", [round(x, 3) for x in brackets[0]], "bm25", [round(x, 3) for x in brackets[1]],
              file=sys.stderr, flush=True)
    # per-rank view of the two dominant kernels (mean launch duration: plain steps / inside the submitted loop): the
    # sharded step runs at the pace of the slowest GPU of the box
    per_rank = None
    if world > 1:
        groups_ = (Bq + 255) // 256
        mine = [bracket_stats(brackets[0], groups_)[0] or 0.0, bracket_stats(brackets[1], 1)[0] or 0.0,
                bracket_stats(brackets_overlapped[0], groups_)[0] or 0.0, bracket_stats(brackets_overlapped[1], 1)[0] or 0.0]
        allr = torch.zeros(world * 4, dtype=torch.float64, device=dev)
        h.dist.all_gather_into_tensor(allr, torch.tensor(mine, dtype=torch.float64, device=dev))
        allr = allr.view(world, 4).cpu().tolist()
        per_rank = {"scan_ms": [round(r[0], 4) for r in allr], "bm25_ms": [round(r[1], 4) for r in allr],
                    "scan_ms_in_timed_loop": [round(r[2], 4) for r in allr],
                    "bm25_ms_in_timed_loop": [round(r[3], 4) for r in allr]}

    # ---- timed: end to end through the public call with HOST buffers, `depth` batches between submission and read-back:
    # while batch i is submitted, the results of batch i - (depth - 1) travel to the host and are read there (ONE host
    # synchronisation per step, on a batch that has had depth - 1 steps to finish)
    import collections
    # (four times the lanes: the host only blocks on a batch that is `depth - 1` submissions old.  With depth = lanes the
    # submission of batch i waited for the completion of batch i - 2, tail included, and the GPU ran with one batch less
    # in flight than the device-timed loop.  Measured at 1.25M rows per GPU on two GPUs, 60 steps, ms per batch, device
    # loop 1.23: depth 3 1.30, 6 1.32, 12 1.25.  ORAG_BENCH_E2E_DEPTH overrides.)
    depth = int(os.environ.get("ORAG_BENCH_E2E_DEPTH") or 4 * max(1, int(getattr(sh, "lanes", 1))))
    dev_sets = [devt] + [[t.clone() for t in devt] for _ in range(depth - 1)]
    outs_h = [(out_ids_h, out_sc_h, status_h)] + [
        (torch.empty_like(out_ids_h).pin_memory(), torch.empty_like(out_sc_h).pin_memory(),
         torch.empty_like(status_h).pin_memory()) for _ in range(depth - 1)]
    state = {"i": 0}
    pending = collections.deque()

    def finish(ticket, slot):
        r = ticket.wait()
        ids_h, sc_h, st_h = outs_h[slot]
        ids_h.copy_(r["ids"], non_blocking=True)
        sc_h.copy_(r["scores"], non_blocking=True)
        st_h.copy_(r["status"], non_blocking=True)   # the overflow flags travel with the result
        torch.cuda.current_stream().synchronize()      # (the batches submitted after this one keep running on their lanes)
        if int(st_h.max()) != 0:                       # rare: repair through the exhaustive kernels
            torch.cuda.synchronize()
            r = sh.search(*dev_sets[slot], k, check_overflow=True)
            ids_h.copy_(r["ids"]); sc_h.copy_(r["scores"])

    host_s = {"enqueue": 0.0, "finish": 0.0}

    def step_e2e():
        i = state["i"]
        slot = i % depth
        t0 = time.perf_counter()
        for dst, src in zip(dev_sets[slot], host):
            dst.copy_(src, non_blocking=True)
        pending.append((sh.submit(*dev_sets[slot], k), slot))
        t1 = time.perf_counter()
        if len(pending) >= depth:
            finish(*pending.popleft())
        host_s["enqueue"] += t1 - t0
        host_s["finish"] += time.perf_counter() - t1
        state["i"] = i + 1

    def drain_e2e():
        while pending:
            finish(*pending.popleft())

    e2e_ms = h.e2e_timed(step_e2e, drain_e2e)
    exchange_used = "none (one shard)" if world == 1 else sh.exchange + (f" ({sh.exchange_note})" if sh.exchange_note else "")

    # ---- outside the timed regions: the last step's results against the CPU oracle at full corpus size
    verified = cpu = None
    res = last["res"]   # the results of the last submitted batch of the device-timed loop
    if rank == 0:
        keys = {"3": ["cos_ids", "cos_scores", "bm25_ids", "bm25_scores", "bm25_max", "ids", "rrf_scores"],
                "2": ["cos_ids", "cos_scores"], "4": ["bm25_ids", "bm25_scores", "bm25_max"]}[args.config]
        got = {key: res[key].cpu().numpy() for key in keys}
        verified, cpu = oracle_check(args, h, got, q_emb_np, qt_np, ql_np, want_cos, want_bm, thr)
    h.barrier()

    pk = peaks()
    roof_scan = scan_rooflines(args, n_local, brackets, pk) if want_cos else None
    roof_bm = bm25_roofline(bm25, devt[-2], devt[-1], brackets, pk, n_local, Bq) if want_bm else None
    groups = (Bq + 255) // 256
    for roof, slot, per in ((roof_scan, 0, groups), (roof_bm, 1, 1)):
        if roof is not None:
            m = bracket_stats(brackets_overlapped[slot], per)[0]
            roof["launch_ms_in_timed_loop"] = None if m is None else m / per
    value = Bq * args.steps / (dev_ms * 1e-3)
    line = {"metric": args.metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": h.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, {
                "first_pass": (args.mode + " tensor cores" if want_cos else "") + (" / " if args.config == "3" else "")
                              + ("fp32 MaxScore over fp16 impacts" if want_bm else ""),
                "arithmetic": "results in the reference's float64 arithmetic (bit-exact); candidate generation in low "
                              "precision with proven error margins, then exact re-score",
                "rows_per_gpu": n_local, "parallelism": f"row-sharded x{world}", "setup_s": round(t_setup, 1),
                "exchange": exchange_used, "batches_in_flight": getattr(sh, "lanes", 1),
                "bm25_co_scheduled_with_scan": cosched if args.config == "3" else None,
                **({"bm25_postings_local": bm25.n_postings, "bm25_index_build_s": round(build_s, 2)} if want_bm else {})}),
            "e2e": {"value": Bq * args.steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "host_enqueue_ms_per_step": host_s["enqueue"] * 1e3 / max(state["i"], 1),
                    "host_wait_ms_per_step": host_s["finish"] * 1e3 / max(state["i"], 1), "batches_in_flight": depth,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": launches, "clocks": h.clocks,
            "roofline": roof_scan if want_cos else roof_bm}
    if args.config == "3":
        line["roofline_bm25"] = roof_bm
    if want_cos:
        line["hbm_roofline_queries_per_sec_fp32_corpus"] = Bq / (n_local * DIM * 4 / (pk["hbm_gbs"] * 1e9))
    line["verified_against_oracle"] = verified
    if per_rank is not None:
        line["per_rank_kernel_ms"] = per_rank
    if cpu is not None and world == 1:
        line["cpu_baseline"] = cpu
    h.finish(line, closer=sh.close)


def pairwise_claims(m: int, dev):
    """Config-5 claims on the device: synthetic rows, every 64th one a noisy copy of another claim (cosine ~0.85-0.97
    with its source) -- the same construction as `pairwise_claims_host`."""
    import torch
    from optimized_rag_b200 import engine, synthetic as syn
    emb = engine.gen_embeddings(m, DIM, 0, syn.SEED_CORPUS, 0, device=dev)
    dst = torch.arange(0, m, 64, device=dev)
    src = (dst * 7919 + 13) % m
    w = 0.25 + 0.4 * ((dst * 2654435761 % 1000).to(torch.float32) / 1000)
    emb[dst] = emb[src] + w[:, None] * emb[dst]
    return emb


def run_pairwise(args):
    """Config 5: one step = every claim pair of the matrix through the tensor-core first pass + float64 re-score."""
    h = Harness(args)
    if h.world > 1:
        raise SystemExit("bench --config pairwise: BASELINE config 5 is a one-GPU configuration")
    torch = h.torch
    from optimized_rag_b200 import engine
    dev, M = h.dev, args.rows
    emb = pairwise_claims(M, dev)
    doc = (torch.arange(M, device=dev) // 16).to(torch.int32)
    checker = engine.PairwiseIndex(emb, doc)
    emb_h = emb.cpu().pin_memory()
    last = {}

    def step():
        last["res"] = checker.pairs(0.85, sync=False)

    dev_ms, launches, brackets = h.device_timed(step)

    def step_e2e():
        emb.copy_(emb_h, non_blocking=True)
        chk = engine.PairwiseIndex(emb, doc)       # ingest (fp16 shadow, norms) is part of the end-to-end call
        i, j, s = chk.pairs(0.85, sync=True)
        last["host"] = (i.cpu(), j.cpu(), s.cpu())

    e2e_ms = h.e2e_timed(step_e2e)
    gi, gj, gs = checker.pairs(0.85, sync=True)
    n_pairs = int(gi.numel())
    # oracle on a sub-block (O(M^2 D) float64 on the CPU): the pairs among the first `sub` claims must be identical
    verified = cpu = None
    sub = min(M, 8192)
    nv = args.verify_queries
    if nv is None or nv > 0:
        import oracle
        cores = os.cpu_count() or 1
        oracle.build()
        oracle.set_threads(cores)
        t0 = time.perf_counter()
        oi, oj, osim = oracle.pairwise_candidates_parallel(emb[:sub].cpu().numpy(), doc[:sub].cpu().numpy(), 0.85)
        t_cpu = time.perf_counter() - t0
        keep = (gi < sub) & (gj < sub)
        a = (gi[keep].cpu().numpy(), gj[keep].cpu().numpy(), gs[keep].cpu().numpy())
        same = (np.array_equal(a[0], oi) and np.array_equal(a[1], oj)
                and np.array_equal(a[2].view(np.uint64), osim.view(np.uint64)))
        if not same:
            raise SystemExit("bench: pair set differs from the CPU oracle on the sub-block")
        verified = {"claims": sub, "pairs_in_sub_block": int(len(oi)), "bitwise": True,
                    "oracle": "oracle.c pairwise candidates (rag/consistency_checker.py:169-189, 241-261)"}
        cpu = {"value": (sub * (sub - 1) / 2) / t_cpu, "unit": "pairs/s", "cores": cores, "kind": "port",
               "sample": f"all pairs of the first {sub} claims ({t_cpu:.1f} s); O(M^2 D): the rate does not depend on M"}
    pk = peaks()
    mean, mn, n = bracket_stats(brackets[0], 1)
    blocks = (M + 255) // 256
    flops_required = float(M) * (M - 1) * DIM
    flops_executed = 2.0 * 128 * 256 * DIM * blocks * (blocks + 1)
    roof = tensor_roofline(flops_executed, mean, mn, n, pk,
                           f"cosine_scan_kernel<f16, 2-SM> in pair mode ({blocks * (blocks + 1)} tiles of 128 x 256)", 1,
                           None, {"flops_required": flops_required, "flops_executed": flops_executed,
                                  "flops": "executed = every 128x256 tile at or below the diagonal blocks; required = "
                                           "M(M-1)D for the strict upper triangle (SURVEY.md §8d)",
                                  "frac_required_flops": flops_required / (mean * 1e-3) / 1e12 / pk["tf_sustained"],
                                  "timing": "CUDA events on the launching stream around every launch of the timed loop "
                                            "(mean; min alongside)"}) \
        if mean else None
    pairs = M * (M - 1) / 2
    line = {"metric": args.metric, "value": pairs * args.steps / (dev_ms * 1e-3), "unit": "pairs/s", "n_gpus": 1,
            "steps": args.steps, "warmup": h.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, {"first_pass": "fp16 shadow (rows scaled by powers of two), tcgen05 kind::f16 "
                                                           "cta_group::2, fixed threshold 0.85 - eps",
                                             "pairs_found": n_pairs}),
            "e2e": {"value": pairs * args.steps / (e2e_ms * 1e-3), "unit": "pairs/s",
                    "h2d_bytes_per_step": emb_h.numel() * 4,
                    "d2h_bytes_per_step": int(sum(t.numel() * t.element_size() for t in last["host"])),
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": launches, "clocks": h.clocks, "roofline": roof, "verified_against_oracle": verified}
    if cpu is not None:
        line["cpu_baseline"] = cpu
    h.finish(line)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "pairwise":
        run_pairwise(args)
    else:
        run_hybrid_like(args)


if __name__ == "__main__":
    main()
