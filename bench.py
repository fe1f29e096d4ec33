#!/usr/bin/env python
"""bench.py — DLSA rows/sec on B200 (BASELINE.json metric), one JSON line on stdout.

A "step" is one whole pass of the hot path over a synthetic workload: every Newton pass (fused K1 kernel + batched
Cholesky update; the last pass's H is Sig_inv at the returned coefficients), the one-round combine (+ ONE NCCL allreduce
when N > 1) and the LARS path with AIC/BIC selection (projects/logistic_dlsa.py:293-345).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c1|c2|c3|c3_cat|c4|c5]

--config names the BASELINE.json configuration (default c2 = configs[1], the one the metric is quoted on):
  c1      simulated logistic n=1e5, p=20, K=16                                  (configs[0]; partitions sharded over N)
  c2      simulated logistic n=1e8, p=50, K=1024 PER GPU                        (configs[1]; weak scaling)
  c3      airline-shaped logistic, dense one-hot p=182, 1.5e7 rows PER GPU      (configs[2] at 15 partitions per GPU)
  c3_cat  airline-shaped logistic n=1.2e8, K=120, compact categorical design    (configs[2]; whole data set, strong scaling)
  c4      OLS n=1e9, p=100, K=1024, generated and consumed in chunks            (configs[3]; strong scaling)
  c5      logistic p=500, K=8192 partitions of 2500 rows, LARS + BIC            (configs[4]; strong scaling)
The default run (c2) also carries an ``other_configs`` block with a short measurement of c3_cat, c4 and c5, so that
every BASELINE configuration has a driver-run number at every N.
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

METRIC = "dlsa_rows_per_sec"
UNIT = "rows/s"
FP64_PEAK_FALLBACK_TFLOPS = 37.1  # profiles/r01_microbench_fp64.json (DMMA.8x8x4 == DFMA peak, measured on this pool)

# the SAME workload string on the GPU arm and on the reference arm (the driver compares them); sampling -> config.sample
WORKLOADS = {
    "c1": "c1: simulated logistic n=1e5, p=20, 16 partitions (BASELINE configs[0])",
    "c2": "c2: simulated logistic n=1e8 per GPU, p=50, 1024 partitions per GPU (BASELINE configs[1])",
    "c3": "c3: airline-shaped logistic, dense one-hot p=182 (181 + intercept), 1.5e7 rows / 15 partitions per GPU "
          "(BASELINE configs[2])",
    "c3_cat": "c3_cat: airline-shaped logistic n=1.2e8, p=182, 120 partitions, compact categorical design "
              "(BASELINE configs[2])",
    "c4": "c4: simulated linear (OLS) DLSA n=1e9, p=100, 1024 partitions (BASELINE configs[3])",
    "c5": "c5: logistic p=500 sparse truth, 8192 partitions of 2500 rows, LARS + BIC (BASELINE configs[4])",
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--rows", type=int, default=None, help="override the row count (c1/c2: per GPU)")
    ap.add_argument("--p", type=int, default=None)
    ap.add_argument("--parts", type=int, default=None, help="override the partition count (c1/c2: per GPU)")
    ap.add_argument("--seed", type=int, default=20210517)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip the other_configs block of the default run")
    ap.add_argument("--cpu-parts", type=int, default=None, help="partitions in the bounded CPU sample")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        hi = sorted(sm)[len(sm) // 2:]  # samples under load = upper half
        return {"sm_mhz": sorted(hi)[len(hi) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "power_w_max": max(power), "samples": len(sm)}


_REAL_STDOUT = None


def quiet_stdout():
    """Library banners (e.g. "NCCL version ...") go to fd 1; keep stdout for the ONE JSON line."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(obj):
    _REAL_STDOUT.write(json.dumps(obj) + "\n")
    _REAL_STDOUT.flush()


# =====================================================================================================================
# reference arm: the reference's own CPU implementation on the box's host cores (oracle/ref_cpu.py)
# =====================================================================================================================
def cpu_sample_spec(cfg, args):
    """The bounded CPU sample of a configuration: (shape, K_sample, n_k, p, text)."""
    cores = os.cpu_count() or 1
    if cfg in ("c1", "c2"):
        rows = args.rows or (100_000 if cfg == "c1" else 100_000_000)
        parts = args.parts or (16 if cfg == "c1" else 1024)
        p = args.p or (20 if cfg == "c1" else 50)
        K = args.cpu_parts or min(parts, 64)
        return "sim", min(K, parts), rows // parts, p, f"{min(K, parts)} of the {parts} partitions ({rows // parts} rows x p={p} each)"
    if cfg in ("c3", "c3_cat"):
        K = args.cpu_parts or cores
        n_k = 200_000
        return "airline", K, n_k, 182, (f"{K} partitions of {n_k} rows (a fifth of a 1e6-row partition each), dummy branch of "
                                        "logistic_model (pd.get_dummies, 181 columns + intercept)")
    if cfg == "c4":
        K = args.cpu_parts or 2 * cores
        return "ols", K, 976_562, 100, f"{K} of the 1024 partitions (976562 rows x p=100 each)"
    K = args.cpu_parts or 4 * cores
    return "sim", K, 2500, 500, f"{K} of the 8192 partitions (2500 rows x p=500 each)"


def run_cpu_sample(cfg, args, reps, warmup):
    from oracle import ref_cpu
    shape, K, n_k, p, text = cpu_sample_spec(cfg, args)
    if shape == "ols":
        res = ref_cpu.run_ols_cpu(K=K, n_k=n_k, p=p, seed=args.seed, reps=reps, warmup=warmup)
        how = ("per-partition OLS by LAPACK lstsq + vcov scaling (the reference's OLS DLSA is R lm(), dlsa/R/wlse_func.R:169-182; "
               "R is absent, so this leg is a NumPy port) + group-sum + lstsq + LARS")
    else:
        res = ref_cpu.run_reference_cpu(K=K, n_k=n_k, p=p, seed=args.seed, reps=reps, warmup=warmup, shape=shape)
        how = ("dlsa.models.logistic_model as shipped (sklearn newton-cg, default tol) + pandas group-sum + lstsq + "
               "lars_lsa" if res["kind"] == "reference" else
               "port of logistic_model (sklearn newton-cg, default tol) + group-sum + lstsq + LARS")
    secs = res["seconds"]
    mean_s = sum(secs) / len(secs)
    sample = (f"{text}; {how}; {res['cores']} processes x 1 BLAS thread; data generation and Spark/JVM/Arrow costs "
              "excluded")
    return res["rows"] / mean_s, mean_s, res["cores"], res["kind"], sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    val, mean_s, cores, kind, sample = run_cpu_sample(args.config, args, args.steps, args.warmup)
    import sklearn
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": mean_s * 1e3, "higher_is_better": True,
           "scaling": "weak" if args.config in ("c2", "c3") else "strong", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic",
           "config": {"workload": WORKLOADS[args.config], "sample": sample, "sklearn": sklearn.__version__},
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    emit(out)


# =====================================================================================================================
# GPU arm
# =====================================================================================================================
class Ctx:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the DLSA hot path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.args = args

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, v):
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())


def fp64_peak_live(ctx):
    """DMMA.8x8x4 peak of THIS device, measured now (dlsa_probe_fp64_dmma, registers only, CUDA events, best of 5)."""
    import ctypes as C
    torch = ctx.torch
    from dlsa_b200 import _lib
    sms = torch.cuda.get_device_properties(ctx.dev).multi_processor_count
    sink = torch.empty(sms * 512, dtype=torch.float64, device=ctx.dev)
    flops = C.c_double(0.0)
    best = 0.0
    st = torch.cuda.current_stream().cuda_stream
    for it in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(_lib.lib.dlsa_probe_fp64_dmma(sms, 4000, sink.data_ptr(), C.byref(flops), st), "dlsa_probe_fp64_dmma")
        e1.record()
        torch.cuda.synchronize()
        if it:
            best = max(best, flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


class Workload:
    """One BASELINE configuration on this rank.  step() runs the whole hot path once and returns (fits, comb, path);
    everything a step launches is inside the caller's CUDA-event bracket except where ``timing`` says otherwise."""
    family = "logistic"
    scaling = "strong"
    timing = "whole step between two CUDA events on the launching stream"
    l2 = "inputs are far larger than the 126 MB L2; no flush needed"
    host_api = None
    step_ms_override = None     # chunked workloads: sum of the consume brackets of the last step

    def free(self):
        for k in list(self.__dict__):
            if k not in ("ctx",):
                self.__dict__.pop(k, None)
        self.ctx.torch.cuda.empty_cache()


class SimLogistic(Workload):
    """c1 / c2 / c5: simulate_logistic law (dlsa/models.py:6-40), dense design, K1 + K2 + K4 + K5 + K6/K7."""

    def __init__(self, ctx, cfg, rows=None, parts=None, p=None):
        from dlsa_b200 import synth
        torch, args = ctx.torch, ctx.args
        self.ctx, self.cfg = ctx, cfg
        if cfg == "c2":     # weak scaling: every rank holds its own shard of the same global stream
            self.n = rows or args.rows or 100_000_000
            self.K = parts or args.parts or 1024
            self.p = p or args.p or 50
            self.scaling = "weak"
            row0, self.n_total, self.K_total = ctx.rank * self.n, self.n * ctx.world, self.K * ctx.world
            sizes = None
        else:               # strong scaling: the K partitions of ONE data set are sharded over the ranks
            n_all = rows or args.rows or (100_000 if cfg == "c1" else 8192 * 2500)
            K_all = parts or args.parts or (16 if cfg == "c1" else 8192)
            self.p = p or args.p or (20 if cfg == "c1" else 500)
            from dlsa_b200 import parallel
            off_all = synth.equal_partition_offsets(n_all, K_all)
            lo, hi = parallel.shard_partitions(K_all, ctx.world, ctx.rank)
            row0, self.n = int(off_all[lo]), int(off_all[hi] - off_all[lo])
            self.K, self.n_total, self.K_total = hi - lo, n_all, K_all
            sizes = off_all[lo:hi + 1] - off_all[lo]
        beta = synth.true_beta(self.p) if cfg != "c5" else None
        if cfg == "c5":     # 5 % sparse truth: the first 25 coefficients are 1
            import numpy as np
            beta = np.zeros(self.p)
            beta[: max(1, self.p // 20)] = 1.0
        self.X = torch.empty(self.n, self.p, dtype=torch.float64, device=ctx.dev)
        self.y = torch.empty(self.n, dtype=torch.float64, device=ctx.dev)
        chunk = max(1, 500_000_000 // self.p)
        for s in range(0, self.n, chunk):
            m = min(chunk, self.n - s)
            synth.simulate_logistic_device(m, self.p, args.seed, ctx.dev, row0=row0 + s, beta=beta,
                                           out=(self.X[s:s + m], self.y[s:s + m]))
        self.off = (synth.equal_partition_offsets(self.n, self.K, ctx.dev) if sizes is None else sizes.to(ctx.dev))
        self.max_rows = int((self.off[1:] - self.off[:-1]).max().item()) if self.K else 0
        self.kernel = "irls_small_kernel (K1 fused Newton pass)" if self.p <= 56 else "irls_tiled_kernel (K1 fused Newton pass)"
        self.bytes_per_row, self.flops_per_row = 8 * (self.p + 1), self.p * self.p + 5 * self.p
        self.rows_local = self.n
        self.host_api = "dlsa_b200.engine.dlsa_fit_host" if cfg in ("c1", "c2") else None
        torch.cuda.synchronize()

    def step(self):
        from dlsa_b200 import engine
        return engine.dlsa_fit(self.X, self.y, self.off, fit_intercept=False, sample_size=float(self.n_total), type="lar",
                               max_part_rows=self.max_rows)

    def h2d_bytes(self):
        return int(self.X.numel() * 8 + self.y.numel() * 8 + self.off.numel() * 8)

    def make_host_step(self, pinned=True):
        from dlsa_b200 import engine
        torch = self.ctx.torch
        Xh = torch.empty(self.n, self.p, dtype=torch.float64, pin_memory=pinned)
        yh = torch.empty(self.n, dtype=torch.float64, pin_memory=pinned)
        Xh.copy_(self.X)
        yh.copy_(self.y)
        offh = self.off.cpu()
        torch.cuda.synchronize()

        def host_step():
            f_, c_, p_ = engine.dlsa_fit_host(Xh, yh, offh, fit_intercept=False, sample_size=float(self.n_total),
                                              type="lar", device=self.ctx.dev, buffers=(self.X, self.y))
            return torch.cat([c_.beta_byOLS, c_.beta_byONESHOT, c_.Sig_inv_sum.view(-1), p_.beta[p_.argmin_BIC],
                              p_.beta[p_.argmin_AIC], p_.BIC, p_.AIC]).cpu()
        return host_step


class AirlineCompact(Workload):
    """c3_cat: the whole airline-shaped data set in its compact form (7 numerics + 5 categorical codes, 84 B/row); the
    fused pass is K1c (dlsa_irls_pass_cat_f64): the (g, H) of the dense pass on the one-hot design, never materialised."""

    def __init__(self, ctx, rows=None, parts=None):
        from dlsa_b200 import parallel, synth
        torch, args = ctx.torch, ctx.args
        self.ctx, self.cfg = ctx, "c3_cat"
        n_all, K_all = rows or 120_000_000, parts or 120
        off_all = synth.equal_partition_offsets(n_all, K_all)
        lo, hi = parallel.shard_partitions(K_all, ctx.world, ctx.rank)
        row0, self.n = int(off_all[lo]), int(off_all[hi] - off_all[lo])
        self.K, self.n_total, self.K_total = hi - lo, n_all, K_all
        self.numeric = torch.empty(self.n, synth.AIRLINE_NUM, dtype=torch.float64, device=ctx.dev)
        self.codes = torch.empty(self.n, len(synth.AIRLINE_LEVELS), dtype=torch.int32, device=ctx.dev)
        self.y = torch.empty(self.n, dtype=torch.float64, device=ctx.dev)
        chunk = 20_000_000
        for s in range(0, self.n, chunk):
            m = min(chunk, self.n - s)
            _, _, self.n_levels, _ = synth.airline_shaped_device(
                m, args.seed, ctx.dev, row0=row0 + s, out=(self.numeric[s:s + m], self.codes[s:s + m], self.y[s:s + m]))
        self.off = (off_all[lo:hi + 1] - off_all[lo]).to(ctx.dev)
        self.max_rows = int((self.off[1:] - self.off[:-1]).max().item()) if self.K else 0
        self.p = 1 + synth.AIRLINE_NUM + sum(self.n_levels)
        self.kernel = "irls_cat_kernel (K1c fused Newton pass on the compact categorical design)"
        self.bytes_per_row = 8 * synth.AIRLINE_NUM + 4 * len(self.n_levels) + 8
        self.flops_per_row = None   # HBM-side kernel: z'z DMMA + scatter-adds; reported against the HBM roof only
        self.rows_local = self.n
        self.host_api = "dlsa_b200.engine.dlsa_fit_host_cat"
        torch.cuda.synchronize()

    def step(self):
        from dlsa_b200 import engine
        fits = engine.fit_logistic_partitions_cat(self.numeric, self.codes, self.n_levels, self.y, self.off,
                                                  fit_intercept=True, defer_check=True, max_part_rows=self.max_rows)
        return engine.finish_fit(fits, True, float(self.n_total), "lar")

    def h2d_bytes(self):
        return int(self.numeric.numel() * 8 + self.codes.numel() * 4 + self.y.numel() * 8 + self.off.numel() * 8)

    def make_host_step(self, pinned=True):
        from dlsa_b200 import engine
        torch = self.ctx.torch
        nh = torch.empty(self.numeric.shape, dtype=torch.float64, pin_memory=pinned)
        ch = torch.empty(self.codes.shape, dtype=torch.int32, pin_memory=pinned)
        yh = torch.empty(self.n, dtype=torch.float64, pin_memory=pinned)
        nh.copy_(self.numeric)
        ch.copy_(self.codes)
        yh.copy_(self.y)
        offh = self.off.cpu()
        torch.cuda.synchronize()

        def host_step():
            f_, c_, p_ = engine.dlsa_fit_host_cat(nh, ch, self.n_levels, yh, offh, fit_intercept=True,
                                                  sample_size=float(self.n_total), device=self.ctx.dev,
                                                  buffers=(self.numeric, self.codes, self.y))
            return torch.cat([c_.beta_byOLS, c_.beta_byONESHOT, c_.Sig_inv_sum.view(-1), p_.beta[p_.argmin_BIC],
                              p_.beta[p_.argmin_AIC], p_.BIC, p_.AIC]).cpu()
        return host_step


class AirlineDense(Workload):
    """c3: the same design expanded to dense one-hot float64 (181 columns + intercept, 1456 B/row): the contraction
    BASELINE.json's north_star names for the airline shape.  15 partitions of 1e6 rows per GPU (weak scaling)."""
    scaling = "weak"

    def __init__(self, ctx, rows=None, parts=None):
        import numpy as np
        from dlsa_b200 import _lib, synth
        torch, args = ctx.torch, ctx.args
        self.ctx, self.cfg = ctx, "c3"
        self.n, self.K = rows or 15_000_000, parts or 15
        self.n_total, self.K_total = self.n * ctx.world, self.K * ctx.world
        row0 = ctx.rank * self.n
        n_levels = [L - 1 for L in synth.AIRLINE_LEVELS]
        px = synth.AIRLINE_NUM + sum(n_levels)
        # odd width: an even row pitch keeps every row 16-byte aligned (TMA row copies instead of 8-byte cp.async)
        self.X = torch.zeros(self.n, px + (px & 1), dtype=torch.float64, device=ctx.dev)[:, :px]
        self.y = torch.empty(self.n, dtype=torch.float64, device=ctx.dev)
        base = torch.as_tensor(np.concatenate([[0], np.cumsum(n_levels)[:-1]]).astype(np.int32) + synth.AIRLINE_NUM,
                               device=ctx.dev)
        chunk = 5_000_000
        for s in range(0, self.n, chunk):
            m = min(chunk, self.n - s)
            num, codes, _, yv = synth.airline_shaped_device(m, args.seed, ctx.dev, row0=row0 + s)
            self.y[s:s + m] = yv
            Xs = self.X[s:s + m]
            _lib.check(_lib.lib.dlsa_onehot_expand_f64(num.data_ptr(), synth.AIRLINE_NUM, None, None, codes.data_ptr(),
                                                       len(n_levels), base.data_ptr(), m, Xs.data_ptr(), Xs.stride(0),
                                                       torch.cuda.current_stream().cuda_stream), "dlsa_onehot_expand_f64")
            torch.cuda.synchronize()
            del num, codes, yv
        self.off = synth.equal_partition_offsets(self.n, self.K, ctx.dev)
        self.max_rows = int((self.off[1:] - self.off[:-1]).max().item())
        self.p = px + 1
        self.kernel = "irls_tiled_kernel (K1 fused Newton pass, CTA-tiled Gram)"
        self.bytes_per_row, self.flops_per_row = 8 * (px + 1), self.p * self.p + 5 * self.p
        self.rows_local = self.n
        torch.cuda.synchronize()

    def step(self):
        from dlsa_b200 import engine
        return engine.dlsa_fit(self.X, self.y, self.off, fit_intercept=True, sample_size=float(self.n_total), type="lar",
                               max_part_rows=self.max_rows)


class OlsChunked(Workload):
    """c4: OLS DLSA, n = 1e9 rows x p = 100 (808 GB of float64): this rank's share of the 1024 partitions is generated on
    the device and consumed chunk by chunk (the shard exceeds HBM below 8 GPUs); only the consume side (fused X'X / X'y /
    y'y pass, batched Cholesky, sigma^2 scaling, partition reduce) and the reduce side (allreduce, solve, LARS) are timed."""
    family = "linear"
    timing = ("sum of the CUDA-event brackets around the consume side of every chunk + the reduce side; the counter-based "
              "generation of each chunk (input set-up, it replaces reading the chunk from storage) is outside the brackets")

    def __init__(self, ctx, rows=None, parts=None, p=None, chunk_parts=32):
        import numpy as np
        from dlsa_b200 import parallel, synth
        torch = ctx.torch
        self.ctx, self.cfg = ctx, "c4"
        self.n_total, self.K_total, self.p = rows or 1_000_000_000, parts or 1024, p or 100
        self.off_all = synth.equal_partition_offsets(self.n_total, self.K_total)
        self.lo, self.hi = parallel.shard_partitions(self.K_total, ctx.world, ctx.rank)
        self.K = self.hi - self.lo
        self.n = int(self.off_all[self.hi] - self.off_all[self.lo])
        self.chunk_parts = chunk_parts
        cmax = int(max(self.off_all[min(k + chunk_parts, self.hi)] - self.off_all[k]
                       for k in range(self.lo, self.hi, chunk_parts))) if self.K else 1
        self.X = torch.empty(cmax, self.p, dtype=torch.float64, device=ctx.dev)
        self.y = torch.empty(cmax, dtype=torch.float64, device=ctx.dev)
        self.beta = np.tile([3, 1.5, 0, 0, 2, 0, 0, 0, 0, 0], (self.p + 9) // 10)[: self.p].astype(np.float64)  # dlsa/R/main.R:26
        self.kernel = "irls_tiled_kernel<LINEAR> (K8 fused X'X / X'y / y'y pass)"
        self.bytes_per_row, self.flops_per_row = 8 * (self.p + 1), self.p * (self.p + 1) + 2 * self.p + 2
        self.rows_local = self.n
        self.chunks_per_step = (self.K + chunk_parts - 1) // chunk_parts
        self.pass_ms = []

    def step(self):
        from dlsa_b200 import engine, synth
        torch, ctx = self.ctx.torch, self.ctx
        packed = None
        consume_ms = []
        last = None
        for k0 in range(self.lo, self.hi, self.chunk_parts):
            k1 = min(k0 + self.chunk_parts, self.hi)
            r0, r1 = int(self.off_all[k0]), int(self.off_all[k1])
            m = r1 - r0
            synth.simulate_logistic_device(m, self.p, ctx.args.seed, ctx.dev, row0=r0, beta=self.beta, law=1,
                                           out=(self.X[:m], self.y[:m]))
            off = (self.off_all[k0:k1 + 1] - r0).to(ctx.dev)
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            fits = engine.fit_linear_partitions(self.X[:m], self.y[:m], off, events=(e1,))
            part = engine.reduce_partitions(fits)
            packed = part if packed is None else packed + part
            e2.record()
            consume_ms.append((e0, e1, e2, m))
            last = fits
        e3, e4 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e3.record()
        out = engine.finish_fit(last, False, float(self.n_total), "lar", packed=packed)
        e4.record()
        torch.cuda.synchronize()
        self.pass_ms = [a.elapsed_time(b) for a, b, _, _ in consume_ms]
        self.step_ms_override = sum(a.elapsed_time(c) for a, _, c, _ in consume_ms) + e3.elapsed_time(e4)
        return out


def make_workload(ctx, cfg):
    sc = float(os.environ.get("DLSA_BENCH_SCALE", "1"))   # debugging only: shrink every workload (never set by the driver)
    if sc != 1.0:
        if cfg == "c2":
            return SimLogistic(ctx, cfg, rows=int(100_000_000 * sc), parts=max(64, int(1024 * sc)))
        if cfg == "c5":
            return SimLogistic(ctx, cfg, rows=max(8, int(8192 * sc)) * 2500, parts=max(8, int(8192 * sc)))
        if cfg == "c3_cat":
            return AirlineCompact(ctx, rows=int(120_000_000 * sc), parts=120)
        if cfg == "c3":
            return AirlineDense(ctx, rows=int(15_000_000 * sc), parts=15)
        if cfg == "c4":
            return OlsChunked(ctx, rows=int(1_000_000_000 * sc), parts=1024)
    if cfg in ("c1", "c2", "c5"):
        return SimLogistic(ctx, cfg)
    if cfg == "c3_cat":
        return AirlineCompact(ctx)
    if cfg == "c3":
        return AirlineDense(ctx)
    return OlsChunked(ctx)


def measure(ctx, wl, steps, warmup, sample_clocks=False):
    """W untimed steps, then K timed steps between barrier + synchronize, CUDA events, max over ranks."""
    torch = ctx.torch
    from dlsa_b200 import _lib, engine
    for _ in range(warmup):
        out = wl.step()
    ctx.barrier()
    sampler = ClockSampler(ctx.local) if (sample_clocks and ctx.rank == 0) else None
    if sampler:
        sampler.start()
    engine.PASS_EVENTS = []
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_ms = {"map": 0.0, "reduce_side": 0.0}
    ctx.barrier()
    e0.record()
    override = 0.0
    for _ in range(steps):
        out = wl.step()
        if wl.step_ms_override is not None:
            override += wl.step_ms_override
    e1.record()
    ctx.barrier()
    ms = e0.elapsed_time(e1) if wl.step_ms_override is None else override
    launches = _lib.launch_count() - launches0
    pass_events = engine.PASS_EVENTS
    engine.PASS_EVENTS = None
    clocks = sampler.stop() if sampler else None
    ms = ctx.max_over_ranks(ms)
    full = [a.elapsed_time(b) for a, b, is_full in pass_events if is_full]
    allp = [a.elapsed_time(b) for a, b, _ in pass_events]
    fits, comb, path = out
    return {"ms_per_step": ms / steps, "wall_ms_per_step": e0.elapsed_time(e1) / steps, "launches": int(launches),
            "pass_ms": (sum(full) / len(full)) if full else None, "passes_full": len(full) / steps,
            "kernel_ms_total": sum(allp), "clocks": clocks, "fits": fits, "comb": comb, "path": path,
            "stage_ms": stage_ms}


def roofline_of(wl, m, steps, f64_peak, f64_src):
    hbm_peak, hbm_src = measured_peaks()
    pass_ms = m["pass_ms"]
    if not pass_ms:
        return None
    n = wl.rows_local if wl.cfg != "c4" else wl.rows_local / max(wl.chunks_per_step, 1)
    ach_gbs = n * wl.bytes_per_row / (pass_ms * 1e-3) / 1e9
    hbm_view = {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                "peak_source": hbm_src}
    out = {"kernel": wl.kernel, "kernel_ms": pass_ms, "rows_per_launch": n,
           "kernel_share_of_step": m["kernel_ms_total"] / (m["wall_ms_per_step"] * steps),
           "algorithmic_bytes_per_row": wl.bytes_per_row, "algorithmic_flops_per_row": wl.flops_per_row, "traffic": None}
    if wl.flops_per_row:
        ach_tf = n * wl.flops_per_row / (pass_ms * 1e-3) / 1e12
        f64_view = {"bound": "tensor", "achieved": ach_tf, "peak": f64_peak, "unit": "TFLOP/s", "frac": ach_tf / f64_peak,
                    "peak_source": f64_src,
                    "note": "FP64 DMMA.8x8x4 and DFMA share one 64 FMA/clk/SM pipe on B200; tcgen05 has no f64 kind"}
        t_hbm = n * wl.bytes_per_row / (hbm_peak * 1e9)
        t_f64 = n * wl.flops_per_row / (f64_peak * 1e12)
        binding = f64_view if t_f64 > t_hbm else hbm_view
        out.update(binding)
        out.update({"time_bound_ms": {"hbm": t_hbm * 1e3, "fp64": t_f64 * 1e3},
                    "binding_roof": "fp64" if t_f64 > t_hbm else "hbm", "hbm": hbm_view, "fp64": f64_view})
    else:
        out.update(hbm_view)
        out.update({"binding_roof": "hbm", "hbm": hbm_view})
    # DRAM bytes of one launch, from the committed ncu --set full capture of the same kernel (profiles/)
    traffic_path = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.isfile(traffic_path):
        try:
            with open(traffic_path) as f:
                tj = json.load(f)
            for ent in (tj if isinstance(tj, list) else [tj]):
                if ent.get("kernel_tag", "small") in wl.kernel and int(ent.get("p", -1)) == wl.p:
                    out["traffic"] = ent["dram_bytes_per_launch"] * (n / float(ent["rows"]))
                    out["traffic_source"] = ent.get("source", "profiles/k1_traffic.json (ncu dram__bytes_read.sum + "
                                                              "dram__bytes_write.sum, scaled by rows)")
        except Exception:
            pass
    return out


def result_check(wl, m):
    comb, path = m["comb"], m["path"]
    sup = path.beta[path.argmin_BIC] != 0
    return {"K_eff": comb.K_eff, "K_total": comb.K_total, "path_rows": int(path.beta.shape[0]),
            "support_BIC": int(sup.sum().item()), "argmin_BIC": int(path.argmin_BIC),
            "theta_head": [round(v, 6) for v in comb.beta_byOLS[:3].tolist()],
            "all_status_ok": bool(int(m["fits"].n_bad.item()) == 0) if m["fits"].n_bad is not None else None}


def combine_check(ctx, wl, m):
    """Hardware proof of the one collective (N > 1): every rank's local packed buffer is all-gathered, summed in rank
    order on the host in float64 and compared with what ncclAllReduce produced; outside the timed region."""
    if ctx.world == 1:
        return None
    torch, dist = ctx.torch, ctx.dist
    from dlsa_b200 import engine
    fits = m["fits"]
    if wl.cfg == "c4":
        return {"note": "chunked map: local buffers are accumulated per chunk; see c2 / c3_cat / c5"}
    local = engine.reduce_partitions(fits)
    gathered = [torch.empty_like(local) for _ in range(ctx.world)]
    dist.all_gather(gathered, local)
    red = local.clone()
    dist.all_reduce(red)
    host = sum(g.cpu().double() for g in gathered[1:]) + gathered[0].cpu().double()
    diff = (red.cpu() - host).abs().max().item()
    scale = host.abs().max().item()
    S = m["comb"].Sig_inv_sum.reshape(-1).cpu()
    p = fits.coef.shape[1]
    return {"ranks": ctx.world, "payload_doubles": int(local.numel()), "max_abs_diff": diff,
            "max_rel_diff": diff / max(scale, 1e-300), "ok": bool(diff <= 1e-12 * max(scale, 1.0)),
            "matches_step_result": bool((S - host[: p * p]).abs().max().item() <= 1e-12 * max(scale, 1.0))}


def measure_e2e(ctx, wl, steps, pinned=True):
    """Same metric through the host-buffer API: host->device copies and the device->host read of the result inside."""
    torch = ctx.torch
    try:
        ctx.barrier()
        mem_ok, mem_msg = 1, ""
        try:
            import psutil
            need = ctx.world * wl.h2d_bytes() * 1.15
            avail = psutil.virtual_memory().available
            if avail < need:
                mem_ok, mem_msg = 0, f"host memory: {avail / 1e9:.0f} GB available, {need / 1e9:.0f} GB needed for {ctx.world} shards"
        except ImportError:
            pass
        flag = torch.tensor([mem_ok], dtype=torch.int32, device=ctx.dev)
        if ctx.world > 1:
            ctx.dist.all_reduce(flag, op=ctx.dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            raise MemoryError(mem_msg or "another rank reported too little host memory")
        host_step = wl.make_host_step(pinned=pinned)
        host_step()
        ctx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        nsteps = max(1, min(steps, 3))
        for _ in range(nsteps):
            res = host_step()
        e1.record()
        ctx.barrier()
        ems = ctx.max_over_ranks(e0.elapsed_time(e1)) / nsteps
        return {"value": wl.n_total / (ems * 1e-3), "unit": UNIT, "h2d_bytes_per_step": wl.h2d_bytes(),
                "d2h_bytes_per_step": int(res.numel() * 8), "ms_per_step": ems, "steps": nsteps,
                "host_memory": "pinned" if pinned else "pageable",
                "api": wl.host_api + ": host X / y / offsets in, partition blocks whose H2D copies overlap the fits of "
                       "earlier blocks, combined estimates + LARS path read back to the host"}
    except Exception as ex:  # e.g. not enough host memory to pin the shard
        return {"value": None, "unit": UNIT, "error": repr(ex)[:200]}


def short_config_line(ctx, cfg, f64_peak, f64_src, steps=2, warmup=2):
    """other_configs entry: a short measurement of another BASELINE configuration on the same box."""
    try:
        wl = make_workload(ctx, cfg)
        m = measure(ctx, wl, steps, warmup)
        rl = roofline_of(wl, m, steps, f64_peak, f64_src)
        fits = m["fits"]
        line = {"workload": WORKLOADS[cfg], "value": wl.n_total / (m["ms_per_step"] * 1e-3), "unit": UNIT,
                "ms_per_step": m["ms_per_step"], "scaling": wl.scaling, "rows_total": int(wl.n_total),
                "partitions_total": int(wl.K_total), "rows_this_rank": int(wl.rows_local), "steps": steps, "warmup": warmup,
                "passes": fits.passes, "pilot_passes": fits.pilot_passes, "timing": wl.timing,
                "gpu_launches": m["launches"], "dominant_kernel": wl.kernel,
                "roofline": None if rl is None else {k: rl[k] for k in ("bound", "achieved", "peak", "unit", "frac",
                                                                        "kernel_ms", "rows_per_launch",
                                                                        "kernel_share_of_step") if k in rl},
                "result_check": result_check(wl, m), "combine_check": combine_check(ctx, wl, m)}
        if cfg == "c5":
            line["stage_ms"] = stage_breakdown(ctx, wl)
        wl.free()
        del wl, m
        ctx.torch.cuda.empty_cache()
        return line
    except Exception as ex:
        ctx.torch.cuda.empty_cache()
        return {"workload": WORKLOADS[cfg], "value": None, "error": repr(ex)[:300]}


def stage_breakdown(ctx, wl):
    """Per-stage device time of one more step (projects/logistic_dlsa.py:293-345 stage boundaries): map = K1 passes + K2
    updates, combine = K4 + allreduce + K5, LARS = K6 + K7."""
    torch = ctx.torch
    from dlsa_b200 import engine
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    engine.PASS_EVENTS = []
    ev[0].record()
    fits = engine.fit_logistic_partitions(wl.X, wl.y, wl.off, False, max_part_rows=wl.max_rows)
    ev[1].record()
    comb = engine.combine_partitions(fits)
    ev[2].record()
    engine.lars_path(comb.Sig_inv_sum, comb.beta_byOLS, False, float(wl.n_total))
    ev[3].record()
    torch.cuda.synchronize()
    k1 = sum(a.elapsed_time(b) for a, b, _ in engine.PASS_EVENTS)
    n_pass = len(engine.PASS_EVENTS)
    engine.PASS_EVENTS = None
    mp = ev[0].elapsed_time(ev[1])
    return {"map_ms": mp, "K1_fused_passes_ms": k1, "K1_launches": n_pass, "K2_and_rest_of_map_ms": mp - k1,
            "K4_allreduce_K5_combine_ms": ev[1].elapsed_time(ev[2]), "K6_K7_lars_ms": ev[2].elapsed_time(ev[3])}


def main():
    args = parse()
    quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
        return
    ctx = Ctx(args)
    torch = ctx.torch
    cfg = args.config
    f64_live = fp64_peak_live(ctx)
    f64_peak, f64_src = ((f64_live, "measured live in this run (dlsa_probe_fp64_dmma: registers-only DMMA.8x8x4 stream, "
                                    "CUDA events, best of 5)") if f64_live > 1.0 else
                         (FP64_PEAK_FALLBACK_TFLOPS, "fallback (profiles/r01_microbench_fp64.json)"))

    wl = make_workload(ctx, cfg)
    m = measure(ctx, wl, args.steps, args.warmup, sample_clocks=True)
    value = wl.n_total / (m["ms_per_step"] * 1e-3)
    roofline = roofline_of(wl, m, args.steps, f64_peak, f64_src)
    rc = result_check(wl, m)
    cc = combine_check(ctx, wl, m)
    fits = m["fits"]

    e2e = None
    if not args.no_e2e and wl.host_api is not None:
        e2e = measure_e2e(ctx, wl, args.steps, pinned=True)
        if cfg == "c2" and ctx.world == 1 and e2e.get("value"):
            pg = measure_e2e(ctx, wl, 1, pinned=False)   # what a Spark worker hands over: pageable memory
            e2e["pageable"] = {k: pg.get(k) for k in ("value", "ms_per_step", "error") if k in pg}
    elif not args.no_e2e:
        e2e = {"value": None, "unit": UNIT, "note": "no host-buffer entry for this configuration (the data set is generated "
                                                     "on the device in chunks / exceeds host memory)"}

    cpu = None
    if ctx.rank == 0 and ctx.world == 1 and not args.no_cpu:
        try:
            v, mean_s, cores, kind, sample = run_cpu_sample(cfg, args, 1, 0)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
        except Exception as ex:
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": repr(ex)[:200]}

    config = {"workload": WORKLOADS[cfg], "rows_total": int(wl.n_total), "partitions_total": int(wl.K_total),
              "rows_this_rank": int(wl.rows_local), "p": int(wl.p), "family": wl.family,
              "parallelism": f"partitions x{ctx.world}", "newton_passes_per_step": fits.passes, "tol": 1e-10,
              "warm_start": {"pilot_partitions": fits.pilot_partitions, "pilot_passes": fits.pilot_passes,
                             "cold_restart": fits.cold_restart,
                             "note": "pilot block fitted from beta=0, its mean coefficients start the other partitions "
                                     "(R init.beta); same MLE, same tol; every pass is inside the timed region"},
              "timing": wl.timing, "l2": wl.l2, "seed": args.seed}
    stage_ms = stage_breakdown(ctx, wl) if cfg == "c5" else None
    wl.free()
    del wl

    others = None
    if cfg == "c2" and not args.no_others:
        others = {}
        for oc in ("c3_cat", "c4", "c5"):
            others[oc] = short_config_line(ctx, oc, f64_peak, f64_src)

    if ctx.rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ctx.world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": m["ms_per_step"], "higher_is_better": True,
               "scaling": "weak" if cfg in ("c2", "c3") else "strong", "vs_baseline": None, "dtype": "f64",
               "data": "synthetic", "config": config, "clocks": m["clocks"], "e2e": e2e,
               "gpu_launches": m["launches"], "roofline": roofline, "cpu_baseline": cpu, "result_check": rc,
               "fp64_peak_probe": {"tflops_live": f64_live, "tflops_round1_microbench": FP64_PEAK_FALLBACK_TFLOPS,
                                   "nominal": "148 SM x 64 FMA/clk x 2 x 1.965 GHz = 37.2"}}
        if cc is not None:
            out["combine_check"] = cc
        if stage_ms is not None:
            out["stage_ms"] = stage_ms
        if others is not None:
            out["other_configs"] = others
        emit(out)
    if ctx.world > 1:
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
