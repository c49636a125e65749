#!/usr/bin/env python3
"""bench.py -- headline benchmark of the Datok matrix-FSA transduction path on B200.

Metric (BASELINE.json): GB/s of input tokenized + sentence-split, tokenizer_de.matok,
synthetic German corpus of ~10 KB EOT-separated documents,
flags TOKENS|SENTENCES|TOKEN_POS|SENTENCE_POS.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--size BYTES] [--impl reference]

One "step" = one pass of the hot path over one corpus batch of --size bytes per GPU
(weak scaling: every rank owns its own EOT-aligned shard, generated from its own seed).
  value       whole-job GB/s, input resident in HBM, offset arrays left in HBM
              (CUDA events on the library's stream, max over ranks)
  e2e         same through datok_transduce() with HOST buffers: pinned host input ->
              H2D -> kernels -> D2H of the offset arrays
  roofline    (N + 8T + 8S + 8D bytes) / device time of the whole path, vs the measured
              HBM copy peak (MEASURED_PEAKS.json)
  cpu_baseline  the CPU oracle port (one worker per document, all host cores) on a
              bounded sample of the same corpus
--impl reference times that CPU path alone (the Go reference cannot be built here).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLAGS = 1 | 2 | 4 | 8  # TOKENS|SENTENCES|TOKEN_POS|SENTENCE_POS  (= `datok tokenize -p --sentence-positions`)
MODEL = os.path.join(ROOT, "testdata", "tokenizer_de.matok")
METRIC = "GB/s input tokenized+sentence-split (de .matok)"


def measured_traffic(n_bytes):
    """DRAM bytes (read+write) of one step from the committed ncu --set full capture of the same
    workload (profiles/r1_traffic.json, written by scripts/ncu_traffic.py); None if absent"""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        if int(t["input_bytes"]) == int(n_bytes):
            return float(t["dram_bytes_per_step"])
    except Exception:
        pass
    return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """samples nvidia-smi clocks / throttle reasons while the timed region runs"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                if out.returncode == 0 and out.stdout.strip():
                    self.rows.append([x.strip() for x in out.stdout.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]),
                "power_w_max": max(float(r[2]) for r in self.rows), "samples": len(self.rows), "reasons": reasons}


def make_corpus(nbytes, seed, pinned_alloc=None):
    from datok_b200 import corpus
    if pinned_alloc is not None:
        ptr = pinned_alloc(nbytes)
        if not ptr:
            raise RuntimeError("pinned allocation failed")
        buf = (C.c_uint8 * nbytes).from_address(ptr)
        arr = np.frombuffer(buf, dtype=np.uint8)
    else:
        arr = np.empty(nbytes, dtype=np.uint8)
    docs = corpus.generate_blocks_into(corpus.GERMAN, seed, arr, block=64 << 20)
    return arr, docs


def cpu_reference_rate(arr, seconds=15.0, threads=None):
    """the oracle port on host cores, one worker per document, on a bounded sample"""
    from oracle import pyoracle
    threads = threads or os.cpu_count() or 1
    ora = pyoracle.OracleModel(MODEL)
    probe = arr[: min(arr.size, 4 << 20)]
    cut = int(np.flatnonzero(probe == 4)[-1]) + 1
    t0 = time.perf_counter()
    ora.transduce_docs_mt(probe[:cut], FLAGS, threads)
    rate = cut / (time.perf_counter() - t0)
    want = int(min(arr.size, max(cut, rate * seconds)))
    eots = np.flatnonzero(arr[:want] == 4)
    n = int(eots[-1]) + 1
    passes = max(1, int(round(rate * seconds / n)))
    t0 = time.perf_counter()
    for _ in range(passes):
        res = ora.transduce_docs_mt(arr[:n], FLAGS, threads)
    dt = time.perf_counter() - t0
    return {"value": passes * n / dt / 1e9, "unit": "GB/s", "cores": threads, "kind": "port",
            "sample": f"{passes} pass(es) over the first {n} bytes ({res['docs']} documents) of the same corpus, "
                      f"{dt:.1f} s, C restatement of matrix.go:348-698 + token_writer.go (Go toolchain absent)",
            "seconds": dt, "bytes": passes * n, "tokens": res["tokens"]}


def run_reference(args, rank, world):
    if rank != 0:
        return
    size = min(args.size, 256 << 20)
    arr, docs = make_corpus(size, 20261018)
    per = []
    base = None
    for i in range(args.warmup + args.steps):
        base = cpu_reference_rate(arr, seconds=max(2.0, 60.0 / max(1, args.warmup + args.steps)))
        if i >= args.warmup:
            per.append(base)
    tot_b = sum(p["bytes"] for p in per)
    tot_s = sum(p["seconds"] for p in per)
    v = tot_b / tot_s / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "GB/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot_s / len(per) * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "C2: tokenizer_de.matok, synthetic German corpus, ~10 KB EOT-separated documents, "
                                   "flags TOKENS|SENTENCES|TOKEN_POS|SENTENCE_POS; each step a bounded sample"},
            "cpu_baseline": {"value": v, "unit": "GB/s", "cores": base["cores"], "kind": "port",
                             "sample": base["sample"]},
            "e2e": {"value": v, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, default=1 << 30, help="corpus bytes per GPU")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import datok_b200 as d
    from datok_b200 import _lib
    tok = d.LoadTokenizerFile(MODEL, device=local)
    if tok is None:
        raise SystemExit("libdatok_b200.so could not load the model on the GPU (no fallback exists)")
    L = _lib.lib()
    arr, docs = make_corpus(args.size, 20261018 + 1000003 * rank, pinned_alloc=L.datok_host_alloc)
    N = arr.size
    d_in = torch.empty(N, dtype=torch.uint8, device="cuda")
    d_in.copy_(torch.from_numpy(arr))
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident leg: delta-coded token spans, two bytes per value (the one-byte form of the
    # host legs only pays off across PCIe) ----
    DFLAGS = FLAGS | d.COMPACT
    for _ in range(args.warmup):
        tok.transduce_device(d_in.data_ptr(), N, DFLAGS).close()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    dev_ms, ktimes, launches = [], {}, 0
    T = S = D = 0
    for _ in range(args.steps):
        r = tok.transduce_device(d_in.data_ptr(), N, DFLAGS)
        dev_ms.append(r.ms_kernels)
        T, S, D = r.n_tokens, r.n_sentences, r.n_texts
        for k, v in tok.kernel_times().items():
            ktimes[k] = ktimes.get(k, 0.0) + v
        launches += tok.launch_count()
        r.close()
    barrier()
    wall_dev = time.perf_counter() - t0
    clocks = sampler.summary()

    # ---- end-to-end legs (host buffers through the C ABI) ----------------------
    # e2e: what the Tokenizer shim calls -- DATOK_COMPACT8, 4-byte delta-coded token spans that the
    # host formatter / replay decode while they walk the tokens anyway.  e2e_absolute: the same call
    # returning absolute (byte, rune) offset pairs, 16 bytes per token.
    def e2e_leg(flags):
        for _ in range(min(2, args.warmup)):
            tok.transduce_arrays(arr, flags).close()
        barrier()
        t0 = time.perf_counter()
        steps = max(1, min(args.steps, 3))
        out_bytes = 0
        for _ in range(steps):
            r = tok.transduce_arrays(arr, flags)
            per_tok = 4 if r.tok_delta8 is not None else 8 if r.tok_delta is not None else 16
            out_bytes = per_tok * r.n_tokens + 4 * (r.n_sent_pos + r.n_sentences + 4 * r.n_texts)
            r.close()
        barrier()
        return (time.perf_counter() - t0) / steps, out_bytes

    wall_abs, d2h_abs = e2e_leg(FLAGS)
    wall_e2e, d2h = e2e_leg(FLAGS | d.COMPACT8)
    h2d = N

    # ---- the same call followed by the host half of the TokenWriter: the exact text
    # NewTokenWriter(w, flags) writes (datok_format, all host cores) ----
    import ctypes as C
    import numpy as np
    # (single-GPU runs only: the text is ~2.7x the input and every rank would hold its own copy)
    fmt_bytes, wall_fmt = 0, float("nan")
    if world == 1:
        fmt_buf = np.empty(4 * N + (1 << 20), dtype=np.uint8)

        def fmt_once():
            r = tok.transduce_arrays(arr, FLAGS | d.COMPACT8)
            need = L.datok_format(r._h, arr.ctypes.data, N, FLAGS, fmt_buf.ctypes.data, fmt_buf.size)
            r.close()
            return int(need)

        fmt_once()
        barrier()
        t0 = time.perf_counter()
        fmt_steps = max(1, min(args.steps, 2))
        for _ in range(fmt_steps):
            fmt_bytes = fmt_once()
        barrier()
        wall_fmt = (time.perf_counter() - t0) / fmt_steps
        del fmt_buf

    # ---- reduce over ranks (max time; counts summed via the per-shard count exchange) -
    ms_step = sum(dev_ms) / len(dev_ms)
    stats = torch.tensor([ms_step, wall_dev / args.steps * 1e3, wall_e2e * 1e3, wall_abs * 1e3, wall_fmt * 1e3], dtype=torch.float64,
                         device="cuda")
    counts = torch.tensor([N, T, S, D], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        # the path's only exchange: per-shard counts -> global offset bases
        gathered = [torch.zeros_like(counts) for _ in range(world)]
        dist.all_gather(gathered, counts)
        allc = torch.stack(gathered)
        bases = torch.cumsum(allc, 0) - allc
        counts = allc.sum(0)
        _ = bases
    ms_step, ms_wall, ms_e2e, ms_abs, ms_fmt = [float(x) for x in stats.tolist()]
    Ntot, Ttot, Stot, Dtot = [int(x) for x in counts.tolist()]

    if rank == 0:
        peak, peak_src = measured_peak()
        alg_bytes = N + 8 * T + 8 * S + 8 * D  # per launch (one GPU's shard)
        achieved = alg_bytes / (ms_step * 1e-3) / 1e9
        cpu = None
        if not args.no_cpu and world == 1:
            try:
                cpu = cpu_reference_rate(arr)
                cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
            except Exception as e:  # the oracle is a checker; its absence must not hide the GPU number
                cpu = {"value": None, "unit": "GB/s", "cores": 0, "kind": "port", "sample": f"failed: {e}"}
        kt = {k: round(v / args.steps, 4) for k, v in ktimes.items()}
        # per-kernel algorithmic traffic of one launch: bytes each kernel must move
        out_b = 8 * T + 8 * S + 16 * D  # delta-coded token spans, sentence entries, per-text bounds
        kalg = {"walk_fused": N + 5 * N // 8, "compact_reduce": 4 * N // 8, "compact_texts": 4 * N // 8,
                "compact_emit": 5 * N // 8 + out_b}
        kroof = {k: {"ms": kt[k], "alg_bytes": b, "gbps": b / (kt[k] * 1e-3) / 1e9,
                     "frac": b / (kt[k] * 1e-3) / 1e9 / peak} for k, b in kalg.items() if kt.get(k)}
        dominant = max(kt, key=kt.get)
        # second bound of the walk (R_gather): the bare dependent shared-memory gather chain
        # of one byte step in the walk's own configuration, measured on this device (not part of the step)
        try:
            gsteps = tok.gather_bound()
            walk_rate = N / (kt["walk_fused"] * 1e-3)
            gather = {"peak": gsteps / 1e9, "achieved": walk_rate / 1e9, "unit": "G byte steps/s (= GB/s of input)",
                      "frac": walk_rate / gsteps,
                      "what": "peak: one ld.shared.u8 (class) + one dependent ld.shared.u16 (row entry) per byte and lane, "
                              "1024-thread CTA per SM, nothing else; achieved: input bytes / walk_fused time"}
        except Exception as e:
            gather = {"peak": None, "error": str(e)}
        line = {
            "metric": METRIC, "value": Ntot / (ms_step * 1e-3) / 1e9, "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"C2: tokenizer_de.matok, {N} B synthetic German corpus per GPU, ~10 KB "
                                   "EOT-separated documents, flags TOKENS|SENTENCES|TOKEN_POS|SENTENCE_POS (device leg: DATOK_COMPACT, "
                                   "8-byte delta-coded token spans; host legs: DATOK_COMPACT8, 4 bytes per token)",
                       "bytes_per_gpu": N, "documents_per_gpu": D, "tokens_per_gpu": T, "sentences_per_gpu": S,
                       "l2": "input (>= 1 GiB) and outputs exceed the 126 MB L2; no flush needed",
                       "chunk_bytes": int(os.environ.get("DATOK_CHUNK", "640")),
                       "calibration": "state order specialised once on the first 8 MiB of the corpus (untimed warm-up)",
                       "timing": "CUDA events on the library's stream around the whole device path, max over ranks",
                       "ms_per_step_wall": ms_wall},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": measured_traffic(N), "peak_source": peak_src,
                         "algorithmic_bytes": alg_bytes, "formula": "N + 8*tokens + 8*sentences + 8*documents",
                         "kernel": "whole device path (all kernels of one step); per-kernel ms in kernel_ms, "
                                   "per-kernel rooflines in kernels",
                         "dominant_kernel": dominant, "kernel_ms": kt, "kernels": kroof, "gather": gather},
            "cpu_baseline": cpu,
            "e2e": {"value": Ntot / (ms_e2e * 1e-3) / 1e9, "unit": "GB/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e,
                    "path": "datok_transduce(DATOK_COMPACT8): pinned host input, EOT-aligned pieces, H2D | kernels | D2H "
                            "overlapped; token spans delta-coded (4 B/token), decoded by the host formatter"},
            "e2e_absolute": {"value": Ntot / (ms_abs * 1e-3) / 1e9, "unit": "GB/s", "h2d_bytes_per_step": h2d,
                             "d2h_bytes_per_step": d2h_abs, "ms_per_step": ms_abs,
                             "path": "same call without DATOK_COMPACT8: absolute (byte, rune) offset pairs, 16 B/token"},
            "e2e_formatted": None if world > 1 else {"value": Ntot / (ms_fmt * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": ms_fmt,
                              "text_bytes_per_step": fmt_bytes, "host_threads": os.cpu_count(),
                              "path": "datok_transduce(DATOK_COMPACT8) + datok_format(): the text NewTokenWriter(w, TOKENS|SENTENCES|"
                                      "TOKEN_POS|SENTENCE_POS) writes, formatted on the host cores"},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        print(json.dumps(line))
    tok.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
