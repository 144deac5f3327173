#!/usr/bin/env python
"""bench.py -- Whisper large-v2 realtime multiple on B200 (BASELINE.json metric).

    python bench.py --gpus 1 --steps K --warmup W                  # our arm (CUDA engine through the C ABI)
    python bench.py --impl reference --gpus 1 --steps K --warmup W # CPU arm: the oracle port of the reference path
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (configs[1] of BASELINE.json): whisper-large-v2, beam 5, one 3.84 s synthetic 16 kHz utterance per step,
padded to a 30 s window exactly as WIS does (main.py:613).  Weights are seeded synthetic (no checkpoints in the image),
so the decode length is pinned as SURVEY.md section 8(d) prescribes: ceil(3.5 tok/s * 3.84 s) + 1 = 15 generated
tokens (max_length = 30, <|endoftext|> suppressed through CT2's own `suppress_tokens` option), after the 4-token prompt.
One "step" = log-mel + encoder + cross-K/V + 3 prefill + 15 beam-search steps for one utterance per GPU.

  value : whole-job audio-seconds / second with the PCM already resident in HBM (wisb_logmel on a device pointer,
          wisb_generate on the device-resident features), timed with the library's CUDA events on its launching stream
  e2e   : the same metric through the reference-facing surface with HOST buffers:
          audio.log_mel_spectrogram(pcm).numpy() -> StorageView.from_array -> Whisper.generate (wall clock, synchronised)
N > 1   : one process per GPU, utterances are independent (weak scaling, no data-path collective); the weight blob is
          generated on rank 0 and broadcast over NCCL at load time only.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL = "large-v2"
AUDIO_SAMPLES = 61440  # 3.84 s
AUDIO_SECONDS = AUDIO_SAMPLES / 16000.0
BEAM = 5
N_OUT = int(math.ceil(3.5 * AUDIO_SECONDS)) + 1  # 15
MAX_LENGTH = 2 * N_OUT  # -> max_new = min(15, 30 - 4) = 15
PROMPT = [50258, 50259, 50359, 50363]
SEED = 0


def synth_utterance(n_samples: int, seed: int = 1234) -> np.ndarray:
    """SURVEY.md section 8(d) input recipe: 0.3 sin(2 pi (200 + 300 t) t) + 0.05 N(0,1), 16 kHz mono float32."""
    rng = np.random.default_rng(seed)
    t = np.arange(n_samples, dtype=np.float64) / 16000.0
    return (0.3 * np.sin(2.0 * np.pi * (200.0 + 300.0 * t) * t) + 0.05 * rng.standard_normal(n_samples)).astype(np.float32)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


def ncu_dram_traffic_per_launch(summary="r01_gemm_tc_full.csv"):
    """dram__bytes_read + dram__bytes_write per launch from a committed ncu --set full capture under profiles/,
    averaged over the captured launches; None if the summary is missing."""
    import csv

    p = os.path.join(ROOT, "profiles", summary)
    if not os.path.exists(p):
        return None
    rows = list(csv.reader(open(p)))
    hdr, units = rows[0], rows[1]
    try:
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    except ValueError:
        return None
    scale = {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Gbyte": 1e9}
    tot = [float(r[ir]) * scale.get(units[ir], 1e6) + float(r[iw]) * scale.get(units[iw], 1e6) for r in rows[2:]]
    return int(sum(tot) / len(tot)) if tot else None


def encoder_gemm_flops(dims, windows=1):
    """Algorithmic FLOPs of the tcgen05 GEMM launches per window: conv2 + 4 GEMMs per encoder
    layer + the cross-K/V projection, 1500 valid rows each."""
    d, L = dims.d_model, dims.n_enc_layers
    conv2 = 2 * 1500 * (3 * d) * d
    layers = L * 2 * 1500 * d * (3 * d + d + 4 * d + 4 * d)
    ckv = 2 * 1500 * d * (dims.n_dec_layers * 2 * d)
    return windows * (conv2 + layers + ckv), 1 + 4 * L + 1


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.idx)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_blob_host(dims):
    import torch

    from willow_inference_server_b200 import weights as W

    tensors = W.synth_engine_tensors(dims, seed=SEED)
    n = W.blob_nbytes(tensors)
    host = torch.empty(n, dtype=torch.uint8).pin_memory()
    W.write_blob_into(host.numpy(), dims, tensors)
    return host, tensors


def run_ours(args):
    import torch
    import torch.distributed as dist

    from willow_inference_server_b200 import _lib, audio, models, weights as W

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dims = W.WhisperDims.for_size(MODEL)
    # ---- load: rank 0 builds the blob, everyone receives it over NCCL (the only collective in the system)
    from willow_inference_server_b200 import parallel

    t0 = time.time()
    if rank == 0:
        host, tensors = make_blob_host(dims)
    else:
        host, tensors = torch.empty(0, dtype=torch.uint8), None
    torch.cuda.synchronize()
    tb = time.time()
    blob_dev = parallel.broadcast_blob(host, torch.device("cuda", local))
    torch.cuda.synchronize()
    t_bcast = (time.time() - tb) if world > 1 else None
    handle = _lib.Handle.from_device(blob_dev.data_ptr(), blob_dev.numel(), local, keepalive=blob_dev)
    load_s = time.time() - t0
    os.environ["WISB_DEVICE"] = str(local)

    pcm = synth_utterance(AUDIO_SAMPLES, seed=1234 + rank)
    pcm_dev = torch.from_numpy(pcm).cuda()
    prompts = np.array([PROMPT], np.int32)
    off, ns = np.zeros(1, np.int64), np.array([AUDIO_SAMPLES], np.int32)
    extra = [dims.eot]

    def step_device():
        handle.logmel(pcm_dev.data_ptr(), off, ns, to_host=False, keep=True, pcm_on_device=True, pcm_dtype=_lib.PCM_F32, B=1)
        t_l = handle.timing()["logmel_ms"]
        ids, _ = handle.generate(None, prompts, BEAM, 1.0, 1.0, MAX_LENGTH, extra, B=1)
        t = handle.timing()
        return ids, t_l + t["generate_ms"], t

    # reuse_encoder=False: the same utterance is replayed every step, nothing may be skipped behind the benchmark's back
    model = models.Whisper(None, device="cuda", _handles=[handle], reuse_encoder=False)
    pcm_pin = torch.from_numpy(pcm).pin_memory().numpy()

    def step_e2e():
        mel = audio.log_mel_spectrogram(pcm_pin).numpy()[None]  # H2D pcm, D2H mel (what WIS does, main.py:613-616)
        res = model.generate(models.StorageView.from_array(mel), [PROMPT], beam_size=BEAM, max_length=MAX_LENGTH,
                             suppress_tokens=[-1, dims.eot])  # H2D mel, D2H ids
        return res[0].sequences_ids[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (graph capture, allocations)
    for _ in range(max(args.warmup, 3)):
        ids, _, _ = step_device()
    assert len(ids[0]) == N_OUT, f"decode length {len(ids[0])} != pinned {N_OUT}"
    for _ in range(2):
        e_ids = step_e2e()
    assert e_ids == ids[0], "host-buffer path and device-resident path disagree"
    gpu_tokens = [int(t) for t in ids[0]]

    # ---- timed region 1: device-resident inputs
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    w0 = time.perf_counter()
    dev_ms, launches, stage = 0.0, 0, {}
    for _ in range(args.steps):
        ids, ms, t = step_device()
        dev_ms += ms
        launches += int(t["launches"]) + 3
        for k in ("encoder_ms", "cross_kv_ms", "decode_ms", "logmel_ms", "h2d_ms"):
            stage[k] = stage.get(k, 0.0) + t[k]
    barrier()
    wall_dev = time.perf_counter() - w0
    # ---- timed region 2: end to end through the reference-facing surface
    barrier()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    wall_e2e = time.perf_counter() - w0
    clocks = sampler.stop()

    # ---- kernel-level profile pass (not timed): per-family CUDA-event sums
    handle.set_option("profile", 1)
    prof = {}
    for _ in range(3):
        _, _, t = step_device()
        for k in ("gemm_ms", "attn_ms", "ln_ms", "conv1_ms", "gemm_launches", "encoder_ms", "cross_kv_ms"):
            prof[k] = prof.get(k, 0.0) + t[k] / 3
    handle.set_option("profile", 0)

    if world > 1:
        tt = torch.tensor([dev_ms, wall_dev, wall_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms, wall_dev, wall_e2e = tt.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk, pk_src = peaks()
    flops, n_gemm = encoder_gemm_flops(dims)
    gemm_tf = flops / (prof["gemm_ms"] * 1e-3) / 1e12 if prof.get("gemm_ms") else None
    peak_tf = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
    ms_per_step = dev_ms / args.steps
    value = world * AUDIO_SECONDS / (ms_per_step * 1e-3)
    e2e_value = world * AUDIO_SECONDS * args.steps / wall_e2e
    dec_bytes = 2 * (16 * dims.n_dec_layers * dims.d_model ** 2 + dims.n_vocab * dims.d_model) + \
        4 * dims.n_dec_layers * 1500 * dims.d_model
    steps_per = int(t["decode_steps"])  # decoder passes per utterance (prompt prefix in one pass + N_OUT search steps)
    out = {
        "metric": "Whisper large-v2 realtime multiple (audio s / s), beam 5, 3.84 s utterance",
        "value": round(value, 2), "unit": "x realtime", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16", "data": "synthetic",
        "config": {"workload": "whisper-large-v2 beam=5, 3.84 s synthetic 16 kHz utterance, 1 utterance per GPU per step "
                               "(BASELINE.json configs[1])",
                   "model": MODEL, "beam": BEAM, "prompt_len": 4, "generated_tokens": N_OUT, "decoder_passes": steps_per,
                   "weights": f"seeded synthetic (seed {SEED}), fp16 weights / fp32 accumulate",
                   "l2": "no explicit flush: each step streams 3.1 GB of weights (>> 126 MB L2)",
                   "timer": "CUDA events on the library's launching stream (wisb_get_timing), max over ranks",
                   "parallelism": f"dp{world} (independent utterances, weights broadcast once over NCCL)"},
        "p50_latency_ms": round(ms_per_step, 3),
        "wall_ms_per_step": round(1e3 * wall_dev / args.steps, 3),
        "stages_ms": {k: round(v / args.steps, 3) for k, v in stage.items()},
        "e2e": {"value": round(e2e_value, 2), "unit": "x realtime",
                "h2d_bytes_per_step": int(AUDIO_SAMPLES * 4 + 80 * 3000 * 4 + 4 * 4),
                "d2h_bytes_per_step": int(80 * 3000 * 4 + N_OUT * 4 + 8), "ms_per_step": round(1e3 * wall_e2e / args.steps, 3)},
        "gpu_launches": launches,
        "clocks": clocks,
        # dominant kernel of the step (88 % of the serialised ncu launch list, profiles/r01_launches_default.csv): the
        # persistent decoder pass; HBM-bound weight streaming (bytes per pass below)
        "roofline": {"bound": "hbm", "kernel": "dec_pass_kernel<5> (persistent decoder pass, one launch per generated token)",
                     "achieved": round(dec_bytes * steps_per / (stage["decode_ms"] / args.steps * 1e-3) / 1e9, 1),
                     "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": None,
                     "traffic": ncu_dram_traffic_per_launch("r01_dec_pass_kernel_full.csv"),
                     "traffic_note": "dram bytes per launch from the ncu --set full capture in profiles/r01_dec_pass_kernel_full.csv",
                     "peak_source": pk_src + " hbm_gbs",
                     "algorithmic_bytes_per_launch": dec_bytes, "launches_per_step": steps_per,
                     "avg_launch_ms": round(stage["decode_ms"] / args.steps / steps_per, 4),
                     "timing_note": "CUDA events around the decode stage on the launching stream / passes; the stage also holds "
                                    "the 3 small search kernels of every pass (about 2 % of it)"},
        "encoder_roofline": {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05, encoder + cross-K/V GEMMs)",
                             "achieved": round(gemm_tf, 1) if gemm_tf else None, "peak": peak_tf, "unit": "TFLOP/s",
                             "frac": round(gemm_tf / peak_tf, 4) if gemm_tf else None,
                             "traffic": ncu_dram_traffic_per_launch("r01_gemm_tc_full.csv"),
                             "traffic_note": "bytes per launch, mean of the ncu --set full capture in "
                                             "profiles/r01_gemm_tc_full.csv (algorithmic operand bytes: 13.7-21.0 MB)",
                             "peak_source": pk_src + " bf16_tflops_sustained (kernel timed inside a long step)",
                             "algorithmic_flops_per_step": flops, "launches_per_step": n_gemm,
                             "avg_launch_ms": round(prof["gemm_ms"] / n_gemm, 4) if prof.get("gemm_ms") else None,
                             "encoder_share": {k: round(prof[k], 3) for k in ("gemm_ms", "attn_ms", "ln_ms", "conv1_ms")}},
        "load": {"seconds": round(load_s, 1), "nccl_broadcast_s": round(t_bcast, 3) if t_bcast else None,
                 "blob_gb": round(blob_dev.numel() / 1e9, 2)},
    }
    out["roofline"]["frac"] = round(out["roofline"]["achieved"] / pk["hbm_gbs"], 4)
    if not args.no_cpu_baseline and world == 1:
        del tensors, host  # free the host copies before the CPU leg builds its own fp32 model
        out["cpu_baseline"] = cpu_baseline_subprocess()
        toks = out["cpu_baseline"].get("tokens")
        # full-size parity: the fp32 CPU oracle decodes the same utterance with the same weights
        out["tokens_identical_to_cpu_oracle"] = (toks == gpu_tokens) if toks is not None else None
        out["gpu_tokens"] = gpu_tokens
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


CPU_THREADS_MAX = 32  # more OpenMP threads than this slow the small decoder GEMVs down on many-core hosts


def cpu_threads() -> int:
    return max(1, min(os.cpu_count() or 1, CPU_THREADS_MAX))


def cpu_baseline_subprocess(timeout_s: int = 420) -> dict:
    """Run the CPU leg in its own process so that a slow host cannot take the GPU result down with it."""
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-baseline-only"], capture_output=True, text=True,
                           timeout=timeout_s)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode == 0 and line:
            return json.loads(line[-1])
        return {"value": None, "unit": "x realtime", "cores": cpu_threads(), "kind": "port", "sample": "failed: " + r.stderr[-300:]}
    except subprocess.TimeoutExpired:
        return {"value": None, "unit": "x realtime", "cores": cpu_threads(), "kind": "port",
                "sample": f"oracle port did not finish one utterance within {timeout_s} s on this host"}


def cpu_baseline(dims, tensors, steps=1):
    """The oracle port of the reference path (fp32 torch on the host cores) on the same utterance/config."""
    import torch

    from oracle import logmel as om
    from oracle.whisper_ref import WhisperOracle

    cores = cpu_threads()
    torch.set_num_threads(cores)
    oracle = WhisperOracle(dims, tensors)
    pcm = synth_utterance(AUDIO_SAMPLES, seed=1234)
    t0 = time.perf_counter()
    for _ in range(steps):
        mel = om.log_mel_spectrogram(om.pad_or_trim(pcm))[None]
        res = oracle.generate(mel, [PROMPT], beam_size=BEAM, max_length=MAX_LENGTH, suppress_tokens=(-1, dims.eot))
    dt = (time.perf_counter() - t0) / steps
    assert len(res[0].sequences_ids[0]) == N_OUT
    return {"value": round(AUDIO_SECONDS / dt, 4), "unit": "x realtime", "cores": cores, "kind": "port",
            "sample": f"{steps} utterance(s) of the same workload (log-mel + large-v2 encoder + {N_OUT}-token beam-5 decode), "
                      f"fp32 torch oracle, {cores} threads; ctranslate2 is not installable here (no wheel, no network)",
            "seconds_per_utterance": round(dt, 2), "tokens": res[0].sequences_ids[0]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from willow_inference_server_b200 import weights as W

    dims = W.WhisperDims.for_size(MODEL)
    tensors = W.synth_engine_tensors(dims, seed=SEED)
    import torch

    from oracle import logmel as om
    from oracle.whisper_ref import WhisperOracle

    cores = cpu_threads()
    torch.set_num_threads(cores)
    oracle = WhisperOracle(dims, tensors)
    pcm = synth_utterance(AUDIO_SAMPLES, seed=1234)

    def step():
        mel = om.log_mel_spectrogram(om.pad_or_trim(pcm))[None]
        return oracle.generate(mel, [PROMPT], beam_size=BEAM, max_length=MAX_LENGTH, suppress_tokens=(-1, dims.eot))

    # warm-up is bounded too (a step is ~10-20 s of host work): at least one step, then stop after ~30 s
    t_w = time.perf_counter()
    warmed = 0
    while warmed < args.warmup and (warmed == 0 or time.perf_counter() - t_w < 30.0):
        step()
        warmed += 1
    args.warmup = warmed
    # bounded: stop after K steps or ~150 s of host work, whichever comes first (slow hosts: a single step)
    t0 = time.perf_counter()
    done = 0
    while done < args.steps and (done == 0 or time.perf_counter() - t0 < 150.0):
        step()
        done += 1
    dt = (time.perf_counter() - t0) / done
    args.steps = done
    v = round(AUDIO_SECONDS / dt, 4)
    sample = (f"each step = 1 utterance of the same workload on the host cores (fp32 torch oracle port, {cores} threads); "
              "the reference's own engine, ctranslate2==4.1.0, is an un-vendored pip dependency that is absent from "
              "this image and cannot be installed offline")
    print(json.dumps({
        "impl": "reference", "metric": "Whisper large-v2 realtime multiple (audio s / s), beam 5, 3.84 s utterance",
        "value": v, "unit": "x realtime", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(dt * 1e3, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "whisper-large-v2 beam=5, 3.84 s synthetic 16 kHz utterance (BASELINE.json configs[1])",
                   "model": MODEL, "beam": BEAM, "generated_tokens": N_OUT},
        "cpu_baseline": {"value": v, "unit": "x realtime", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "x realtime", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-baseline-only", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.cpu_baseline_only:
        from willow_inference_server_b200 import weights as W

        dims = W.WhisperDims.for_size(MODEL)
        print(json.dumps(cpu_baseline(dims, W.synth_engine_tensors(dims, seed=SEED), steps=1)))
        return
    if args.impl == "reference":
        if args.steps > 3:
            args.steps = 3  # bounded: each step is ~10-30 s of host work
        args.warmup = min(args.warmup, 1)
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
