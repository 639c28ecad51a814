#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native Qwen3-TTS hot path.

Metric (BASELINE.json): real-time factor = audio seconds / wall seconds, plus p50 first-audio
latency. A "step" is one pass of the hot path over one batch of synthetic input: at N GPUs every
rank synthesises `--utterances` C2-shaped utterances (BASELINE.json configs[1]: 0.6B-Base, batch 1,
30 s = 375 frames, English, temp 0.8 / top-k 50 / top-p 0.95, seeded Philox) through
lqt_synthesize_tokens (prompt assembly -> prefill -> frame loop A+B -> vocoder, all on the device).
Request-level data parallelism, no collective on the data path => "scaling": "weak".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--frames F] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

value  : audio-s / s with inputs resident on the device (CUDA-event time measured by the library on
         its own stream, prompt build -> last vocoder kernel), max over ranks.
e2e    : the same through the C-ABI with HOST buffers: token ids in, PCM + codes out (pinned),
         wall clock between barrier+synchronize pairs, max over ranks.
roofline: the frame loop (weight-streaming GEMV kernels) against the measured HBM peak:
         algorithmic bytes per frame x frames / CUDA-event time of the frame loop.
cpu_baseline / --impl reference: the CPU oracle executing the REFERENCE's schedule (48 graph calls
         per frame, cache-less code predictor, whole-KV copy per step: src/tts_onnx.cpp:782-872) on
         the box's host cores -- the stand-in for the ORT CPU path, which cannot be built here
         (no onnxruntime, no .onnx graphs).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

load_package()
from leaxer_qwen3_tts_b200 import modelspec as ms  # noqa: E402

METRIC = "real-time factor (audio s / wall s)"
UNIT = "audio-s/s"
FRAME_S = 0.08                      # 1920 samples @ 24 kHz per frame (src/tts_onnx.h:69)
WORKLOAD = ("0.6B-Base batch-1 decode, 30 s English prompt (375 frames, 90 synthetic text ids), "
            "temp 0.8 / top-k 50 / top-p 0.95, Philox seed 1234 [BASELINE.json configs[1]]")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def frame_bytes(spec: ms.ModelSpec, mean_kv_pos: float) -> dict:
    """Algorithmic bytes one frame must move at batch 1 (bf16 weights, no cross-pass reuse):
    talker step (all layers + final norm + codec head) + KV read + 15 predictor passes
    (5 layers + one head each; 16 token passes through the body incl. the extra row of pass 0 is
    still one weight read per pass in the minimum-traffic schedule: pass 0 reads the body once for
    its 2 rows). SURVEY.md §8d."""
    def layer(H, qd, kvd, I):
        return H * (qd + 2 * kvd) + qd * H + 3 * H * I
    talker = spec.layers * layer(spec.hidden, spec.q_dim, spec.kv_dim, spec.inter) + spec.vocab * spec.hidden
    cp_body = spec.cp_layers * layer(spec.cp_hidden, spec.cp_q_dim, spec.cp_kv_dim, spec.cp_inter)
    cp_head = spec.cp_vocab * spec.cp_hidden
    kv = spec.layers * 2 * spec.kv_dim * 2 * mean_kv_pos
    return {"talker": 2 * talker, "predictor": spec.cp_steps * 2 * (cp_body + cp_head), "kv": kv,
            "total": 2 * talker + spec.cp_steps * 2 * (cp_body + cp_head) + kv}


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, dev: int):
        self.dev, self.proc, self.lines = dev, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.dev), f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower() == "active":
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU stand-in for the reference's ORT CPU path (oracle, reference schedule)
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(mdir: str, token_ids, frames: int, threads: int, seed: int = 1234, warm: bool = True):
    """-> (audio seconds, wall seconds) of one bounded sample: prompt + `frames` frames + vocoder."""
    import torch
    from oracle import qwen3_tts_oracle as orc
    torch.set_num_threads(threads)
    m = cpu_reference_run._model if getattr(cpu_reference_run, "_mdir", None) == mdir else None
    if m is None:
        m = orc.OracleModel(mdir)
        cpu_reference_run._model, cpu_reference_run._mdir = m, mdir
        if warm:                      # materialise the fp32 weight views outside the timed region
            p = orc.SamplingParams(max_new_tokens=1, seed=seed)
            orc.synthesize_tokens(m, token_ids, "en", p, schedule="reference")
    p = orc.SamplingParams(temperature=0.8, top_k=50, top_p=0.95, max_new_tokens=frames, seed=seed)
    t0 = time.perf_counter()
    audio, codes = orc.synthesize_tokens(m, token_ids, "en", p, schedule="reference")
    dt = time.perf_counter() - t0
    return codes.shape[0] * FRAME_S, dt


def run_reference_arm(a, mdir, token_ids):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    # size the per-step sample so that (steps + warmup) steps end within ~3 minutes
    t0 = time.perf_counter()
    cpu_reference_run(mdir, token_ids, 1, cores)
    _, t1f = cpu_reference_run(mdir, token_ids, 2, cores)
    per_frame = max(t1f / 2.0, 1e-3)
    budget = 150.0 / max(a.steps + a.warmup, 1)
    frames = int(max(1, min(a.frames, budget / per_frame)))
    log(f"[reference] oracle load+warm {time.perf_counter() - t0:.1f}s, ~{per_frame * 1e3:.0f} ms/frame "
        f"-> {frames} frames per step")
    for _ in range(a.warmup):
        cpu_reference_run(mdir, token_ids, frames, cores)
    audio_s = wall = 0.0
    for _ in range(a.steps):
        s, dt = cpu_reference_run(mdir, token_ids, frames, cores)
        audio_s += s; wall += dt
    v = audio_s / wall
    sample = (f"first {frames} frames of the C2 utterance per step (prompt assembly + prefill + frames + "
              f"vocoder), reference schedule (48 graph calls/frame), torch CPU fp32, {cores} threads")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": wall / a.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames": frames, "frames_full_workload": a.frames,
                       "frames_note": "each step is a bounded sample: the FIRST `frames` frames of the same utterance (prompt + prefill + frames + "
                                      "vocoder), sized so that the run ends within minutes; the b200 arm runs all frames_full_workload frames. "
                                      "Shorter runs flatter the CPU side (the reference's whole-KV copy per step grows with the position).",
                       "parallelism": "cpu host cores, rank 0 only"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "ONNX Runtime and the .onnx graphs are not available offline; this is the CPU oracle "
                    "(oracle/qwen3_tts_oracle.py) running the reference's schedule"}
    print(json.dumps(line), flush=True)
    return 0


def run_c5(a, rank, world, local, token_ids):
    """BASELINE configs[4]: 1.7B talker (hidden 2048, MLP 6144; the predictor stays 1024 wide behind in_proj), long-form
    utterances of `--frames` (default 2048 = 163.84 s) frames, 64 utterances in total split over the GPUs, through
    lqt_synthesize_batch (the persistent batch-1 kernel does not take this shape; lqt_stats.frame_impl_active says so)."""
    import torch
    spec = ms.spec_1p7b(0)
    mdir = ms.default_model_dir(spec)
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if local == 0:
        ms.generate_model_dir(mdir, spec)
    if dist:
        dist.barrier()
    from leaxer_qwen3_tts_b200 import engine
    eng = engine.Engine(mdir, device=local, frame_impl="auto")
    frames = a.frames if a.frames != 375 else 2048
    per = max(1, 64 // world)
    spf = eng.info.samples_per_frame
    ids_np = np.asarray(token_ids, np.int64)
    reqs = [{"token_ids": ids_np, "lang": "en", "utterance_id": rank * per + u, "max_new_tokens": frames} for u in range(per)]
    pins = [(torch.empty(frames * spf, dtype=torch.float32).pin_memory().numpy(),
             torch.empty(frames * 16, dtype=torch.int64).pin_memory().numpy()) for _ in range(per)]
    best = None
    for rep in range(2):
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        t0 = time.perf_counter()
        outs = eng.synthesize_batch(reqs, 0.8, 50, 0.95, seed=1234, max_concurrent=per, planes=a.c4_planes, pinned=pins)
        torch.cuda.synchronize()
        best = (time.perf_counter() - t0, sum(o[1].shape[0] for o in outs))
    v = torch.tensor([best[0]], dtype=torch.float64, device=f"cuda:{local}")
    n = torch.tensor([float(best[1])], dtype=torch.float64, device=f"cuda:{local}")
    if dist:
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
    if rank == 0:
        line = {"metric": METRIC, "value": float(n.item()) * FRAME_S / float(v.item()), "unit": UNIT, "n_gpus": world, "steps": 1, "warmup": 1,
                "ms_per_step": float(v.item()) * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic",
                "config": {"workload": f"1.7B talker, {per * world} utterances of {frames} frames ({frames * FRAME_S:.2f} s) each, {per} per GPU, "
                                       f"planes={a.c4_planes} [BASELINE.json configs[4]]", "spec": spec.name, "frames": frames,
                           "frame_impl_active": int(eng.stats().frame_impl_active), "parallelism": f"dp{world} (request-level, no collective)"},
                "e2e": {"value": float(n.item()) * FRAME_S / float(v.item()), "unit": UNIT, "h2d_bytes_per_step": int(ids_np.nbytes * per),
                        "d2h_bytes_per_step": int((frames * spf * 4 + frames * 16 * 8) * per)},
                "gpu_launches": int(eng.stats().kernel_launches)}
        print(json.dumps(line), flush=True)
    eng.close()
    if dist:
        dist.destroy_process_group()
    return 0


def parity_check(eng, a) -> dict:
    """Outside the timed region: the benchmarked configuration against the committed oracle golden
    (tests/golden/c2_full_375.npz, written by tests/golden/make_c2_golden.py from oracle/qwen3_tts_oracle.py).
    (1) teacher-forced over all 375 frames: the engine is fed the oracle's codes and its logits are compared at the stored
    frames (north_star: <= 2e-2 max-abs); (2) free-running with the benchmark's sampler settings and Philox key (1234, 0):
    length of the token-exact prefix (a draw within float noise of a CDF boundary ends it; tests/test_gpu_parity.py)."""
    gpath = os.path.join(ROOT, "tests", "golden", "c2_full_375.npz")
    if not (a.spec == "0.6b" and a.frames == 375 and os.path.exists(gpath)):
        return {"parity_checked": False, "why": "no golden for this configuration (only BASELINE configs[1] has one)"}
    g = np.load(gpath)
    tol = 2e-2
    prompt, trailing, pad = eng.build_prompt(g["token_ids"], "en")
    sp = eng.sampling(0.8, 50, 0.95, 375, seed=4321, utterance_id=9)
    codes, tb = eng.generate(prompt, trailing, pad, sp, forced_codes=g["codes"], trace=True)
    worst = 0.0
    for i, f in enumerate(g["frames"]):
        ref0, refc = g["talker_logits"][i], g["cp_logits"][i]
        fin = np.isfinite(ref0)
        worst = max(worst, float(np.abs(tb[f, 0, :3072][fin] - ref0[fin]).max()), float(np.abs(tb[f, 1:, :2048] - refc).max()))
    del tb
    _, free = eng.synthesize_tokens(g["token_ids"], "en", 0.8, 50, 0.95, 375, 1234, 0)
    diff = np.argwhere(free != g["codes"])
    prefix = 375 if len(diff) == 0 else int(diff[0][0])
    ok = bool(np.array_equal(codes, g["codes"]) and worst < tol and prefix >= 8)
    return {"parity_checked": ok, "golden": "tests/golden/c2_full_375.npz (CPU oracle, 375 frames)",
            "teacher_forced_frames": 375, "logit_frames_compared": int(len(g["frames"])), "max_abs_logit_err": worst,
            "logit_tol": tol, "free_running_exact_prefix_frames": prefix}


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=375, help="max_new_tokens per utterance (375 = 30 s)")
    ap.add_argument("--utterances", type=int, default=1, help="utterances per GPU per step (sequential)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--spec", default="0.6b", choices=["0.6b", "1.7b", "tiny"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true", help="skip the (untimed) comparison with the committed oracle golden")
    ap.add_argument("--c4-utterances", type=int, default=256,
                    help="BASELINE configs[3] leg (after the headline): this many concurrent utterances IN TOTAL, split evenly over the "
                         "GPUs, through lqt_synthesize_batch (tcgen05 GEMM path); 0 = skip")
    ap.add_argument("--no-c3", action="store_true", help="skip the voice-clone leg (BASELINE configs[2])")
    ap.add_argument("--c5", action="store_true",
                    help="BASELINE configs[4] leg instead of the headline: 1.7B talker, 2048 frames (163.84 s) per utterance, 64 utterances "
                         "split over the GPUs (8 per GPU at 8 GPUs), batched tcgen05 path")
    ap.add_argument("--c4-planes", type=int, default=2, help="bf16 planes per activation in the batched leg (3 = fp32-exact, 2 = 16-bit mantissa)")
    ap.add_argument("--cpu-frames", type=int, default=150, help="frames in the cpu_baseline sample (~10-30 s of CPU work)")
    ap.add_argument("--frame-impl", default="persistent", choices=["persistent", "graph"],
                    help="persistent = one cooperative kernel per utterance (default); graph = round-1 v1 schedule (A/B only)")
    a = ap.parse_args()
    if a.warmup < 3 and a.impl == "b200":
        log("[bench] note: timing rules ask for >= 3 warm-up steps")

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    spec = {"0.6b": ms.spec_0p6b, "1.7b": ms.spec_1p7b, "tiny": ms.spec_tiny}[a.spec](0)
    mdir = ms.default_model_dir(spec)
    from leaxer_qwen3_tts_b200.engine import wrap_text_ids
    token_ids = wrap_text_ids(ms.synthetic_text_ids(90, 1234))

    if a.impl == "reference":
        if rank == 0:
            ms.generate_model_dir(mdir, spec)
        return run_reference_arm(a, mdir, token_ids)
    if a.c5:
        return run_c5(a, rank, world, local, token_ids)

    import torch
    if not torch.cuda.is_available():
        log("bench.py: no CUDA device; the product path has no CPU fallback")
        return 2
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # weights: rank 0 of the node writes the seeded random-init model dir, the others wait
    if local == 0:
        t0 = time.time()
        ms.generate_model_dir(mdir, spec)
        log(f"[bench] model dir {mdir} ready ({time.time() - t0:.1f}s)")
    if dist:
        dist.barrier()

    from leaxer_qwen3_tts_b200 import dispatch, engine
    eng = engine.Engine(mdir, device=local, frame_impl=a.frame_impl)
    my_utts = dispatch.assign([a.frames] * (a.utterances * world), world)[rank]      # weak scaling: `--utterances` per GPU
    spf = eng.info.samples_per_frame
    audio_pin = torch.empty(a.frames * spf, dtype=torch.float32).pin_memory().numpy()
    codes_pin = torch.empty(a.frames * 16, dtype=torch.int64).pin_memory().numpy()
    ids_np = np.asarray(token_ids, np.int64)

    def one_step(step_idx):
        dev_ms = gen_ms = voc_ms = 0.0
        nfr = 0
        first = []
        for u in my_utts:                                                 # this rank's share (dispatch.assign: LPT, no collective)
            utt = dispatch.utterance_key(1234, u * 1000003 + step_idx)[1]   # Philox key from the GLOBAL utterance index: invariant to the GPU count
            audio, codes = eng.synthesize_tokens(ids_np, "en", 0.8, 50, 0.95, a.frames, 1234, utt,
                                                 audio_out=audio_pin, codes_out=codes_pin)
            st = eng.stats()
            dev_ms += st.last_total_ms; gen_ms += st.last_generate_ms; voc_ms += st.last_vocoder_ms
            first.append(st.first_audio_ms)
            nfr += codes.shape[0]
            assert audio.shape[0] == codes.shape[0] * spf
        return dev_ms, gen_ms, voc_ms, nfr, first

    def fence():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    parity = parity_check(eng, a) if rank == 0 and not a.no_parity_check else None
    if parity is not None:
        log(f"[bench] parity: {parity}")
    for i in range(a.warmup):
        one_step(-1 - i)
    eng.reset_stats()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    fence()
    t0 = time.perf_counter()
    dev_ms = gen_ms = voc_ms = 0.0
    frames = 0
    first_ms = []
    for i in range(a.steps):
        d, g, v, n, fa = one_step(i)
        dev_ms += d; gen_ms += g; voc_ms += v; frames += n
        first_ms += fa
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    fence()
    clocks = sampler.stop() if rank == 0 else None
    st = eng.stats()

    # ---- BASELINE configs[3]: 256 concurrent utterances split over the GPUs, batched tcgen05 path (strong scaling) ----------
    c4 = None
    if a.c4_utterances > 0 and a.spec == "0.6b":
        per = max(1, a.c4_utterances // world)
        reqs = [{"token_ids": ids_np, "lang": "en", "utterance_id": rank * per + u, "max_new_tokens": a.frames} for u in range(per)]
        pins = [(torch.empty(a.frames * spf, dtype=torch.float32).pin_memory().numpy(),
                 torch.empty(a.frames * 16, dtype=torch.int64).pin_memory().numpy()) for _ in range(per)]
        best = None
        for rep in range(2):                                      # one warm-up (context + graph creation), one timed
            fence()
            t0 = time.perf_counter()
            outs = eng.synthesize_batch(reqs, 0.8, 50, 0.95, seed=1234, max_concurrent=per, planes=a.c4_planes, pinned=pins)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = (dt, sum(o[1].shape[0] for o in outs), eng.stats().last_frames)
        c4v = torch.tensor([best[0]], dtype=torch.float64, device=f"cuda:{local}")
        c4n = torch.tensor([float(best[1])], dtype=torch.float64, device=f"cuda:{local}")
        if dist:
            dist.all_reduce(c4v, op=dist.ReduceOp.MAX)
            dist.all_reduce(c4n, op=dist.ReduceOp.SUM)
        c4 = {"workload": f"{per * world} concurrent C2-shaped utterances ({per} per GPU), {a.frames} frames each, planes={a.c4_planes} "
                          "[BASELINE.json configs[3]]", "value": float(c4n.item()) * FRAME_S / float(c4v.item()), "unit": UNIT,
              "utterances_per_gpu": per, "wall_s": float(c4v.item()), "lockstep_frames_rank0": int(best[2]),
              "note": "end to end through lqt_synthesize_batch with pinned host buffers (token ids in, PCM + codes out), wall clock, "
                      "max over ranks; strong scaling: the total is fixed, each GPU takes total / n_gpus utterances"}
        log(f"[bench] c4: {c4}")

    # ---- BASELINE configs[2]: voice clone (3 s synthetic reference clip -> device log-mel -> speaker encoder -> prompt row),
    # zh/ja/ko prompts, 125 frames, batch 1 (rank 0 only: a latency figure, not a scaling one) ------------------------------
    c3 = None
    if rank == 0 and a.spec == "0.6b" and not a.no_c3:
        t = np.arange(72000) / 24000.0
        clip = (0.4 * np.sin(2 * np.pi * 220 * t) + 0.25 * np.sin(2 * np.pi * 1330 * t) + 0.1 * np.sin(2 * np.pi * 5100 * t)
                + 0.02 * np.random.default_rng(3).standard_normal(72000)).astype(np.float32)
        c3_ids = np.asarray(wrap_text_ids(ms.synthetic_text_ids(30, 77)), np.int64)
        walls = []
        for i, lang in enumerate(["zh", "ja", "ko", "zh", "ja", "ko"]):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            spk = eng.speaker_embed_audio(clip)
            audio, codes = eng.synthesize_tokens(c3_ids, lang, 0.8, 50, 0.95, 125, 1234, 9000 + i, speaker_embed=spk,
                                                 audio_out=audio_pin, codes_out=codes_pin)
            torch.cuda.synchronize()
            if i >= 3:
                walls.append(time.perf_counter() - t0)
            assert codes.shape[0] == 125
        c3 = {"workload": "voice clone: 3 s 24 kHz synthetic reference clip (host f32) -> log-mel + speaker encoder on the device -> "
                          "125 frames (10 s), zh/ja/ko, batch 1 [BASELINE.json configs[2]]",
              "value": 125 * FRAME_S / statistics.median(walls), "unit": UNIT, "ms_per_utterance_p50": statistics.median(walls) * 1e3,
              "h2d_bytes": int(clip.nbytes + c3_ids.nbytes)}
        log(f"[bench] c3: {c3}")

    vals = torch.tensor([wall, dev_ms, gen_ms, voc_ms], dtype=torch.float64, device=f"cuda:{local}")
    tot = torch.tensor([float(frames), float(st.kernel_launches)], dtype=torch.float64, device=f"cuda:{local}")
    if dist:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    wall_max, dev_ms_max, gen_ms_max, voc_ms_max = [float(x) for x in vals.tolist()]
    frames_all, launches_all = [float(x) for x in tot.tolist()]
    audio_s = frames_all * FRAME_S
    value = audio_s / (dev_ms_max * 1e-3)
    e2e = audio_s / wall_max

    if rank != 0:
        eng.close()
        if dist:
            dist.destroy_process_group()
        return 0

    # roofline of the frame loop (rank 0's own loop; ranks are identical replicas)
    hbm_peak, peak_src = peaks()
    P = 9
    fb = frame_bytes(spec, mean_kv_pos=P + (a.frames - 1) / 2.0)
    frames_r0 = frames if frames else 1
    gen_s = gen_ms * 1e-3
    achieved = fb["total"] * frames_r0 / gen_s / 1e9 if gen_s > 0 else 0.0
    traffic = None                      # DRAM bytes of one frame_kernel launch from the committed ncu capture (same workload only)
    tpath = os.path.join(ROOT, "profiles", "r2s3_traffic.json")
    if a.frames == 375 and a.spec == "0.6b" and a.frame_impl == "persistent" and os.path.exists(tpath):
        traffic = json.load(open(tpath))["dram_bytes_per_launch"]
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic,
                "traffic_note": "DRAM bytes per utterance (9 prefill rows + 375 frames, one frame_kernel launch), ncu capture of the same command (profiles/r2s3_summary.md section 1, r2s3_final_launches.csv.gz); algorithmic bytes per utterance = total x 375",
                "kernel": "frame_kernel (persistent cluster kernel: TMA weight stream + tensor-core matrix-vector phases + attention + sampler), per frame",
                "algorithmic_bytes_per_frame": fb, "us_per_frame": gen_s / frames_r0 * 1e6,
                "peak_source": peak_src}

    cpu = None
    if not a.no_cpu_baseline:
        cores = os.cpu_count() or 1
        t0 = time.time()
        s, dt = cpu_reference_run(mdir, token_ids, a.cpu_frames, cores)
        cpu = {"value": s / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"first {a.cpu_frames} frames of the same C2 utterance (prompt + prefill + frames + vocoder), "
                         f"reference schedule, torch CPU fp32, {cores} threads, {dt:.1f}s of CPU work"}
        log(f"[bench] cpu_baseline took {time.time() - t0:.1f}s incl. oracle load")

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": dev_ms_max / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD if a.frames == 375 and a.spec == "0.6b" else f"{a.spec} {a.frames} frames (non-headline)",
                   "spec": spec.name, "frames": a.frames, "utterances_per_gpu_per_step": a.utterances,
                   "parallelism": f"dp{world} (request-level, no collective)",
                   "l2": "inputs larger than L2: 3.3 GB of weights streamed per frame vs 126 MB L2",
                   "kv_cache": "paged bf16", "frame_impl": a.frame_impl},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(ids_np.nbytes * a.utterances),
                "d2h_bytes_per_step": int((a.frames * spf * 4 + a.frames * 16 * 8) * a.utterances),
                "ms_per_step": wall_max / a.steps * 1e3,
                "first_audio_ms_p50": statistics.median(first_ms) if first_ms else None,
                "full_utterance_ms_p50": wall_max / a.steps / a.utterances * 1e3},
        "c3_clone": c3,
        "c4_batched": c4,
        "parity_checked": bool(parity and parity.get("parity_checked")),
        "parity": parity,
        "gpu_launches": int(launches_all),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "first_audio_ms_p50": statistics.median(first_ms) if first_ms else None,
        "full_utterance_ms_p50": wall_max / a.steps / a.utterances * 1e3,
        "first_audio_note": "rank 0, CUDA events: request start -> first chunk (4 frames = 320 ms of PCM; later chunks 25 frames) vocoded on a second stream and copied into the "
                            "caller's pinned buffer while the frame kernel generates the rest; the reference has no streaming (first audio = full utterance)",
        "breakdown_ms_per_step": {"frame_loop": gen_ms_max / a.steps, "vocoder": voc_ms_max / a.steps,
                                  "device_total": dev_ms_max / a.steps, "host_wall": wall_max / a.steps * 1e3},
    }
    print(json.dumps(line), flush=True)
    eng.close()
    if dist:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
