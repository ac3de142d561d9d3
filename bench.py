#!/usr/bin/env python
"""
bench.py — headline benchmark of the DeTikZify hot path on B200 (contract in the task brief, tier ④).

Workload (BASELINE.json configs[1], named in ``config.workload``): detikzify-ds-1.3b shape, random-init
bf16 weights, ONE synthetic 384x384 figure per GPU, batch-1 greedy generation: ViT encode + concat-3
projector + 243-token image-prefix prefill, then KV-cached single-token decode with the fused sampler
up to a total length of 2048 (1805 new tokens). One "step" = one such figure.

  value  = decoded tokens/s, whole job (sum over GPUs), pixels already resident in HBM, the decode loop
           enqueued as one CUDA-graph launch per token with NO per-token host synchronisation.
  e2e    = the same metric through the public API ``model.generate()``: pixel_values start in pinned
           host memory (H2D inside the timed region) and every generated token is read back by the host
           (the streamer/stopping-criteria contract of the reference) before the next one is consumed.
  roofline = algorithmic HBM bytes of the decode steps (weights once per token + KV read at the running
           context; dtk_decode_bytes) / CUDA-event time of the decode region, vs MEASURED_PEAKS.json.
  cpu_baseline = the oracle (HF Llama+SigLIP wired like the reference; oracle/hf_oracle.py) on the host cores,
           bounded sample.

``--impl reference`` times that CPU path alone (the reference package itself is pure Python glue over
HF modules and does not import offline).
Multi-GPU: figures are independent -> one engine per rank, ONE NCCL broadcast of the weight arena at
load, no per-step collective; scaling is weak (one figure per GPU per step).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="nllg/detikzify-ds-1.3b")
    ap.add_argument("--total-len", type=int, default=2048)
    ap.add_argument("--cpu-tokens", type=int, default=32, help="decode tokens of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-vit-sweep", action="store_true")
    return ap.parse_args()


# ---------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of ``kernel`` from the committed `ncu --set full` capture
    (profiles/r*_ncu_full_<kernel>.json, written by tools/ncu_summary.py); None when no capture is committed."""
    best = None
    for f in sorted((ROOT / "profiles").glob(f"r*_ncu_full_{kernel}.json")):
        try:
            d = json.loads(f.read_text())
            best = {"bytes_per_launch": d["dram_bytes_read"] + d["dram_bytes_write"], "ctx": d.get("ctx"), "source": f"profiles/{f.name}"}
        except (OSError, ValueError, KeyError):
            continue
    return best


# ---------------------------------------------------------------------------------- CPU reference arm
def cpu_decode_tokens_per_s(model_name: str, n_tokens: int, steps: int = 1, warmup: int = 0):
    """Oracle on host cores: ViT + 243-token prefill + n_tokens greedy KV-cached decode steps (fp32 eager).
    Returns (tok/s over the timed steps, seconds per step, cores)."""
    from detikzify_b200.model.configuration import preset
    from detikzify_b200.model.weights import random_init
    from oracle.hf_oracle import Oracle, synthetic_pixels
    # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1 to its workers)
    try:
        import psutil
        phys = psutil.cpu_count(logical=False) or os.cpu_count() or 1
    except Exception:
        phys = os.cpu_count() or 1
    torch.set_num_threads(max(torch.get_num_threads(), int(os.environ.get("DTK_CPU_THREADS", phys))))
    cfg = preset(model_name)
    sd = random_init(cfg, seed=0)
    oracle = Oracle(cfg.to_dict(), sd)
    del sd
    pix = synthetic_pixels(1, cfg.vision_config.image_size)
    ids = torch.full((1, cfg.num_patches), cfg.patch_token_id, dtype=torch.long)
    cores = torch.get_num_threads()

    def one():
        t0 = time.perf_counter()
        logits, cache = oracle.forward_logits(ids, pix, use_cache=True)
        nxt = logits[:, -1].argmax(-1, keepdim=True)
        t1 = time.perf_counter()
        for _ in range(n_tokens):
            logits, cache = oracle.decode_logits(nxt, cache)
            nxt = logits[:, -1].argmax(-1, keepdim=True)
        t2 = time.perf_counter()
        return t2 - t1, t1 - t0

    for _ in range(warmup):
        one()
    dec, pre = 0.0, 0.0
    for _ in range(steps):
        d, p = one()
        dec += d; pre += p
    return n_tokens * steps / dec, dec / steps, pre / steps, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = min(args.cpu_tokens, 16)
    tps, sec, pre, cores = cpu_decode_tokens_per_s(args.model, n, steps=args.steps, warmup=min(args.warmup, 1))
    sample = f"per step: 1 figure, ViT+243-token prefill ({pre:.2f}s, untimed) then {n} greedy KV-cached decode tokens at ctx 243..{243 + n}, fp32 HF eager"
    line = {
        "impl": "reference", "metric": "TikZ tokens/sec/GPU (decode, 384px cond, 2k ctx)", "value": tps, "unit": "tokens/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.model} random-init, 1x384px synthetic figure, batch-1 greedy decode (bounded CPU sample)"},
        "cpu_baseline": {"value": tps, "unit": "tokens/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": tps, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))

    from detikzify_b200.model import load
    from oracle.hf_oracle import synthetic_pixels  # input generator only (seeded uniform pixels)

    model, processor = load(args.model, device_map=local, torch_dtype=torch.bfloat16, broadcast=world > 1, seed=0)
    cfg, eng = model.config, model.engine
    dev = model.device
    P, total = cfg.num_patches, min(args.total_len, eng.max_len)
    n_new = total - P
    ids = torch.full((P,), cfg.patch_token_id, dtype=torch.int64, device=dev)
    pix_host = synthetic_pixels(1, cfg.vision_config.image_size, seed=1000 + rank).pin_memory()
    pix_dev = pix_host.to(dev)
    # greedy, EOS suppressed for the throughput run so every figure decodes the full 1805 tokens
    params = eng.sampling(do_sample=False, bad_token=cfg.image_token_id, begin_suppress_token=-1)
    stream = torch.cuda.Stream(device=dev)
    slot = eng.seq_alloc()

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    def figure(timed: bool):
        """ViT -> projector -> prefill -> first token -> (n_new - 1) graph-launched decode+sample steps."""
        if timed:
            ev[0].record(stream)
        img = eng.image_embeds(pix_dev)[0]
        last, _ = eng.prefill(slot, ids, 0, img, 0)
        first, _ = eng.sample(last, params, suppress=[0])
        tok0 = int(first.item())
        eng.gen_begin([slot], [P], [tok0], params)
        if timed:
            ev[1].record(stream)
        for _ in range(n_new - 1):
            eng.gen_step()
        if timed:
            ev[2].record(stream)
        out = eng.gen_wait(n_new - 2)  # last token has landed on the host
        eng.gen_end()
        return out

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            figure(False)
        barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        launches0 = eng.launch_count
        t_all = t_dec = 0.0
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record(stream)
        dec_ms = []
        for _ in range(args.steps):
            figure(True)
            stream.synchronize()
            dec_ms.append(ev[1].elapsed_time(ev[2]))
        stop.record(stream)
        barrier()
        t_all = start.elapsed_time(stop) / 1e3
        t_dec = sum(dec_ms) / 1e3
        launches = eng.launch_count - launches0
        clocks = sampler.stop() if rank == 0 else None

        # ---- e2e through the public API (host buffers, per-token host visibility)
        e2e_t = None
        if not args.no_e2e:
            from detikzify_b200.util.generation import TokenStreamer
            ids_host = ids.cpu()[None]

            def api_figure():
                st = TokenStreamer()
                out = model.generate(input_ids=ids_host, pixel_values=pix_host, bad_words_ids=[[cfg.image_token_id]],
                                     begin_suppress_tokens=[cfg.eos_token_id], streamer=st, do_sample=False,
                                     max_length=total, eos_token_id=-1)
                assert out.shape[1] == total, out.shape
                return out
            model._img_cache = None
            api_figure()  # warm-up (graph capture for this sampling config)
            barrier()
            s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            for _ in range(args.steps):
                model._img_cache = None      # a new figure every step: ViT + full prefill inside the timed region
                model._slot_tokens = []
                api_figure()
            torch.cuda.synchronize()
            e2e_t = time.perf_counter() - t0
            barrier()

        # ---- secondary metric of BASELINE.json ("ViT encode ms/img", configs[2]: batch sweep @384px), rank 0 only
        vit = None
        if rank == 0 and not args.no_vit_sweep:
            vit = {}
            v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for B in (1, 8, 64):
                pix_b = synthetic_pixels(B, cfg.vision_config.image_size, seed=7).to(dev)
                for _ in range(2):
                    eng.vit_encode(pix_b)
                reps = 5 if B < 64 else 3
                v0.record(stream)
                for _ in range(reps):
                    eng.vit_encode(pix_b)
                v1.record(stream)
                stream.synchronize()
                vit[str(B)] = v0.elapsed_time(v1) / reps / B
                del pix_b
        barrier()

    # max over ranks
    vals = torch.tensor([t_all, t_dec, e2e_t or 0.0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    t_all, t_dec, e2e_t = vals.tolist()

    if rank == 0:
        new_per_step = n_new
        value = world * args.steps * new_per_step / t_all
        # roofline of the decode region (dominant: the per-token decode step = weights + KV stream)
        bytes_dec = sum(eng.decode_bytes(P + 1 + i) for i in range(n_new - 1))
        peak, peak_src = peaks()
        achieved = bytes_dec * args.steps / t_dec / 1e9
        persistent = eng.get_option("decode_persistent") == 1
        kernel_name = ("decode_mega_kernel (persistent cooperative weight-streaming decode kernel, 1 launch per token) + sample_kernel"
                       if persistent else "decode step (CUDA graph: fused RMSNorm+GEMV / split-K attention / sampler kernels of one token)")
        traffic = ncu_traffic("decode_mega_kernel") if persistent else None
        line = {
            "metric": "TikZ tokens/sec/GPU (decode, 384px cond, 2k ctx)", "value": value, "unit": "tokens/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_all / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{args.model} random-init bf16, 1x384px synthetic figure per GPU, batch-1 greedy: ViT + projector + "
                                   f"{P}-token prefill + {n_new} decoded tokens to total length {total}",
                       "l2": "inputs larger than L2: 2.56 GB of weights streamed per token (126 MB L2)",
                       "parallelism": f"figure-sharded dp{world}, 1 NCCL weight broadcast at load, no per-step collective"},
            "roofline": {"bound": "hbm", "kernel": kernel_name,
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "bytes_per_token_avg": bytes_dec / (n_new - 1),
                         "decode_ms_per_token": t_dec / args.steps / (n_new - 1) * 1e3},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if vit:
            # tensor roofline of the ViT (dense contractions, SURVEY.md section 8 a2: 666 GFLOP per image at so400m/14@384)
            vc = cfg.vision_config
            n_tok = (vc.image_size // vc.patch_size) ** 2
            D, Iv, Lv = vc.hidden_size, vc.intermediate_size, vc.num_hidden_layers
            flop_img = Lv * (2 * n_tok * (4 * D * D + 2 * D * Iv) + 4 * n_tok * n_tok * D) + 2 * n_tok * D * 3 * vc.patch_size ** 2
            tpeak, tsrc = 1590.0, "fallback (B200_PROFILING.md 1.59 PFLOP/s)"
            pk = ROOT / "MEASURED_PEAKS.json"
            if pk.exists() and json.loads(pk.read_text()).get("bf16_tflops_sustained"):
                tpeak, tsrc = float(json.loads(pk.read_text())["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
            best_b = min(vit, key=vit.get)
            ach = flop_img / (vit[best_b] * 1e-3) / 1e12
            line["vit_encode_ms_per_img"] = {
                "batch": vit, "note": "SigLIP-so400m/14@384 tokens + pooled output, pixels resident, CUDA events",
                "roofline": {"bound": "tensor", "gflop_per_img": flop_img / 1e9, "achieved": ach, "unit": "TFLOP/s", "at_batch": int(best_b),
                             "peak": tpeak, "frac": ach / tpeak, "peak_source": tsrc}}
        if e2e_t:
            line["e2e"] = {"value": world * args.steps * new_per_step / e2e_t, "unit": "tokens/s",
                           "h2d_bytes_per_step": int(pix_host.numel() * 4 + P * 8), "d2h_bytes_per_step": int(new_per_step * 4)}
        if not args.no_cpu_baseline and world == 1:
            tps, sec, pre, cores = cpu_decode_tokens_per_s(args.model, args.cpu_tokens)
            line["cpu_baseline"] = {"value": tps, "unit": "tokens/s", "cores": cores, "kind": "port",
                                    "sample": f"1 figure: ViT+243-token prefill ({pre:.2f}s, untimed) then {args.cpu_tokens} greedy KV-cached decode tokens "
                                              f"at ctx 243..{243 + args.cpu_tokens} ({sec:.2f}s), fp32 HF eager on {cores} threads"}
        print(json.dumps(line), flush=True)
    eng.seq_free(slot)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
