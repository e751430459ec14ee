#!/usr/bin/env python
"""bench.py - train images/s of the CLIP ViT-B/16 LoRA online step (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

A "step" is one pass of the hot path over one combined stream+replay batch: image tower forward,
head + loss, backward to the LoRA factors (frozen-backbone weight gradients skipped), all-reduce
of the flat LoRA gradient (N > 1) and AdamW. Workload = BASELINE.json configs[1]: ViT-B/16,
stream+replay batch 256, 100 visible classes, synthetic images, random-init weights. N = 1 runs the
256 images on one GPU; N > 1 shards the SAME global batch (256 / N per GPU: strong scaling, the
configuration BASELINE.json states); --scaling weak keeps 256 per GPU instead.

    python bench.py --mode eval                              # BASELINE config 5: inference-only,
                                                             # 4096 images x 1000 cached classes
    python bench.py --peft both                              # LoRA text tower recomputed per step

  value   images/s with the inputs already resident in HBM (CUDA events, max over ranks)
  e2e     images/s through the public API LoRAClipTrainer.online_step(images, labels, idx) fed
          from pinned HOST memory with the RAW CIFAR-shaped uint8 batch the reference's DataLoader
          yields (the reference, too, resizes on the GPU: methods/_trainer.py:236-242); the H2D
          copy of the step's images and labels, the fused resize/crop/flip/normalise and the D2H
          read of (loss, acc) are inside the timed region
  roofline       the tcgen05 GEMM kernel: algorithmic FLOPs / CUDA-event time per launch
  cpu_baseline   the oracle port of the reference path on this box's host cores (bounded sample)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "train img/s, CLIP ViT-B/16 LoRA online step"
UNIT = "img/s"
MODELS = {  # name -> (image, patch, width, layers, heads, embed)
    "ViT-B/16": (224, 16, 768, 12, 12, 512),
    "ViT-L/14": (224, 14, 1024, 24, 16, 768),
}


def train_flops_per_image(model: str) -> float:
    """BASELINE.md §4 / SURVEY.md §8d: forward + activation-gradient backward, no frozen dW."""
    S, p, D, layers, H, E = MODELS[model]
    L, m, r, hd = (S // p) ** 2 + 1, 4 * D, 4, 64
    lin = 2 * L * (3 * D * D + D * D + 2 * D * m)
    attn = 2 * 2 * H * L * L * hd
    lora = 2 * L * (D * r + r * 3 * D + D * r + r * D)
    fwd = 2 * (L - 1) * D * 3 * p * p + layers * (lin + attn + lora) + 2 * D * E
    bwd = layers * (lin + 2 * attn + 2 * lora) + 2 * D * E
    return float(fwd + bwd)


def executed_flops_per_image(model: str) -> float:
    """What the step actually executes: the last block runs class-token-only (out-proj, MLP and
    the Q projection on 1 of L rows; K, V and their backward stay full size), see
    llc_vit_forward_cls. Everything else as train_flops_per_image."""
    S, p, D, layers, H, E = MODELS[model]
    L, m, r, hd = (S // p) ** 2 + 1, 4 * D, 4, 64
    full = train_flops_per_image(model)
    # per image, last block: rows L -> 1 on out-proj + MLP + the Q third of the in-projection (fwd
    # and bwd), attention L*L -> L
    lin_tail = 2 * (L - 1) * (D * D + 2 * D * m + D * D)
    attn_tail = 2 * 2 * H * (L - 1) * L * hd
    lora_tail = 2 * (L - 1) * (D * r + r * D)
    saved = (lin_tail + attn_tail + lora_tail) + (lin_tail + 2 * attn_tail + 2 * lora_tail)
    return float(full - saved)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16": p["bf16_tflops"], "bf16_sustained": p.get("bf16_tflops_sustained"),
                "hbm": p["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clocks and throttle reasons DURING the timed region (B200_PROFILING.md). NVML in a
    sampling thread (every 5 ms, so that even the ~90 ms region of a 32-image-per-GPU run gets a
    dozen samples); falls back to `nvidia-smi -lms 50` when the NVML binding is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.path = gpu_index, None, f"/tmp/llc_clocks_{os.getpid()}.csv"
        self.thread, self.stop_flag, self.samples = None, False, []

    def _nvml_loop(self, nv, handle):
        names = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown),
                 ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown),
                 ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap))
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)
                try:
                    pw = nv.nvmlDeviceGetPowerUsage(handle) / 1000.0
                except Exception:
                    pw = 0.0
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(handle)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(handle)
                self.samples.append((float(sm), pw, [n for n, bit in names if mask & bit]))
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it lists indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.gpu
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                idx = int(vis.split(",")[self.gpu])
            handle = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, handle), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                 "50", "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            if not self.samples:
                return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples"]}
            sm = [s[0] for s in self.samples]
            pw = [s[1] for s in self.samples]
            reasons = sorted({r for s in self.samples for r in s[2]})
            lo = (min(pw) + max(pw)) / 2
            load = [a for a, p in zip(sm, pw) if p >= lo] or sm
            return {"sm_mhz": statistics.median(load), "sm_max_mhz": self.max_mhz,
                    "power_w_max": max(pw), "samples": len(sm), "reasons": reasons,
                    "how": "NVML, 5 ms period, during the timed region"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.f.close()
        sm, mx, pw, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        with open(self.path) as f:
            for line in f:
                c = [x.strip() for x in line.split(",")]
                if len(c) < 8:
                    continue
                try:
                    sm.append(float(c[1])); mx.append(float(c[2])); pw.append(float(c[3]))
                except ValueError:
                    continue
                for n, v in zip(names, c[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load": samples in the upper half of the observed power range
        lo = (min(pw) + max(pw)) / 2
        load = [s for s, p in zip(sm, pw) if p >= lo] or sm
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": max(mx),
                "power_w_max": max(pw), "samples": len(sm), "reasons": sorted(reasons),
                "how": "nvidia-smi -lms 50 during the timed region"}


# ------------------------------------------------------------------------------------------------
class CpuReference:
    """The reference's CPU path: its OWN modules when a checkout is reachable ($LLC_REFERENCE,
    /root/reference, baseline/_ref; kind = "reference"), else the restatement in
    oracle/vit_oracle.py (the reference is Python and absent on the GPU box; kind = "port").
    fp32 torch CPU ops on all host threads: forward + reference loss + backward + AdamW over the
    LoRA tensors."""

    def __init__(self, model: str, classes: int, batch: int, seed: int = 0):
        import numpy as np
        import torch
        from oracle import ref_runner
        from oracle import vit_oracle as vo
        self.torch, self.vo = torch, vo
        S, p, D, layers, H, E = MODELS[model]
        self.cfg = vo.VitCfg(image_size=S, patch=p, width=D, layers=layers, heads=H, embed_dim=E)
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        w_np = vo.synth_weights(self.cfg, seed)
        text_np = vo.synth_text_features(classes, E, seed + 1)
        rng = np.random.default_rng(seed + 2)
        self.x = torch.from_numpy(rng.standard_normal((batch, 3, S, S)).astype(np.float32))
        self.y = torch.from_numpy(rng.integers(0, classes, size=(batch,)).astype(np.int64))
        self.batch = batch
        ref_root = ref_runner.find_reference()
        self.kind = "reference" if ref_root else "port"
        if ref_root:
            self.ref = ref_runner.ReferenceStep(ref_root, self.cfg, w_np, text_np)
        else:
            self.w = vo.to_torch(w_np, torch.float32)
            self.text = torch.from_numpy(text_np)
            self.opt = torch.optim.AdamW([t for t in self.w.values() if t.requires_grad],
                                         lr=1e-3, weight_decay=1e-5)

    def describe(self) -> str:
        return ("the reference's own models/clip modules" if self.kind == "reference" else
                "fp32 oracle port of the reference's PyTorch path")

    def step(self) -> float:
        if self.kind == "reference":
            return self.ref.step(self.x, self.y)
        vo = self.vo
        self.opt.zero_grad(set_to_none=True)
        feat = vo.vit_forward(self.x, self.w, self.cfg)
        probs, logits, _ = vo.head_forward(feat, self.text, 1.0 / 0.07)
        loss = vo.reference_loss(probs, self.y, logits, True)
        loss.backward()
        self.opt.step()
        return float(loss.detach())

    def time_steps(self, steps: int, warmup: int) -> float:
        for _ in range(warmup):
            self.step()
        t0 = time.perf_counter()
        for _ in range(steps):
            self.step()
        return (time.perf_counter() - t0) / steps


ADAPTER_METRIC = "train img/s, adapter-clip online step"


def adapter_workload_config(args, world: int) -> dict:
    cfg = workload_config(args, world)
    cfg["workload"] = (f"CLIP {args.model} adapter-clip online step (bottleneck adapters, "
                       f"ffn 64, dropout 0.1), stream+replay batch {args.batch}, bf16 operands")
    cfg["method"] = "adapter-clip"
    return cfg


class CpuAdapterPort:
    """CPU arm of `--method adapter`: the oracle's restatement of the adapter-clip step
    (oracle/vit_oracle.py: adapter_block_forward in the towers `peft` names, dropout p = 0.1 from
    torch's generator as the reference draws it, reference loss, backward, AdamW over adaptmlp.*)
    in fp32 on all host threads. kind = "port"."""

    kind = "port"

    def __init__(self, model: str, classes: int, batch: int, peft: str, seed: int = 0):
        import numpy as np
        import torch
        from oracle import vit_oracle as vo
        self.torch, self.vo, self.peft = torch, vo, peft
        S, p, D, layers, H, E = MODELS[model]
        self.cfg = vo.VitCfg(image_size=S, patch=p, width=D, layers=layers, heads=H, embed_dim=E)
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        f32 = lambda d: {k: torch.from_numpy(v) for k, v in d.items()}
        self.w = f32(vo.strip_lora(vo.synth_weights(self.cfg, seed)))
        self.w.update({k: torch.zeros(sh) for k, sh in vo.param_shapes(self.cfg).items()
                       if "lora" in k})
        self.train = {}
        if peft in ("image", "both"):
            self.train.update(f32(vo.synth_adapter_weights(
                D, layers, "visual.transformer.resblocks.", seed + 3)))
        self.text = torch.from_numpy(vo.synth_text_features(classes, E, seed + 1))
        self.tcfg = None
        if peft in ("both", "text"):
            from lifelong_clip_b200.adapter_clip import TEXT_CONFIGS
            ctx, vocab, tw, th, tl = TEXT_CONFIGS[model]
            self.tcfg = vo.TextCfg(context=ctx, vocab=vocab, width=tw, heads=th, layers=tl,
                                   embed_dim=E)
            self.wt = f32(vo.strip_lora(vo.synth_text_weights(self.tcfg, seed + 4)))
            self.train.update(f32(vo.synth_adapter_weights(tw, tl, "transformer.resblocks.",
                                                           seed + 5)))
            self.tokens = torch.from_numpy(vo.synth_tokens(classes, self.tcfg, seed + 6))
        for t in self.train.values():
            t.requires_grad_(True)
        rng = np.random.default_rng(seed + 2)
        self.x = torch.from_numpy(rng.standard_normal((batch, 3, S, S)).astype(np.float32))
        self.y = torch.from_numpy(rng.integers(0, classes, size=(batch,)).astype(np.int64))
        self.batch = batch
        self.opt = torch.optim.AdamW(list(self.train.values()), lr=1e-3, weight_decay=1e-5)

    def describe(self) -> str:
        return f"fp32 oracle port of the reference's adapter-clip path (peft_encoder={self.peft})"

    def _tower(self, x, w, prefix, cfg, layers, causal, adapters):
        vo, torch = self.vo, self.torch
        for i in range(layers):
            pre = f"{prefix}{i}."
            if adapters:
                shape = x.shape[:-1] + (vo.ADAPTER_DIM,)
                masks = tuple(torch.rand(shape) >= vo.ADAPTER_DROPOUT for _ in range(2))
                x = vo.adapter_block_forward(x, w, pre, cfg, causal=causal, masks=masks,
                                             p=vo.ADAPTER_DROPOUT)
            else:
                x = vo.block_forward(x, w, pre, cfg, causal=causal)
        return x

    def step(self) -> float:
        vo, torch = self.vo, self.torch
        self.opt.zero_grad(set_to_none=True)
        w = {**self.w, **self.train}
        x = vo.patch_embed(self.x, w, self.cfg)
        x = self._tower(x, w, "visual.transformer.resblocks.", self.cfg, self.cfg.layers, False,
                        self.peft in ("image", "both"))
        feat = vo.layer_norm(x[:, 0, :], w["visual.ln_post.weight"],
                             w["visual.ln_post.bias"]) @ w["visual.proj"]
        if self.tcfg is not None:
            wt = {**self.wt, **self.train}
            wt.update({k: torch.zeros(sh) for k, sh in vo.text_param_shapes(self.tcfg).items()
                       if "lora" in k and k not in wt})
            t = wt["token_embedding.weight"][self.tokens] + wt["positional_embedding"]
            bcfg = vo.VitCfg(width=self.tcfg.width, heads=self.tcfg.heads, layers=self.tcfg.layers)
            t = self._tower(t, wt, "transformer.resblocks.", bcfg, self.tcfg.layers, True, True)
            t = vo.layer_norm(t, wt["ln_final.weight"], wt["ln_final.bias"])
            tf = t[torch.arange(t.shape[0]), self.tokens.argmax(dim=-1)] @ wt["text_projection"]
            text = tf / tf.norm(dim=-1, keepdim=True)
        else:
            text = self.text
        probs, logits, _ = vo.head_forward(feat, text, 1.0 / 0.07)
        loss = vo.reference_loss(probs, self.y, logits, True)
        loss.backward()
        self.opt.step()
        return float(loss.detach())

    time_steps = CpuReference.time_steps


def run_reference(args):
    """--impl reference: rank 0 alone times the CPU path; other ranks exit 0 without work."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    b = args.cpu_batch
    if args.method == "adapter":
        ref = CpuAdapterPort(args.model, args.classes, b, args.peft)
    else:
        ref = CpuReference(args.model, args.classes, b)
    sec = ref.time_steps(args.steps, args.warmup)
    v = b / sec
    sample = (f"{b}-image batch per step (of the {global_batch(args, args.gpus)}-image workload), "
              f"{ref.describe()}, fp32, fwd+loss+bwd+AdamW, {ref.cores} threads")
    adapter = args.method == "adapter"
    print(json.dumps({
        "impl": "reference", "metric": ADAPTER_METRIC if adapter else METRIC, "value": v,
        "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": (adapter_workload_config if adapter else workload_config)(args, args.gpus),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": ref.cores, "kind": ref.kind,
                         "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


def per_gpu_batch(args, world: int) -> int:
    if args.scaling == "strong":
        if args.batch % world:
            raise SystemExit(f"--batch {args.batch} is not divisible by {world} GPUs")
        return args.batch // world
    return args.batch


def global_batch(args, world: int) -> int:
    return args.batch if args.scaling == "strong" else args.batch * world


def workload_config(args, world: int) -> dict:
    """The workload both arms report: BASELINE.json configs[1] (train) / configs[4] (eval)."""
    S, p = MODELS[args.model][0], MODELS[args.model][1]
    return {"workload": workload_name(args), "global_batch": global_batch(args, world),
            "batch_per_gpu": per_gpu_batch(args, world), "classes": args.classes,
            "tokens": (S // p) ** 2 + 1, "parallelism": f"dp{world}", "weights": "random-init",
            "peft_encoder": args.peft, "mode": args.mode,
            "l2": "inputs larger than L2 (154 MB of images per 256; activations 1 GB/layer)",
            "loss": "CE on probabilities (reference double softmax)"}


def workload_name(args) -> str:
    if args.mode == "eval":
        return (f"CLIP {args.model} inference-only eval: masked cosine logits over {args.classes} "
                f"cached class text embeddings, batch {args.batch} (BASELINE.json configs[4])")
    return (f"CLIP {args.model} LoRA online step, stream+replay batch {args.batch} "
            f"(BASELINE.json configs[1]; data-parallel shards of the same global batch), "
            f"bf16 operands")


# ------------------------------------------------------------------------------------------------
def run_eval(args, model, names, dev, world, rank, local_rank):
    """BASELINE.json configs[4]: inference-only evaluation, ViT-B/16, masked cosine logits over
    1000 cached class text embeddings, batch 4096 (sharded over the GPUs: no collective, every
    rank evaluates its own images). A step = tower forward (no saved activations) + the
    tensor-core head (ln_post -> proj GEMM -> L2 norm -> logit GEMM -> softmax/arg-max) + the
    on-device confusion/per-task counters."""
    import torch
    import torch.distributed as dist
    from lifelong_clip_b200 import ops
    from lifelong_clip_b200.trainer import LoRAClipTrainer
    from lifelong_clip_b200.transform import GpuTransform

    S, p, D, layers, H, E = MODELS[args.model]
    B, C = per_gpu_batch(args, world), args.classes
    gB = global_batch(args, world)
    trainer = LoRAClipTrainer(model, names, n_classes=C, n_tasks=100, visible_classes="all")
    model.set_token(names)
    eng = model.model.visual.engine()
    seen = torch.arange(0, C, 2, device=dev)          # seen-class mask: every other class
    mask = torch.full((C,), float("-inf"), device=dev)
    mask[seen] = 0.0
    model.set_additive_mask(mask)
    gen = torch.Generator(device=dev).manual_seed(5 + rank)
    dev_x = [torch.randn(B, 3, S, S, generator=gen, device=dev) for _ in range(2)]
    dev_y = [torch.randint(0, C, (B,), generator=gen, device=dev) for _ in range(2)]
    cm = torch.zeros(C, C, dtype=torch.int64, device=dev)
    counts = torch.zeros(22, dtype=torch.int64, device=dev)
    scale = model.model.logit_scale_exp()

    def step(i):
        eng.forward(dev_x[i % 2], training=False)
        if args.peft in ("both", "text"):
            raise SystemExit("--mode eval measures the cached-text configuration (--peft image)")
        head = eng.eval_head(model._text_all, scale, cls_idx=model._cls_idx,
                             add_mask=model._add_mask, want_probs=True)
        ops.eval_accum(dev_y[i % 2], head.pred, trainer.n_tasks, C, cm, counts)
        return head

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    with torch.no_grad():
        for i in range(args.warmup):
            step(i)
        barrier()
        sampler = ClockSampler(local_rank) if rank == 0 else None
        if sampler:
            sampler.start()
        l0 = ops.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            step(i)
        e1.record()
        barrier()
        launches = ops.launch_count() - l0
        ms_dev = max_over_ranks(e0.elapsed_time(e1)) / args.steps
        clocks = sampler.stop() if sampler else None
        # e2e: online_evaluate over a loader of RAW uint8 32x32 host batches (test_transform =
        # Resize + Normalize fused into the tower's first kernel); the dict comes back on the host
        mean, std = (0.5071, 0.4867, 0.4408), (0.2675, 0.2565, 0.2761)
        trainer.test_transform = GpuTransform.test(S, mean, std)
        hg = torch.Generator().manual_seed(50 + rank)
        host = [(torch.randint(0, 256, (B, 3, 32, 32), generator=hg, dtype=torch.uint8).pin_memory(),
                 torch.randint(0, C, (B,), generator=hg).pin_memory()) for _ in range(2)]
        trainer.online_evaluate([host[i % 2] for i in range(2)], 0)      # warm-up (arena, caches)
        barrier()
        t0 = time.perf_counter()
        res = trainer.online_evaluate([host[i % 2] for i in range(args.steps)], 0)
        torch.cuda.synchronize()
        ms_e2e = max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
        # per-kernel pass
        ops.prof_enable(True)
        for i in range(args.prof_steps):
            step(i)
        torch.cuda.synchronize()
        recs = ops.prof_read()
        ops.prof_enable(False)
    by_kind = {}
    for kind, m_, n_, k_, ms, fl, by in recs:
        d = by_kind.setdefault(kind, {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
        d["launches"] += 1; d["ms"] += ms; d["flops"] += fl; d["bytes"] += by
    peaks = load_peaks()
    gemm = by_kind.get("gemm", {"launches": 0, "ms": 1e-9, "flops": 0.0})
    total_ms = sum(d["ms"] for d in by_kind.values()) or 1.0
    achieved = gemm["flops"] / (gemm["ms"] * 1e-3) / 1e12
    peak = peaks["bf16_sustained"] or peaks["bf16"]
    fwd_flops = train_flops_per_image(args.model)   # forward share computed below
    S_, p_, D_, layers_, H_, E_ = MODELS[args.model]
    L_ = (S_ // p_) ** 2 + 1
    lin = 2 * L_ * (3 * D_ * D_ + D_ * D_ + 2 * D_ * 4 * D_)
    attn = 2 * 2 * H_ * L_ * L_ * 64
    lora = 2 * L_ * (D_ * 4 + 4 * 3 * D_ + D_ * 4 + 4 * D_)
    fwd_flops = float(2 * (L_ - 1) * D_ * 3 * p_ * p_ + layers_ * (lin + attn + lora) +
                      2 * D_ * E_ + 2 * E_ * C)
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    value = gB / (ms_dev * 1e-3)
    out = {
        "metric": "eval img/s, CLIP ViT-B/16 masked cosine logits over cached class text embeddings",
        "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(args, world), "clocks": clocks,
        "e2e": {"value": gB / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": (host[0][0].numel() + host[0][1].numel() * 8) * world,
                "d2h_bytes_per_step": (C * C * 8 + 22 * 8) * world / args.steps,
                "api": "LoRAClipTrainer.online_evaluate(loader of raw uint8 host batches) -> dict",
                "avg_acc": float(res["avg_acc"])},
        "gpu_launches": launches * world,
        "step_tensor_frac": {"achieved_tflops_per_gpu": value / world * fwd_flops / 1e12,
                             "of_burst_peak": value / world * fwd_flops / 1e12 / peaks["bf16"],
                             "flop_per_image": fwd_flops},
        "roofline": {"bound": "tensor", "kernel": "gemm2_kernel / gemm_tn_kernel (tcgen05): tower "
                     "GEMMs + the feature and logit GEMMs of the evaluation head",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": achieved / peak, "traffic": None,
                     "peak_source": peaks["source"],
                     "share_of_step": gemm["ms"] / total_ms},
        "kernel_breakdown": {k: {"ms_per_step": d["ms"] / args.prof_steps,
                                 "launches_per_step": d["launches"] / args.prof_steps}
                             for k, d in sorted(by_kind.items())},
        "cpu_baseline": None,
    }
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------
def run_maple(args, dev, world, rank, local_rank):
    """--method maple (BASELINE.json configs[3]): MaPLe multi-modal deep prompts on frozen ViT-B/16
    towers (models/maple.py:74-253): text encoder over the class prompts (C x 77 tokens, compound
    prompts spliced at layers 1-2) + image encoder (197 + 3 prompt tokens) forward and backward
    down to the prompt rows, cosine logits, CE on the logits (methods/maple.py:96), AdamW on
    prompt_learner.* (torch's: ~1.2 M parameters in 9 tensors). Data parallel: images shard, every
    rank runs the text side, the prompt gradients are all-reduced."""
    import torch
    import torch.distributed as dist
    from lifelong_clip_b200 import ops
    from lifelong_clip_b200.adapter_clip import SyntheticTokenizer
    from lifelong_clip_b200.maple import MaPLe

    S, p, D, layers, H, E = MODELS[args.model]
    B, C = per_gpu_batch(args, world), args.classes
    gB = global_batch(args, world)
    torch.manual_seed(0)
    m = MaPLe(model_name=args.model, vision_config=(S, p, D, layers, E)).to(dev)
    m.set_tokenizer(SyntheticTokenizer())
    for k, prm in m.named_parameters():          # methods/maple.py: only the prompt learner trains
        prm.requires_grad = "prompt_learner" in k
    m.update_class_names([f"class {i}" for i in range(C)])
    params = [prm for prm in m.parameters() if prm.requires_grad]
    opt = torch.optim.AdamW(params, lr=1e-3, weight_decay=1e-5)
    gen = torch.Generator().manual_seed(100 + rank)
    host_x = [torch.randn(B, 3, S, S, generator=gen).pin_memory() for _ in range(2)]
    host_y = [torch.randint(0, C, (B,), generator=gen).pin_memory() for _ in range(2)]
    dev_x = [x.to(dev) for x in host_x]
    dev_y = [y.to(dev) for y in host_y]

    graphed = None
    if not args.no_graph:
        from lifelong_clip_b200.maple import GraphedStep
        graphed = GraphedStep(m, dev_x[0], dev_y[0], gB)
    def step(x, y):
        if graphed is not None:
            loss = graphed(x, y)
        else:
            opt.zero_grad(set_to_none=True)
            logits = m(x)
            loss = torch.nn.functional.cross_entropy(logits, y, reduction="sum") / gB
            loss.backward()
        if world > 1:
            flat = torch.cat([prm.grad.reshape(-1) for prm in params])
            dist.all_reduce(flat)
            off = 0
            for prm in params:
                prm.grad.copy_(flat[off:off + prm.numel()].view_as(prm))
                off += prm.numel()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for i in range(args.warmup):
        step(dev_x[i % 2], dev_y[i % 2])
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(dev_x[i % 2], dev_y[i % 2])
    e1.record()
    barrier()
    launches = ops.launch_count() - l0
    ms_dev = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    clocks = sampler.stop() if sampler else None
    if graphed is not None:     # libllc kernel nodes per replay = launches of one eager step
        g_keep, graphed = graphed, None
        n0 = ops.launch_count()
        step(dev_x[0], dev_y[0])
        launches = (ops.launch_count() - n0) * args.steps
        graphed = g_keep
    # e2e: pinned host fp32 batches -> device -> step -> loss float on the host
    last = None
    for i in range(args.warmup + args.steps):
        if i == args.warmup:
            barrier()
            e0.record()
        x = host_x[i % 2].to(dev, non_blocking=True)
        y = host_y[i % 2].to(dev, non_blocking=True)
        last = float(step(x, y))
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    graphed = None              # per-launch events need eager launches
    ops.prof_enable(True)
    for i in range(args.prof_steps):
        step(dev_x[i % 2], dev_y[i % 2])
    torch.cuda.synchronize()
    recs = ops.prof_read()
    ops.prof_enable(False)
    by_kind = {}
    for kind, m_, n_, k_, ms, fl, by in recs:
        d = by_kind.setdefault(kind, {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
        d["launches"] += 1; d["ms"] += ms; d["flops"] += fl; d["bytes"] += by
    peaks = load_peaks()
    gemm = by_kind.get("gemm", {"launches": 0, "ms": 1e-9, "flops": 0.0})
    total_ms = sum(d["ms"] for d in by_kind.values()) or 1.0
    achieved = gemm["flops"] / (gemm["ms"] * 1e-3) / 1e12
    peak = peaks["bf16_sustained"] or peaks["bf16"]
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    cfg = workload_config(args, world)
    cfg["workload"] = (f"MaPLe multi-modal prompt tuning on CLIP {args.model}: text + image encoder "
                       f"forward / backward with deep prompts (n_ctx 3, depth 3), batch {args.batch}, "
                       f"{C} classes (BASELINE.json configs[3])")
    cfg["method"] = "maple"
    cfg["tokens"] = (S // p) ** 2 + 1 + 3
    cfg["loss"] = "CE on the logits (methods/maple.py:96)"
    out = {
        "metric": "train img/s, MaPLe prompt-tuning step", "value": gB / (ms_dev * 1e-3),
        "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": cfg,
        "clocks": clocks, "trainable_params": sum(prm.numel() for prm in params),
        "impl_details": {"cuda_graph": not args.no_graph},
        "e2e": {"value": gB / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": (host_x[0].numel() * 4 + host_y[0].numel() * 8) * world,
                "d2h_bytes_per_step": 4 * world,
                "api": "logits = MaPLe(images); CE; backward; AdamW(prompt_learner) -> loss float",
                "last_loss": last},
        "gpu_launches": launches * world,
        "roofline": {"bound": "tensor", "kernel": "gemm2_kernel (tcgen05 cta_group::2 / TMEM): the "
                     "frozen blocks' dense contractions, forward and activation-gradient",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "traffic": None, "peak_source": peaks["source"],
                     "share_of_step": gemm["ms"] / total_ms},
        "kernel_breakdown": {k: {"ms_per_step": d["ms"] / args.prof_steps,
                                 "launches_per_step": d["launches"] / args.prof_steps}
                             for k, d in sorted(by_kind.items())},
        "cpu_baseline": None,
    }
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------
def run_adapter(args, dev, world, rank, local_rank):
    """--method adapter: the reference's second PEFT method on the same trainer
    (scripts/adapter_clip.sh -> ResidualAttentionBlock_Adapter in the towers peft_encoder names,
    models/clip/model.py:418-442). A step = block-by-block forward (dropout p = 0.1 on the
    bottleneck) + loss + backward + all-reduce of the flat adapter gradient + fused AdamW."""
    import torch
    import torch.distributed as dist
    from lifelong_clip_b200 import ops
    from lifelong_clip_b200.adapter_clip import AdapterCLIP, SyntheticTokenizer
    from lifelong_clip_b200.trainer import DevicePrefetcher, LoRAClipTrainer
    from lifelong_clip_b200.transform import GpuTransform

    S, p, D, layers, H, E = MODELS[args.model]
    B, C = per_gpu_batch(args, world), args.classes
    gB = global_batch(args, world)
    torch.manual_seed(0)
    model = AdapterCLIP(model_name=args.model, peft_method="adapter", peft_encoder=args.peft,
                        vision_config=(S, p, D, layers, E)).to(dev)
    # the reference's init leaves up_proj at zero (an identity adapter): randomise it so that the
    # timed step differentiates a non-trivial function
    with torch.no_grad():
        for a in model.adapters():
            a.up_proj.weight.normal_(0, 0.02)
    names = [f"class {i}" for i in range(C)]
    if args.peft in ("both", "text"):
        model.set_tokenizer(SyntheticTokenizer())
    else:
        model.set_text_features(names, torch.randn(C, E, generator=torch.Generator().manual_seed(1)))
    trainer = LoRAClipTrainer(model, names, n_classes=C, n_tasks=5, lr=1e-3, online_iter=1,
                              visible_classes="all", sharded_input=True)
    trainer.online_before_task(0)
    trainer.add_new_class(torch.arange(C))
    model.set_token(trainer.exposed_classes_names)
    lut = trainer._class_lut(trainer.exposed_classes)
    gen = torch.Generator().manual_seed(100 + rank)
    host_raw = [torch.randint(0, 256, (B, 3, 32, 32), generator=gen, dtype=torch.uint8).pin_memory()
                for _ in range(3)]
    host_y = [torch.randint(0, C, (B,), generator=gen).pin_memory() for _ in range(3)]
    dev_x = [torch.randn(B, 3, S, S, device=dev,
                         generator=torch.Generator(device=dev).manual_seed(7 + i + 10 * rank))
             for i in range(2)]
    dev_y = [ops.label_remap(y.to(dev), lut) for y in host_y[:2]]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for i in range(args.warmup):
        trainer.block_step(dev_x[i % 2], dev_y[i % 2], gB, sync=False)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        trainer.block_step(dev_x[i % 2], dev_y[i % 2], gB, sync=False)
    e1.record()
    barrier()
    launches = ops.launch_count() - l0
    ms_dev = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    clocks = sampler.stop() if sampler else None

    mean, std = (0.5071, 0.4867, 0.4408), (0.2675, 0.2565, 0.2761)
    trainer.train_transform = GpuTransform.train(S, mean, std)
    idx = torch.arange(B)

    def host_loader(n):
        for i in range(n):
            yield host_raw[i % 3], host_y[i % 3], idx

    last = None
    for i, (images, labels, ids) in enumerate(
            DevicePrefetcher(host_loader(args.warmup + args.steps), dev)):
        if i == args.warmup:
            barrier()
            e0.record()
        last = trainer.online_step(images, labels, ids)
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    trainer.train_transform = (lambda x: x)

    ops.prof_enable(True)
    for i in range(args.prof_steps):
        trainer.block_step(dev_x[i % 2], dev_y[i % 2], gB, sync=False)
    torch.cuda.synchronize()
    recs = ops.prof_read()
    ops.prof_enable(False)
    agg = {}
    for kind, m_, n_, k_, ms, fl, by in recs:
        d = agg.setdefault((kind, m_, n_, k_), [0, 0.0, 0.0])
        d[0] += 1; d[1] += ms; d[2] += by
    by_shape = sorted(({"kind": kk[0], "m": kk[1], "n": kk[2], "k": kk[3],
                        "launches_per_step": v[0] / args.prof_steps,
                        "us_per_launch": 1e3 * v[1] / v[0],
                        "gbs": (v[2] / (v[1] * 1e-3) / 1e9) if v[1] else 0.0}
                       for kk, v in agg.items()),
                      key=lambda r: -r["us_per_launch"] * r["launches_per_step"])[:16]
    by_kind = {}
    for kind, m_, n_, k_, ms, fl, by in recs:
        if kind == "gemm" and (n_ < 256 or k_ < 256):
            kind = "gemm_skinny"        # adapter projections (N = 64 / K = 64), rank-r products
        d = by_kind.setdefault(kind, {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
        d["launches"] += 1; d["ms"] += ms; d["flops"] += fl; d["bytes"] += by
    peaks = load_peaks()
    gemm = by_kind.get("gemm", {"launches": 0, "ms": 1e-9, "flops": 0.0})
    total_ms = sum(d["ms"] for d in by_kind.values()) or 1.0
    achieved = gemm["flops"] / (gemm["ms"] * 1e-3) / 1e12
    peak = peaks["bf16_sustained"] or peaks["bf16"]
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        ref = CpuAdapterPort(args.model, C, args.cpu_batch, args.peft)
        sec = ref.time_steps(1, 1)
        cpu = {"value": args.cpu_batch / sec, "unit": UNIT, "cores": ref.cores, "kind": ref.kind,
               "sample": f"{args.cpu_batch}-image batch x 1 timed step (1 warm-up) of the same "
                         f"model/classes: {ref.describe()}, fp32, fwd+loss+bwd+AdamW"}
    cfg = adapter_workload_config(args, world)
    n_adapter = sum(p_.numel() for a in model.adapters() for p_ in a.parameters())
    out = {
        "metric": ADAPTER_METRIC, "value": gB / (ms_dev * 1e-3),
        "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": cfg,
        "clocks": clocks, "trainable_params": n_adapter,
        "e2e": {"value": gB / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": (host_raw[0].numel() + host_y[0].numel() * 8) * world,
                "d2h_bytes_per_step": 8 * world,
                "api": "for images, labels, idx in DevicePrefetcher(pinned host loader): "
                       "LoRAClipTrainer.online_step(images, labels, idx) -> (loss, acc)",
                "last_loss_acc": list(last) if last else None},
        "gpu_launches": launches * world,
        "roofline": {"bound": "tensor", "kernel": "gemm2_kernel (tcgen05 cta_group::2 / TMEM): the "
                     "frozen blocks' dense contractions", "achieved": achieved, "peak": peak,
                     "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                     "peak_source": peaks["source"], "share_of_step": gemm["ms"] / total_ms},
        "kernel_breakdown": {k: {"ms_per_step": d["ms"] / args.prof_steps,
                                 "launches_per_step": d["launches"] / args.prof_steps,
                                 "gbs": (d["bytes"] / (d["ms"] * 1e-3) / 1e9) if d["ms"] else 0.0}
                             for k, d in sorted(by_kind.items())},
        "top_launches": by_shape,
        "cpu_baseline": cpu,
    }
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun (one process per GPU): python -m "
                             f"torch.distributed.run --nnodes=1 --nproc-per-node {args.gpus} "
                             f"--master-addr 127.0.0.1 bench.py --gpus {args.gpus} ...")
        raise SystemExit(f"WORLD_SIZE={world} does not match --gpus {args.gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this framework has no CPU path "
                         "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from lifelong_clip_b200 import ops
    from lifelong_clip_b200.adapter_clip import AdapterCLIP
    from lifelong_clip_b200.trainer import LoRAClipTrainer

    from lifelong_clip_b200.adapter_clip import SyntheticTokenizer
    from lifelong_clip_b200.transform import GpuTransform

    if args.method in ("adapter", "maple"):
        if args.mode != "train":
            raise SystemExit(f"--method {args.method} measures the training step")
        run = run_adapter if args.method == "adapter" else run_maple
        return run(args, dev, world, rank, local_rank)
    S, p, D, layers, H, E = MODELS[args.model]
    B, C = per_gpu_batch(args, world), args.classes
    gB = global_batch(args, world)
    torch.manual_seed(0)                       # identical replicas on every rank
    model = AdapterCLIP(model_name=args.model, peft_encoder=args.peft,
                        vision_config=(S, p, D, layers, E)).to(dev)
    names = [f"class {i}" for i in range(C)]
    g = torch.Generator().manual_seed(1)
    if args.peft in ("both", "text"):
        model.set_tokenizer(SyntheticTokenizer())
    else:
        model.set_text_features(names, torch.randn(C, E, generator=g))
    if args.mode == "eval":
        return run_eval(args, model, names, dev, world, rank, local_rank)
    trainer = LoRAClipTrainer(model, names, n_classes=C, n_tasks=5, lr=1e-3, online_iter=1,
                              visible_classes="all", sharded_input=True,
                              use_cuda_graph=not args.no_graph)
    trainer.online_before_task(0)

    # synthetic stream: a pool of distinct batches per rank. `value` reads pre-transformed fp32
    # 224x224 images resident in HBM; `e2e` starts from the RAW CIFAR-shaped uint8 batch in
    # pinned host memory (what the reference's DataLoader yields before its GPU transform).
    n_pool = 3
    RAW = 32
    gen = torch.Generator().manual_seed(100 + rank)
    host_raw = [torch.randint(0, 256, (B, 3, RAW, RAW), generator=gen, dtype=torch.uint8)
                .pin_memory() for _ in range(n_pool)]
    host_y = [torch.randint(0, C, (B,), generator=gen).pin_memory() for _ in range(n_pool)]
    idx = torch.arange(B)
    mean, std = (0.5071, 0.4867, 0.4408), (0.2675, 0.2565, 0.2761)   # datasets/__init__.py:38-39

    # every class exposed once up front so the visible-class list (C columns) is fixed
    trainer.add_new_class(torch.arange(C))
    model.set_token(trainer.exposed_classes_names)
    lut = trainer._class_lut(trainer.exposed_classes)
    dev_x = [torch.randn(B, 3, S, S, generator=torch.Generator(device=dev).manual_seed(7 + i + 10 * rank),
                         device=dev) for i in range(2)]
    dev_y = [ops.label_remap(y.to(dev), lut) for y in host_y[:2]]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------------------------------------------------------- DP equivalence (N > 1)
    # one-off: the all-reduced gradient of the sharded global batch against the gradient rank 0
    # computes alone on the concatenated batch
    dp_check = None
    if world > 1:
        gg = torch.Generator().manual_seed(999)
        gx = torch.randn(gB, 3, S, S, generator=gg)
        gy = torch.randint(0, C, (gB,), generator=gg)
        engines = trainer.optimizer.engines()
        trainer._step_body(gx[rank::world].to(dev), ops.label_remap(gy[rank::world].to(dev), lut),
                           gB)
        red = torch.cat([e.grad_flat for e in engines]).clone()
        dist.all_reduce(red)
        if rank == 0:
            shard_sum = None
            for r in range(world):       # every shard again on this one GPU
                trainer._step_body(gx[r::world].to(dev),
                                   ops.label_remap(gy[r::world].to(dev), lut), gB)
                part = torch.cat([e.grad_flat for e in engines])
                shard_sum = part.clone() if shard_sum is None else shard_sum + part
            trainer._step_body(gx.to(dev), ops.label_remap(gy.to(dev), lut), gB)
            full = torch.cat([e.grad_flat for e in engines])
            dp_check = {"rel_l2": float((red - full).norm() / full.norm()),
                        "rel_l2_vs_sum_of_shards_on_one_gpu":
                            float((red - shard_sum).norm() / shard_sum.norm()),
                        "max_abs": float((red - full).abs().max()),
                        "grad_norm": float(full.norm()),
                        "what": f"all-reduce of {world} shard gradients vs one GPU on the "
                                f"concatenated {gB}-image batch (rel_l2: the fp32 order of "
                                f"largely cancelling token sums) and vs the sum of the same "
                                f"shards computed on one GPU (the exchange itself)"}
            for e in engines:           # back to the per-GPU shard size
                e.arena, e.arena_key, e.dx = None, None, None
        barrier()

    # ---------------------------------------------------------------- device-resident: `value`
    for i in range(args.warmup):
        trainer.fused_step(dev_x[i % 2], dev_y[i % 2], gB, sync=False)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        trainer.fused_step(dev_x[i % 2], dev_y[i % 2], gB, sync=False)
    e1.record()
    barrier()
    launches = ops.launch_count() - l0     # eager llc_* launches (AdamW, and everything w/o graph)
    if trainer.use_cuda_graph:             # + the kernel nodes of every graph replay
        launches += trainer.graph_kernels * args.steps
    ms_dev = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    clocks = sampler.stop() if sampler else None
    value = gB / (ms_dev * 1e-3)

    # ---------------------------------------------------------------- end to end: `e2e`
    # the user-facing loop: DevicePrefetcher over a loader of pinned HOST batches (the H2D copy of
    # batch i+1 overlaps step i on a side stream) -> online_step -> (loss, acc) floats on the
    # host. The batches are RAW uint8 32x32 (CIFAR-shaped); the reference's train_transform
    # (Resize 224 / RandomCrop / flip / Normalize, methods/_trainer.py:236-242) runs fused in
    # front of the patch embedding.
    from lifelong_clip_b200.trainer import DevicePrefetcher
    trainer.train_transform = GpuTransform.train(S, mean, std)

    def host_loader(n):
        for i in range(n):
            yield host_raw[i % n_pool], host_y[i % n_pool], idx

    # one continuous loop in steady state (as a long run is): W untimed steps, then exactly K
    # timed ones. The prefetcher stages one batch ahead, so the copies of the first timed batch
    # (and of batch W+1) are issued before e0; the other K-1 (K-2) are inside the timed region,
    # as are all K transforms, steps and (loss, acc) reads.
    last = None
    for i, (images, labels, ids) in enumerate(
            DevicePrefetcher(host_loader(args.warmup + args.steps), dev)):
        if i == args.warmup:
            barrier()
            e0.record()
        last = trainer.online_step(images, labels, ids)
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    e2e_value = gB / (ms_e2e * 1e-3)
    h2d = host_raw[0].numel() + host_y[0].numel() * 8
    d2h = 8
    trainer.train_transform = (lambda x: x)

    # ---------------------------------------------------------------- per-kernel pass (roofline)
    trainer.use_cuda_graph = False          # per-launch events need eager launches
    ops.prof_enable(True)
    for i in range(args.prof_steps):
        trainer.fused_step(dev_x[i % 2], dev_y[i % 2], gB, sync=False)
    torch.cuda.synchronize()
    recs = ops.prof_read()
    ops.prof_enable(False)
    trainer.use_cuda_graph = not args.no_graph
    # per (kind, m, n, k) launch statistics of the instrumented steps
    agg = {}
    for kind, m, n, k, ms, fl, by in recs:
        d = agg.setdefault((kind, m, n, k), [0, 0.0])
        d[0] += 1; d[1] += ms
    by_shape = [{"kind": kk[0], "m": kk[1], "n": kk[2], "k": kk[3],
                 "launches_per_step": v[0] / args.prof_steps, "us_per_launch": 1e3 * v[1] / v[0],
                 "ms_per_step": v[1] / args.prof_steps} for kk, v in agg.items()]
    by_shape.sort(key=lambda r: -r["ms_per_step"])
    by_kind = {}
    for kind, m, n, k, ms, fl, by in recs:
        if kind == "gemm" and n < 256:
            kind = "gemm_skinny"   # rank-r row products (N = 16): HBM-bound gemm_tn_kernel<32>
        d = by_kind.setdefault(kind, {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
        d["launches"] += 1; d["ms"] += ms; d["flops"] += fl; d["bytes"] += by
    peaks = load_peaks()
    gemm = by_kind.get("gemm", {"launches": 0, "ms": 0.0, "flops": 0.0})
    total_ms = sum(d["ms"] for d in by_kind.values()) or 1.0
    # DRAM bytes per launch of that kernel from the committed ncu --set full capture (forward
    # launches of one block inside this same bench command; profiles/r02_ncu_gemm2_in_step.csv)
    traffic = None
    try:
        if B != 256:
            raise LookupError("the ncu capture was taken at 256 images per GPU")
        import csv
        with open(os.path.join(ROOT, "profiles", "r02_ncu_gemm2_in_step.csv")) as f:
            rows = list(csv.DictReader(f))
        tot = [(float(r["dram__bytes_read.sum"]) + float(r["dram__bytes_write.sum"])) * 1e6
               for r in rows]
        traffic = sum(tot) / len(tot)
    except Exception:
        traffic = None
    roof = None
    if gemm["launches"]:
        achieved = gemm["flops"] / (gemm["ms"] * 1e-3) / 1e12
        # the launches are timed inside instrumented steps that follow the timed region directly
        # (a long, power-capped run): the denominator is the SUSTAINED measured cuBLAS figure; the
        # burst figure (a kernel timed alone on a cool part) is reported next to it
        peak = peaks["bf16_sustained"] or peaks["bf16"]
        roof = {"bound": "tensor", "kernel": "gemm2_kernel (tcgen05 cta_group::2 / TMEM): the 8 dense contractions of every block, forward and backward",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": traffic,
                "peak_burst": peaks["bf16"], "frac_of_burst_peak": achieved / peaks["bf16"],
                "traffic_note": "mean dram read+write bytes per launch, ncu --set full, 8 forward "
                                "launches inside this bench (profiles/r02_ncu_gemm2_in_step.csv)",
                "peak_source": peaks["source"] + (", sustained (kernel timed inside a long step)"
                                                  if peaks["bf16_sustained"] else ", burst"),
                "flop_per_launch": gemm["flops"] / gemm["launches"],
                "us_per_launch": gemm["ms"] * 1e3 / gemm["launches"],
                "launches_per_step": gemm["launches"] / args.prof_steps,
                "share_of_step": gemm["ms"] / total_ms,
                "timed": f"CUDA events around each launch, {args.prof_steps} instrumented steps "
                         "right after the timed region"}
    breakdown = {k: {"ms_per_step": d["ms"] / args.prof_steps,
                     "launches_per_step": d["launches"] / args.prof_steps,
                     "tflops": (d["flops"] / (d["ms"] * 1e-3) / 1e12) if d["ms"] else 0.0,
                     "gbs": (d["bytes"] / (d["ms"] * 1e-3) / 1e9) if d["ms"] else 0.0}
                 for k, d in sorted(by_kind.items())}
    if args.dump_prof and rank == 0:
        os.makedirs(os.path.dirname(os.path.abspath(args.dump_prof)), exist_ok=True)
        with open(args.dump_prof, "w") as f:
            json.dump({"by_shape": by_shape, "by_kind": breakdown, "records": recs}, f)

    # ---------------------------------------------------------------- CPU baseline (rank 0, N=1)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ref = CpuReference(args.model, C, args.cpu_batch)
        sec = ref.time_steps(2, 1)
        cpu = {"value": args.cpu_batch / sec, "unit": UNIT, "cores": ref.cores, "kind": ref.kind,
               "sample": f"{args.cpu_batch}-image batch x 2 timed steps (1 warm-up) of the same "
                         f"model/classes: {ref.describe()}, fp32, fwd+loss+bwd+AdamW"}

    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    fl_ref = train_flops_per_image(args.model)
    cls_only = os.environ.get("LLC_FULL_LAST_BLOCK") is None
    fl = executed_flops_per_image(args.model) if cls_only else fl_ref
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(args, world),
        "impl_details": {"cuda_graph": not args.no_graph,
                         "last_block": ("class-token rows only (llc_vit_forward_cls: identical "
                                        "outputs)" if os.environ.get("LLC_FULL_LAST_BLOCK") is None
                                        else "full"),
                         "value_input": "pre-transformed fp32 224x224 images resident in HBM",
                         "e2e_input": "raw uint8 32x32 batches in pinned host memory; resize / "
                                      "crop / flip / normalise fused into the step "
                                      "(llc_vit_forward_tx)"},
        "dp_check": dp_check,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                "api": "for images, labels, idx in DevicePrefetcher(pinned host loader): "
                       "LoRAClipTrainer.online_step(images, labels, idx) -> (loss, acc)",
                "last_loss_acc": list(last) if last else None},
        "gpu_launches": launches * world,
        "step_tensor_frac": {"achieved_tflops_per_gpu": value / world * fl / 1e12,
                             "of_burst_peak": value / world * fl / 1e12 / peaks["bf16"],
                             "of_sustained_peak": (value / world * fl / 1e12 / peaks["bf16_sustained"]
                                                   if peaks["bf16_sustained"] else None),
                             "flop_per_image": fl,
                             "flop_note": ("executed FLOPs: the last block is computed for the "
                                           "class-token rows only (identical results)"
                                           if cls_only else "reference algorithm, every row"),
                             "reference_flop_per_image": fl_ref,
                             "of_sustained_peak_at_reference_flops": (
                                 value / world * fl_ref / 1e12 / peaks["bf16_sustained"]
                                 if peaks["bf16_sustained"] else None)},
        "roofline": roof,
        "kernel_breakdown": breakdown,
        "cpu_baseline": cpu,
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="ViT-B/16", choices=sorted(MODELS))
    ap.add_argument("--batch", type=int, default=None,
                    help="images per step: the GLOBAL batch under --scaling strong (default 256 "
                         "train / 4096 eval), per GPU under --scaling weak")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--mode", default="train", choices=["train", "eval"])
    ap.add_argument("--peft", default="image", choices=["image", "both", "text"],
                    help="'both': LoRA text tower recomputed every step (scripts/lora_clip.sh); "
                         "'text': only the text tower trains, the image tower runs forward only")
    ap.add_argument("--method", default="lora", choices=["lora", "adapter", "maple"],
                    help="'adapter': the adapter-clip method (scripts/adapter_clip.sh); 'maple': "
                         "MaPLe prompt tuning (BASELINE configs[3]) - secondary lines; the "
                         "headline metric is the lora-clip step")
    ap.add_argument("--classes", type=int, default=None)
    ap.add_argument("--cpu-batch", type=int, default=16, help="images per CPU-baseline step")
    ap.add_argument("--prof-steps", type=int, default=2)
    ap.add_argument("--dump-prof", default=None, help="write per-launch records (json) here")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly")
    args = ap.parse_args()
    if args.batch is None:
        args.batch = 4096 if args.mode == "eval" else 256
    if args.classes is None:
        args.classes = 1000 if args.mode == "eval" else 100
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
