"""Calibration of the peft_encoder='both' parity tolerance: PyTorch's own bf16 autocast of the
oracle's two-tower step (CPU) against the reference goldens. Writes nothing; prints one line per
tower. (Run in the authoring container: python tools/parity_both_calib.py)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vit_oracle as vo  # noqa: E402
from tests.golden.make_golden import BOTH_CASES, synth_inputs  # noqa: E402


def rel(a, b):
    a = np.asarray(a, np.float64).ravel(); b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


torch.set_num_threads(os.cpu_count() or 1)
for name in ["both_tiny", "both_vitb16"]:
    cfg, tcfg, n, c, seed = BOTH_CASES[name]
    gold = np.load(os.path.join("tests", "golden", f"ref_{name}.npz"))
    wv, wt = vo.synth_weights(cfg, seed), vo.synth_text_weights(tcfg, seed + 1)
    images, labels = synth_inputs(cfg, n, c, seed + 100)
    tokens = vo.synth_tokens(c, tcfg, seed + 300)
    wvt, wtt = vo.to_torch(wv, torch.float32), vo.to_torch(wt, torch.float32)
    with torch.autocast("cpu", dtype=torch.bfloat16):
        feat = vo.vit_forward(torch.from_numpy(images), wvt, cfg)
        tfeat = vo.text_forward(torch.from_numpy(tokens), wtt, tcfg)
    tn = tfeat.float() / tfeat.float().norm(dim=-1, keepdim=True)
    probs, logits, f = vo.head_forward(feat.float(), tn, float(gold["logit_scale_exp"]))
    loss = vo.reference_loss(probs, torch.from_numpy(labels))
    loss.backward()
    want = {k[5:]: gold[k] for k in gold.files if k.startswith("grad:")}
    got = {k: v.grad.float().numpy() for k, v in {**wvt, **wtt}.items() if v.requires_grad}
    print(f"{name}: tfeat rel {rel(tfeat.float().detach().numpy(), gold['tfeat']):.2e}  probs rel "
          f"{rel(probs.detach().numpy(), gold['probs']):.2e}")
    for tower in ("visual.", "transformer."):
        keys = sorted(k for k in want if k.startswith(tower))
        rels = [rel(got[k], want[k]) for k in keys]
        flat = rel(np.concatenate([got[k].ravel() for k in keys]),
                   np.concatenate([want[k].ravel() for k in keys]))
        print(f"  {tower:13s} torch bf16 autocast vs reference fp32: flat {flat:.2e}  median "
              f"{np.median(rels):.2e}  worst {max(rels):.2e} ({keys[int(np.argmax(rels))]})")
