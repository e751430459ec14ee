"""Calibration of the adapter-clip parity tolerance: PyTorch's own bf16 autocast of the oracle's
adapter step (CPU, same injected dropout masks) against the reference goldens. The bottleneck's
ReLU gate makes the gradient discontinuous: a bf16-rounded pre-activation opens a few gates
differently from fp32, and each flipped gate is an O(1) error of that element's gradient.
Prints one line per tower. (Run in the authoring container: python tools/parity_adapter_calib.py)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vit_oracle as vo  # noqa: E402
from tests.golden.make_golden import load_grads  # noqa: E402
from tests.test_oracle_golden import adapter_case_inputs  # noqa: E402


def rel(a, b):
    a = np.asarray(a, np.float64).ravel(); b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


torch.set_num_threads(os.cpu_count() or 1)
for name in ["adapter_tiny", "adapter_vitb16"]:
    cfg, tcfg, wv, wt, wa, wta, images, labels, tokens, masks, tmasks = adapter_case_inputs(name)
    gold = np.load(os.path.join("tests", "golden", f"ref_{name}.npz"))
    with torch.autocast("cpu", dtype=torch.bfloat16):
        out = vo.adapter_step_oracle(
            images, labels, wv, wa, None, cfg, logit_scale_exp=float(gold["logit_scale_exp"]),
            dtype=torch.float32, masks=vo.masks_sample_major(masks), p=vo.ADAPTER_DROPOUT,
            tokens=tokens, wt_np=wt, wta_np=wta, tcfg=tcfg, tmasks=vo.masks_sample_major(tmasks))
    want = load_grads(gold)
    got = out["grads"]
    print(f"{name}: probs rel {rel(out['probs'], gold['probs']):.2e}  loss "
          f"{float(out['loss']):.5f} vs {float(gold['loss']):.5f}")
    for tower in ("visual.", "transformer."):
        keys = sorted(k for k in want if k.startswith(tower))
        rels = [rel(got[k], want[k]) for k in keys]
        flat = rel(np.concatenate([got[k].ravel() for k in keys]),
                   np.concatenate([want[k].ravel() for k in keys]))
        print(f"  {tower:13s} torch bf16 autocast vs reference fp32: flat {flat:.2e}  median "
              f"{np.median(rels):.2e}  worst {max(rels):.2e} ({keys[int(np.argmax(rels))]})")
