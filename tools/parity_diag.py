"""Per-layer forward/backward error of the CUDA blocks vs the fp64 oracle (dev tool): where does
the LoRA-gradient error of the 12-layer tower come from?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import vit_oracle as vo
from lifelong_clip_b200.clip_modules import ResidualAttentionBlock_LoRA

cfg = vo.VIT_B16
N = int(os.environ.get("N", 2))
w = vo.synth_weights(cfg, 7)
wd = vo.to_torch(w, torch.float64)
L, D = cfg.tokens, cfg.width
g = torch.Generator().manual_seed(0)
x0 = torch.randn(N, L, D, generator=g)             # stand-in for ln_pre output
dy = torch.randn(N, L, D, generator=g) * 1e-3

def rel(a, b):
    a = a.double().flatten(); b = b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))

# oracle
xs_o = [x0.double().requires_grad_(True)]
for i in range(cfg.layers):
    y = vo.block_forward(xs_o[-1], wd, f"visual.transformer.resblocks.{i}.", cfg)
    y.retain_grad(); xs_o.append(y)
xs_o[-1].backward(dy.double())

# CUDA blocks on [L, N, D]
blocks = []
for i in range(cfg.layers):
    pre = f"visual.transformer.resblocks.{i}."
    b = ResidualAttentionBlock_LoRA(D, cfg.heads, None, {"lora_alpha": 1, "lora_r": 4})
    b.load_state_dict({k[len(pre):]: torch.from_numpy(v) for k, v in w.items() if k.startswith(pre)})
    blocks.append(b.cuda())
xs_c = [x0.transpose(0, 1).contiguous().cuda().requires_grad_(True)]
for b in blocks:
    y = b(xs_c[-1]); y.retain_grad(); xs_c.append(y)
xs_c[-1].backward(dy.transpose(0, 1).contiguous().cuda())
torch.cuda.synchronize()
print("layer  fwd x_out rel   bwd dx_in rel   grad rel: in_A in_B out_A out_B")
for i in range(cfg.layers):
    pre = f"visual.transformer.resblocks.{i}."
    f = rel(xs_c[i + 1].detach().cpu().transpose(0, 1), xs_o[i + 1].detach())
    bw = rel(xs_c[i].grad.cpu().transpose(0, 1), xs_o[i].grad)
    gs = []
    for k in ("attn.in_proj_weight_lora_A", "attn.in_proj_weight_lora_B", "attn.out_proj.lora_A",
              "attn.out_proj.lora_B"):
        p = dict(blocks[i].named_parameters())[k]
        gs.append(rel(p.grad.cpu(), wd[pre + k].grad))
    print(f"{i:3d}   {f:.3e}   {bw:.3e}   " + "  ".join(f"{v:.3e}" for v in gs))
