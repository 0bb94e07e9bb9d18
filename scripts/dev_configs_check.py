"""Functional check of the other BASELINE.json configs (L4 depth ablation, R101 + H2048) at a small batch, plus a
short throughput probe at batch 256 (developer tool, run under gpurun)."""
import sys, time
import torch
sys.path.insert(0, ".")
from bench import synth_host_batch
from virtex_b200.config import Config
from virtex_b200.factories import PretrainingModelFactory
from virtex_b200.trainer import Trainer

cases = [("depth_ablations/bicaptioning_R_50_L4_H1024.yaml", []),
         ("backbone_ablations/bicaptioning_R_101_L1_H1024.yaml", ["MODEL.TEXTUAL.NAME", "transdec_postnorm::L1_H2048_A32_F8192"]),
         ("_base_bicaptioning_R_50_L1_H1024.yaml", ["MODEL.TEXTUAL.NAME", "transdec_prenorm::L2_H1024_A16_F4096"])]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
for cfg_file, over in cases:
    cfg = Config(cfg_file, ["OPTIM.BATCH_SIZE", B] + over)
    torch.manual_seed(0)
    model = PretrainingModelFactory.from_config(cfg).cuda().train()
    tr = Trainer(model, cfg)
    batch = {k: v.cuda() for k, v in synth_host_batch(B, pin=False).items()}
    losses = []
    for i in range(4):
        losses.append(float(tr.step(batch).sum()))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for i in range(5):
        tr.step(batch)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{cfg_file} {over}: params {sum(p.numel() for p in model.parameters())/1e6:.1f}M losses {[round(l,3) for l in losses]} "
          f"{ms:.1f} ms/step {B/ms*1e3:.0f} pairs/s grad_norm {float(tr.grad_norm):.3f}", flush=True)
    del tr, model
    torch.cuda.empty_cache()
