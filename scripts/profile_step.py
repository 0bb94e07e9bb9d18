"""One profiled optimisation step (after warm-up) between cudaProfilerStart/Stop, for ncu --profile-from-start off."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth_host_batch  # noqa: E402
from virtex_b200.config import Config  # noqa: E402
from virtex_b200.factories import PretrainingModelFactory  # noqa: E402
from virtex_b200.trainer import Trainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--config", default="_base_bicaptioning_R_50_L1_H1024.yaml")
args = ap.parse_args()
cfg = Config(args.config, ["OPTIM.BATCH_SIZE", args.batch])
torch.manual_seed(0)
model = PretrainingModelFactory.from_config(cfg).cuda().train()
trainer = Trainer(model, cfg)
batch = {k: v.cuda() for k, v in synth_host_batch(args.batch, pin=False).items()}
for _ in range(args.warmup):
    trainer.step(batch)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(args.steps):
    trainer.step(batch)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled", args.steps, "step(s); workspace GB", trainer.engine.ws.nbytes() / 1e9)
