"""Layer-by-layer comparison of the engine's backbone against the oracle (developer tool, run under gpurun)."""
import sys
sys.path.insert(0, ".")
import torch
from oracle import virtex_oracle as O
from tests.test_gpu_parity import build_model, rel

spec = O.Spec(hidden=128, layers=1, heads=2, ffn=256)
state = O.synth_state(spec, 5, bn3_gain=0.25)
model = build_model(spec, state)
B = 4
batch = O.synth_batch(B, seed=3)
eng = model.engine
model.train()
feat, h, w = eng.backbone_forward(batch["image"].cuda(), training=True)
torch.cuda.synchronize()
rec = {}
with torch.no_grad():
    ref = O.backbone_forward(state, batch["image"], spec, training=True, record=rec, emulate_bf16=True)


def nhwc(t):
    return t.permute(0, 2, 3, 1).reshape(-1, t.shape[1])


tape = eng._tape
print("stem.y", rel(tape["stem"]["y"], nhwc(rec["stem.y"])))
print("stem.pool", rel(eng.ws.flat["stem.pool"][: rec["stem.pool"].numel()].view(-1, 64), nhwc(rec["stem.pool"])))
for r in tape["blocks"]:
    q = r["name"] + "."
    print(r["name"], "y1", f'{rel(r["y1"], nhwc(rec[q + "y1"])):.5f}', "a1", f'{rel(r["a1"], nhwc(rec[q + "a1"])):.5f}',
          "y2", f'{rel(r["y2"], nhwc(rec[q + "y2"])):.5f}', "out", f'{rel(r["out"], nhwc(rec[q + "out"])):.5f}')

import os
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
    if os.path.exists(f):
        print(f, open(f).read().strip())
