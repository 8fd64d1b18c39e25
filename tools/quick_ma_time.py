"""Ad-hoc timing of an M-A train step (dev tool)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import cvad_b200
from cvad_b200.ma import CausalAnomalyDetector, MATrainer
import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
prec = sys.argv[2] if len(sys.argv) > 2 else "fp32"
dev = torch.device("cuda:0")
torch.manual_seed(0)
tr = MATrainer(CausalAnomalyDetector(), dev, precision=prec)
tr.model.train()
x = synth.ma_clips(B, 16, 240, 360, 1, True).to(dev)
y = (torch.rand(B) < 0.3).long().to(dev)
for i in range(3):
    comp, _ = tr.train_step(x, y)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
n0 = cvad_b200.ops.LAUNCHES[0]
ev[0].record()
K = 5
for i in range(K):
    comp, _ = tr.train_step(x, y)
ev[1].record()
torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / K
print(f"M-A train step B={B} {prec}: {ms:.2f} ms/step -> {B / ms * 1000:.1f} clips/s; loss {comp.tolist()}; launches/step {(cvad_b200.ops.LAUNCHES[0]-n0)/K}; mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.train_step(x, y)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
