cd $GRAFT_REPO_ROOT
python - <<'PY'
import torch, sys, os
sys.path.insert(0, os.getcwd())
from improver_b200 import engine
g = torch.Generator(device="cuda").manual_seed(4)
d = torch.rand((18, 970, 1042), generator=g, device="cuda")
for _ in range(2): engine.neighbourhood_percentiles(d, 10, (5, 25, 50, 75, 95))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): out, st = engine.neighbourhood_percentiles(d, 10, (5, 25, 50, 75, 95))
e1.record(); torch.cuda.synchronize()
print("percentiles ms", e0.elapsed_time(e1) / 3)
PY
