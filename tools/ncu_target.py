"""One eager forward+backward(+optimizer) of the native engine: the target process for `ncu` captures.
    ncu --set full --clock-control none --import-source on -k regex:<pattern> -c <n> -o gpurun_out/<name> python tools/ncu_target.py [batch]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from deepfake_detection_b200.arch import get_spec  # noqa: E402
from deepfake_detection_b200.trainer import Trainer  # noqa: E402
from deepfake_detection_b200.models import init_state_dict  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
arch = sys.argv[2] if len(sys.argv) > 2 else "efficientnet_b0"
res = int(sys.argv[3]) if len(sys.argv) > 3 else 224
tr = Trainer(arch, batch, res, res, dtype="bf16", use_graph=False)
tr.load_state_dict(init_state_dict(get_spec(arch), seed=42))
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(batch, 3, res, res, device="cuda", generator=g)
y = torch.randint(0, 2, (batch,), device="cuda", generator=g)
for _ in range(int(os.environ.get("NCU_STEPS", "1"))):
    tr.train_step(x, y)
torch.cuda.synchronize()
print("loss", float(tr.engine.loss))
