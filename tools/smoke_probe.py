"""Diagnostics: the smoke-size parity metrics (EfficientNet-B0, tiny batch) for both dtypes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import engine_checks as EC
for (b, r, dt) in [(4, 64, "bf16"), (4, 64, "fp16"), (8, 64, "bf16"), (16, 96, "bf16")]:
    rep = EC.run_parity("efficientnet_b0", b, r, r, dtype=dt, steps=1)
    e = rep["steps"][0]["emul"]
    print(b, r, dt, {k: round(v, 5) for k, v in e.items() if isinstance(v, float)}, flush=True)
