"""Diagnostics: run-to-run spread of the whole-step parity metrics (fp32 atomics reorder between runs)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import engine_checks as EC
arch, b, res, dt, reps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], int(sys.argv[5])
for r in range(reps):
    rep = EC.run_parity(arch, b, res, res, dtype=dt, steps=2)
    for i, st in enumerate(rep["steps"]):
        fp, yd = st["fp32"], st["yard"]
        print("run %d step %d: logits %.4f (lim %.4f) grad %.4f (lim %.4f) loss %.5f (lim %.5f) param_worst %.4f  eval %.4f" % (
            r, i, fp["logits_rel"], 2.0 * yd["logits_rel"] + 1e-2, fp["grad_rel_total"], 1.5 * yd["grad_rel_total"] + 2e-2,
            abs(fp["loss_native"] - fp["loss_oracle"]), 2.0 * yd["loss_abs"] + 5e-3, fp["param_rel_worst"][0][1], rep["eval_logits_rel"]), flush=True)
