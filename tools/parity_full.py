"""Diagnostics: whole-step parity at a BASELINE configuration size (oracle on the box's host cores).
usage: parity_full.py arch batch res dtype [steps] [tame]  -> JSON on stdout and under gpurun_out/"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
torch.set_num_threads(int(os.environ.get("DFD_ORACLE_THREADS", "32")))
import engine_checks as EC
arch, b, res, dt = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 1
t0 = time.time()
rep = EC.run_parity(arch, b, res, res, dtype=dt, steps=steps, tame=len(sys.argv) > 6 and sys.argv[6] == "tame")
rep["wall_s"] = time.time() - t0
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
name = "parity_%s_%d_%d_%s.json" % (arch, b, res, dt)
json.dump(rep, open(os.path.join(ROOT, "gpurun_out", name), "w"), indent=1, default=str)
for i, st in enumerate(rep["steps"]):
    for k in ("emul", "fp32"):
        r = st[k]
        print("step %d %s: loss %.6f/%.6f logits_rel %.3e grad_tot %.3e grad_med %.3e param_worst %s buf_worst %s" % (
            i, k, r["loss_native"], r["loss_oracle"], r["logits_rel"], r["grad_rel_total"], r["grad_rel_median"],
            r["param_rel_worst"][0], r["buffer_rel_worst"][0]))
    print("step %d yard: %s" % (i, st["yard"]))
print("eval_logits_rel %.3e wall %.1fs" % (rep["eval_logits_rel"], rep["wall_s"]))
