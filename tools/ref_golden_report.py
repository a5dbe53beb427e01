"""Error table of the CUDA path against the reference-source goldens (tests/golden/ref_path*.pt).  Run on a GPU box:
python tools/ref_golden_report.py > gpurun_out/ref_golden_accuracy.txt"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import test_gpu_ref_golden as T  # noqa: E402

print(f"{'case':26s} {'b':>4s} {'q':>2s} {'N':>5s} {'temp':>7s} | {'max|d acq|':>11s} {'/max|acq|':>10s} {'max|d mask|':>11s} "
      f"{'max rel d grid':>14s} {'argmax':>10s}")
for fname in ("ref_path.pt", "ref_path_n1024.pt"):
    for c in T._cases(fname)["acq"]:
        out = T._call_abi(c)
        w = c["out"]
        b, G, q = c["b"], c["G"], c["q"] + c["x_pending"]
        acq = out["acq"].cpu()
        da = float((acq - w["acq"]).abs().max())
        dm = float((out["conf_pred_mask"].cpu().reshape(b, G, q) - w["conf_pred_mask"][..., 0]).abs().max())
        tg = w["target_grid"][..., 0]
        dg = float(((out["target_grid"].cpu().reshape(b, G, q) - tg).abs() / tg.abs().clamp_min(1e-3)).max())
        same = int(acq.argmax()) == w["argmax"]
        print(f"{c['name']:26s} {b:4d} {q:2d} {c['N']:5d} {c['temp']:7.0e} | {da:11.3e} {da / float(w['acq'].abs().max()):10.2e} "
              f"{dm:11.3e} {dg:14.3e} {'same' if same else 'DIFFERENT':>10s}")
