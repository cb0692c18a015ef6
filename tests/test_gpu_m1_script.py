"""GPU: driver M1 — the UNMODIFIED reference script (baseline/_ref/train_final_voc.py, a verbatim copy) trains on the
drop-in modules.  The script is launched twice under torchrun through baseline/run_script.py on a generated mini-dataset:
once on the reference's own modules (stock PyTorch on the GPU) and once with `--dropin` (dupl_b200's model_dupl / PAR /
losses / cam_helper bound under the reference's module names).  Same seed => same initial parameters
(tests/test_m1_compat.py) and the same batches, so the loss parts the script itself computes every iteration
(avg_meter.add, train_final_voc.py:461-468) must agree iteration by iteration, across the phase-A -> phase-B boundary
(--cam_iters 3) and through real AdamW updates under DistributedDataParallel(find_unused_parameters=True)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline import compat, make_dataset  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not compat.available(), reason="baseline/_ref not installed")]

ITERS, CAM_ITERS = 6, 3


def _run(tmp, tag, dropin, nproc, port):
    root, lists = make_dataset.make_voc_like(os.path.join(tmp, "data"), n_train=16, n_val=2, seed=0)
    trace = os.path.join(tmp, f"trace_{tag}.jsonl")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "baseline", "run_script.py")]
    if dropin:
        cmd.append("--dropin")
    cmd += ["--trace", trace, "train_final_voc.py", "--",
            "--data_folder", root, "--list_folder", lists, "--work_dir", os.path.join(tmp, f"work_{tag}"),
            "--samples_per_gpu", "2", "--num_workers", "2", "--max_iters", str(ITERS), "--cam_iters", str(CAM_ITERS),
            "--log_iters", "2", "--eval_iters", "100000"]
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, (out.stdout[-1500:], out.stderr[-3000:])
    rows = [json.loads(line) for line in open(trace)]
    assert len(rows) == ITERS
    return rows, out


@pytest.mark.parametrize("nproc", [1, 2])
def test_unmodified_script_runs_on_the_dropin_modules_and_tracks_the_stock_run(tmp_path, nproc):
    if torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    ref_rows, _ = _run(str(tmp_path), f"ref{nproc}", False, nproc, 29611 + nproc)
    got_rows, out = _run(str(tmp_path), f"dropin{nproc}", True, nproc, 29621 + nproc)
    print(json.dumps({"reference": ref_rows, "dropin": got_rows}))
    assert "Iter: %d" % ITERS in (out.stdout + out.stderr)          # the script's own logger reached the last iteration
    for r, g in zip(ref_rows, got_rows):
        it = r["iter"]
        assert (r["seg_loss"] == 1.0) == (it < CAM_ITERS)            # phase A before cam_iters, phase B after
        for k, tol in (("cls_loss", 2e-3), ("ptc_loss", 2e-3), ("sim_loss", 2e-3), ("seg_loss", 2e-2)):
            assert abs(g[k] - r[k]) <= tol * max(1.0, abs(r[k])), (it, k, g[k], r[k])
    # the parameters really moved: the classification loss of the last iteration differs from the first
    assert got_rows[-1]["cls_loss"] != got_rows[0]["cls_loss"]
