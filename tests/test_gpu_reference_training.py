"""The UNMODIFIED reference training script, run twice: as the CPU reference and with the sm_100a drop-ins installed.

``tools/run_set_modeling.py`` executes ``experiments/set_modeling/train.py`` (BASELINE config 1: set shuffling, batch 256,
``--encoding_num_flows 4`` + the README's defaults) from ``baseline/_ref`` as ``__main__`` - argparse, ``TrainSetModeling``,
the data-dependent initialisation, initial evaluation, 20 iterations of ``train_model`` with the reference's RAdam (which
updates weights through ``p.data``), gradient clipping, LR scheduler, final test - once with the GPUs hidden (the reference's
own CPU path) and once through ``categoricalnf_b200.install`` on cuda:0.  Both start from the state dict of the model the
REFERENCE constructed (loaded with ``strict=True``: the App. A key contract) and see the same encoding noise (host RNG,
``CNF_B200_HOST_NOISE``).  The per-iteration training loss, the evaluation / test bits-per-dimension and the trained
parameters must agree: bpd within 1e-3 (north_star).  Skipped when ``baseline/_ref`` did not travel with the repo
(``tools/vendor_reference.sh`` creates it where ``/root/reference`` exists).
"""
import json
import math
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
RUNNER = os.path.join(ROOT, "tools", "run_set_modeling.py")
LOG2E = math.log2(math.e)

pytestmark = pytest.mark.gpu

ITERS = 20
TRAIN_ARGS = ["--dataset", "shuffling", "--set_size", "16", "--max_iterations", str(ITERS), "--batch_size", "256",
              "--encoding_dim", "4", "--encoding_num_flows", "4", "--optimizer", "4", "--learning_rate", "7.5e-4",
              "--cluster", "--debug", "--print_freq", "5", "--no_model_checkpoints", "--seed", "42"]


def run(impl, tmp, extra):
    out = os.path.join(tmp, impl + ".json")
    cmd = [sys.executable, RUNNER, "--impl", impl, "--ref-root", REF, "--out", out] + extra + ["--"] + TRAIN_ARGS + \
        ["--checkpoint_path", os.path.join(tmp, "ckpt_" + impl)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, "%s run failed:\n%s\n%s" % (impl, r.stdout[-3000:], r.stderr[-3000:])
    with open(out) as f:
        return json.load(f)


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "layers", "flows")), reason="baseline/_ref absent (tools/vendor_reference.sh)")
def test_set_modeling_training_matches_cpu_reference(tmp_path):
    tmp = str(tmp_path)
    init = os.path.join(tmp, "init.pt")
    ref = run("reference", tmp, ["--state-out", init])
    gpu = run("b200", tmp, ["--state-in", init])

    assert ref["device"] == "cpu" and gpu["device"].startswith("cuda")
    # the reference run used only reference modules; the patched run got the drop-in layers under the same names
    assert not any(m.startswith("categoricalnf_b200") for m in ref["layer_modules"])
    dropins = [m for m in gpu["layer_modules"] if m.startswith("categoricalnf_b200.layers")]
    assert len(dropins) >= 6, gpu["layer_modules"]
    assert not any(m.startswith("layers.") for m in gpu["layer_modules"])
    assert gpu["model_class"] == ref["model_class"] == "experiments.set_modeling.flow_model.FlowSetModeling"
    assert gpu["num_parameters"] == ref["num_parameters"]
    assert gpu["cnf_launches"] > 100 * ITERS            # the CUDA library did the work
    assert gpu["param_epoch"] >= ITERS                   # every RAdam step invalidated the parameter-derived caches

    assert len(ref["train_loss"]) == len(gpu["train_loss"]) == ITERS
    worst = max(abs(a - b) * LOG2E for a, b in zip(ref["train_loss"], gpu["train_loss"]))
    assert worst <= 1e-3, "training loss diverges: worst |delta bpd| %.3e\nref %s\ngpu %s" % (worst, ref["train_loss"], gpu["train_loss"])
    assert ref["train_loss"][-1] < ref["train_loss"][0] - 0.1          # it did train

    assert len(ref["eval"]) == len(gpu["eval"]) == 2     # initial evaluation + final test
    for a, b in zip(ref["eval"], gpu["eval"]):
        assert abs(a["bpd"] - b["bpd"]) <= 1e-3, (a, b)

    # trained parameters: per-tensor |.|-sums within 1e-3 relative (20 RAdam steps from identical starting points)
    for name, want in ref["final_param_abs_sum"].items():
        got = gpu["final_param_abs_sum"][name]
        assert abs(got - want) <= 1e-3 * abs(want) + 1e-5, "%s: %r vs %r" % (name, got, want)
