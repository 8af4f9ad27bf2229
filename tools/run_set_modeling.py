#!/usr/bin/env python
"""Run the reference's OWN set-modeling training script (experiments/set_modeling/train.py:57-83 ->
general/train.py:363-444 ``start_training`` -> ``TrainTemplate.train_model`` :80-254) for a few iterations and record what
it computed - either as the unmodified CPU reference, or with the sm_100a drop-in layers patched in by
``categoricalnf_b200.install`` (nothing in the checkout is edited either way).

    python tools/run_set_modeling.py --impl reference --ref-root baseline/_ref --out ref.json --state-out init.pt -- <train.py args>
    python tools/run_set_modeling.py --impl b200      --ref-root baseline/_ref --out gpu.json --state-in  init.pt -- <train.py args>

The script is executed with ``runpy`` as ``__main__`` (its argparse, ``args_to_params``, ``TrainSetModeling``, RAdam, gradient
clipping, LR scheduler, data-dependent initialisation, initial evaluation and final test all run as upstream wrote them).
Test hooks, installed around the reference's classes without touching their code:
  * after ``TrainTemplate.__init__``: the freshly built model's state dict is saved (``--state-out``) or replaced
    (``--state-in``), then all host RNGs are re-seeded - so both implementations start from the same parameters even
    though their constructors consume random numbers in a different order;
  * ``TaskTemplate.train_step`` / ``TaskTemplate.eval`` are wrapped to record every training loss and evaluation result.
``--impl reference`` hides the GPUs (``CUDA_VISIBLE_DEVICES=""``): the reference's CPU path, as BASELINE config 1 names it.
``--impl b200`` sets ``CNF_B200_HOST_NOISE=1``: the categorical encoding draws its uniform noise from torch's CPU generator
exactly where the reference does (layers/categorical_encoding/linear_encoding.py:78), so both runs see the same noise.
This file is test / benchmark infrastructure (tests/test_gpu_reference_training.py, bench.py --impl reference).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
import types


def stub_matplotlib():
    """general/task.py:8 imports matplotlib.pyplot without using it on this path; the image has no matplotlib."""
    try:
        import matplotlib  # noqa: F401
        return
    except ImportError:
        pass
    mpl = types.ModuleType("matplotlib")
    mpl.use = lambda *a, **k: None
    pyplot, colors = types.ModuleType("matplotlib.pyplot"), types.ModuleType("matplotlib.colors")
    colors.hsv_to_rgb = lambda x: x
    mpl.pyplot, mpl.colors = pyplot, colors
    sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": pyplot, "matplotlib.colors": colors})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", choices=["reference", "b200"], required=True)
    ap.add_argument("--ref-root", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--state-out", default=None)
    ap.add_argument("--state-in", default=None)
    ap.add_argument("--final-state-out", default=None)
    ap.add_argument("--reseed", type=int, default=1234)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("train_args", nargs="*")
    args = ap.parse_args()
    ref_root = os.path.abspath(args.ref_root)
    repo_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

    if args.impl == "reference":
        os.environ["CUDA_VISIBLE_DEVICES"] = ""
    else:
        os.environ["CNF_B200_HOST_NOISE"] = "1"
    import numpy as np
    import random
    import torch
    if args.threads > 0:
        torch.set_num_threads(args.threads)
    stub_matplotlib()
    if args.impl == "b200":
        sys.path.insert(0, repo_root)
        import categoricalnf_b200.install as cnf_install
        cnf_install.install(ref_root)
        from categoricalnf_b200 import ops
    elif ref_root not in sys.path:
        sys.path.insert(0, ref_root)

    import general.task as gtask
    import general.train as gtrain

    record = {"impl": args.impl, "device": None, "train_loss": [], "eval": [], "step_seconds": []}

    orig_init = gtrain.TrainTemplate.__init__

    def init_hook(self, *a, **k):
        orig_init(self, *a, **k)
        record["device"] = str(next(self.model.parameters()).device)
        record["model_class"] = type(self.model).__module__ + "." + type(self.model).__name__
        record["layer_modules"] = sorted({type(m).__module__ for m in self.model.modules()})
        record["num_parameters"] = sum(p.numel() for p in self.model.parameters())
        if args.state_out:
            torch.save({k_: v.detach().cpu() for k_, v in self.model.state_dict().items()}, args.state_out)
        if args.state_in:
            missing = self.model.load_state_dict(torch.load(args.state_in, map_location="cpu"), strict=True)
            record["state_in"] = str(missing)
        np.random.seed(args.reseed)
        random.seed(args.reseed)
        torch.manual_seed(args.reseed)
        main.trainer = self

    gtrain.TrainTemplate.__init__ = init_hook

    orig_step = gtask.TaskTemplate.train_step

    def step_hook(self, iteration=0):
        t0 = time.perf_counter()
        loss = orig_step(self, iteration=iteration)
        record["train_loss"].append(float(loss.item()))
        record["step_seconds"].append(time.perf_counter() - t0)        # forward only (the backward follows in train_model)
        return loss

    gtask.TaskTemplate.train_step = step_hook

    orig_eval = gtask.TaskTemplate.eval

    def eval_hook(self, *a, **k):
        loss_metric, detailed = orig_eval(self, *a, **k)
        record["eval"].append({"nll": float(detailed["negative_log_likelihood"]), "bpd": float(detailed["bpd"])})
        return loss_metric, detailed

    gtask.TaskTemplate.eval = eval_hook

    import runpy
    script = os.path.join(ref_root, "experiments", "set_modeling", "train.py")
    sys.argv = [script] + list(args.train_args)
    cwd = os.getcwd()
    os.chdir(os.path.dirname(script))         # the script appends "../../" to sys.path
    t0 = time.perf_counter()
    try:
        runpy.run_path(script, run_name="__main__")
    finally:
        os.chdir(cwd)
    record["wall_seconds"] = time.perf_counter() - t0
    trainer = getattr(main, "trainer", None)
    if trainer is not None:
        sd = {k_: v.detach().double().cpu() for k_, v in trainer.model.state_dict().items()}
        record["final_param_abs_sum"] = {k_: float(v.abs().sum()) for k_, v in sd.items()}
        if args.final_state_out:
            torch.save({k_: v.float() for k_, v in sd.items()}, args.final_state_out)
    if args.impl == "b200":
        record["cnf_launches"] = ops.launch_count()
        record["param_epoch"] = ops.param_epoch()
    with open(args.out, "w") as f:
        json.dump(record, f)


if __name__ == "__main__":
    main()
