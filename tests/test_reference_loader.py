"""CPU: the REFERENCE's own loader helpers, extracted unmodified from /root/reference, driven against the drop-in classes.

`app.load_model` (app.py:1327-1769) cannot be imported here (flask is absent), so its nested helpers
`_normalize_state_dict_keys` (:1413-1432), `_safe_load_state_dict` (:1476-1488) and `_load_stats` (:1490-1528) are
lifted out of the source with `ast` and executed as they are; `InferenceAgent` (src/agent_system.py:66-117) is lifted the
same way and bound to the drop-in `PretrainedBackboneDetector`.  Build-container test: skipped where /root/reference is
absent (the GPU box)."""
import ast
import logging
import os
from abc import ABC, abstractmethod
from datetime import datetime
from typing import Any, Tuple

import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "app.py")), reason="reference tree not present")


def _lift(path, names, ns):
    """exec the (possibly nested) function / class definitions called `names` from `path`, unchanged, into `ns`."""
    tree = ast.parse(open(path).read())
    found = {}
    for node in ast.walk(tree):
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names and node.name not in found:
            found[node.name] = node
    assert set(found) == set(names), f"missing in {path}: {set(names) - set(found)}"
    exec(compile(ast.Module([found[n] for n in names], []), path, "exec"), ns)
    return ns


@pytest.fixture(scope="module")
def loader():
    return _lift(os.path.join(REF, "app.py"), ["_normalize_state_dict_keys", "_safe_load_state_dict", "_load_stats"], {"torch": torch})


def _fresh():
    from deepfake_video_detection_b200 import PretrainedBackboneDetector
    return PretrainedBackboneDetector("efficientnet_b0", pretrained=False, num_classes=2, dropout_rate=0.5, use_temporal_attention=True)


def test_reference_load_stats_and_safe_load_on_the_dropin(loader, synth_sd):
    model = _fresh()
    assert not hasattr(model, "models")                              # app.py:2121 selects the ensemble path on this attribute
    stats = loader["_load_stats"](model, synth_sd)                   # app.py:1717
    assert stats["model_keys"] == 366 and stats["ckpt_keys"] == 366
    assert stats["matched"] == 366 and stats["mismatched"] == 0 and stats["missing"] == 0 and stats["unexpected"] == 0
    assert stats["match_ratio"] == 1.0 and stats["match_ratio"] >= 0.80          # the gate of app.py:1735-1738
    res = loader["_safe_load_state_dict"](model, synth_sd)           # app.py:1718
    assert list(res.missing_keys) == [] and list(res.unexpected_keys) == []
    for k, v in model.state_dict().items():
        assert torch.equal(v, synth_sd[k]), k


def test_reference_prefix_normalisation_then_load(loader, synth_sd):
    """DataParallel / wrapper prefixes, also stacked ('model.module.'), are stripped by the reference before the load; the
    engine's packer strips them the same way (engine.normalize_key), so a FrameScorer built on a wrapped dict packs too."""
    from deepfake_video_detection_b200.engine import normalize_key
    wrapped = {"model.module." + k: v for k, v in synth_sd.items()}
    sd = loader["_normalize_state_dict_keys"](wrapped)
    assert list(sd) == list(synth_sd)
    assert [normalize_key(k) for k in wrapped] == list(synth_sd)
    assert loader["_load_stats"](_fresh(), sd)["match_ratio"] == 1.0


def test_reference_gate_rejects_a_foreign_checkpoint(loader, synth_sd):
    """A checkpoint of another architecture (shape mismatches / missing keys) falls under the 0.80 match-ratio gate exactly
    as it does with the reference class: shapes and key names of the drop-in are the reference's."""
    model = _fresh()
    foreign = {k: (v if i % 3 else torch.zeros(tuple(v.shape) + (2,))) for i, (k, v) in enumerate(synth_sd.items()) if "se." not in k}
    stats = loader["_load_stats"](model, foreign)
    assert stats["match_ratio"] < 0.80 and stats["mismatched"] > 0 and stats["missing"] == 64       # 16 blocks x 4 SE tensors
    before = {k: v.clone() for k, v in model.state_dict().items()}
    res = loader["_safe_load_state_dict"](model, foreign)            # only the shape-compatible keys are copied, nothing raises
    assert 0 < len(res.missing_keys) <= 366 - stats["matched"]     # BatchNorm fills an absent num_batches_tracked itself
    changed = sum(not torch.equal(before[k], v) for k, v in model.state_dict().items())
    assert 0 < changed <= stats["matched"]


def test_reference_inference_agent_strict_load(tmp_path, synth_sd):
    """agent_system.py:82-91: construct, `.to(device).eval()`, torch.load of a RAW state_dict, strict load_state_dict."""
    from deepfake_video_detection_b200 import PretrainedBackboneDetector
    ns = {"ABC": ABC, "abstractmethod": abstractmethod, "Any": Any, "Tuple": Tuple, "datetime": datetime, "torch": torch,
          "logger": logging.getLogger("agent_system"), "PretrainedBackboneDetector": PretrainedBackboneDetector}
    _lift(os.path.join(REF, "src", "agent_system.py"), ["Agent", "InferenceAgent"], ns)
    path = tmp_path / "checkpoint_best_efficientnet_b0.pt"
    torch.save(synth_sd, path)
    agent = ns["InferenceAgent"](str(path), backbone_name="efficientnet_b0", device="cpu")
    assert not agent.model.training
    for k, v in agent.model.state_dict().items():
        assert torch.equal(v, synth_sd[k]), k
    bad = dict(synth_sd)
    bad.pop("fc2.bias")
    torch.save(bad, path)
    with pytest.raises(RuntimeError):                                # strict load: a missing key is an error, as in the reference
        ns["InferenceAgent"](str(path), device="cpu")
    # no CPU fallback: the agent's forward on a CPU tensor raises instead of silently running eager PyTorch
    with pytest.raises(RuntimeError):
        agent.process(torch.zeros(1, 2, 3, 224, 224))
