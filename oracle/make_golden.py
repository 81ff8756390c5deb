"""Freeze golden vectors from the UNMODIFIED reference (build container only).

    python -m oracle.make_golden

Imports `/root/reference/src/pretrained_detector.py` as-is (over `oracle/timm_standin`, because timm is
not installed — SURVEY.md F3), loads the calibrated synthetic checkpoint strictly (the way
`agent_system.py:89-91` does), feeds it the reference's own tensor prep (`app.py:2084-2086`) and writes

    tests/golden/synth_calib_seed0.npz      data-dependent part of the synthetic checkpoint
    tests/golden/ref_outputs_seed0.npz      reference features / frame_scores / logits / verdict inputs

`/root/reference` does not exist on the GPU box: tests only read the committed `.npz` files.
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "timm_standin"))
sys.path.insert(0, "/root/reference")

from oracle import synth_checkpoint as S   # noqa: E402

# ragged video lengths exercised by the golden set (T=1 edge case included; app.py allows 1..64)
GOLDEN_LENGTHS = (8, 8, 8, 8, 1, 3, 5, 2)
SEED = 0


def reference_imagenet_normalize():
    """`imagenet_normalize` lives in app.py, which cannot be imported (flask missing); exec just that def."""
    import ast
    src = open("/root/reference/app.py").read()
    tree = ast.parse(src)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "imagenet_normalize")
    ns = {"torch": torch}
    exec(compile(ast.Module([fn], []), "/root/reference/app.py", "exec"), ns)
    return ns["imagenet_normalize"]


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    from src.pretrained_detector import PretrainedBackboneDetector   # the unmodified reference class
    ref_norm = reference_imagenet_normalize()

    sd = S.make_checkpoint(SEED, freeze=True)
    model = PretrainedBackboneDetector("efficientnet_b0", pretrained=False, num_classes=2,
                                       dropout_rate=0.5, use_temporal_attention=True)
    missing = model.load_state_dict(sd, strict=True)
    model.eval()
    print("strict load:", missing)

    crops, offsets = S.synth_crops(SEED + 7, len(GOLDEN_LENGTHS), list(GOLDEN_LENGTHS))
    out = {"offsets": offsets, "crops_sha256": np.frombuffer(hashlib.sha256(crops.tobytes()).digest(), np.uint8)}
    logits, scores, feats = [], [], []
    with torch.no_grad():
        for v in range(len(GOLDEN_LENGTHS)):
            faces = crops[offsets[v]:offsets[v + 1]]
            x = torch.from_numpy(faces).permute(0, 3, 1, 2).float() / 255.0      # app.py:2084
            x = ref_norm(x).unsqueeze(0)                                         # app.py:2085-2086
            lg, fs = model(x)                                                    # app.py:2089
            logits.append(lg[0].numpy()); scores.append(fs[0].numpy())
            feats.append(model.backbone(x[0]).numpy())
        # batched call on the four equal-length videos (validate_improvements.py:146-167 contract)
        xb = torch.stack([ref_norm(torch.from_numpy(crops[offsets[v]:offsets[v + 1]]).permute(0, 3, 1, 2).float() / 255.0)
                          for v in range(4)])
        lgb, fsb = model(xb)
        # mean-pool mode (pretrained_detector.py:132-135)
        model.use_temporal_attention = False
        lgm, fsm = model(xb)
    out.update(logits=np.stack(logits), frame_scores=np.concatenate(scores), features=np.concatenate(feats),
               batched_logits=lgb.numpy(), batched_frame_scores=fsb.numpy(),
               meanpool_logits=lgm.numpy(), meanpool_frame_scores=fsm.numpy(),
               prepped_frame0=ref_norm(torch.from_numpy(crops[:1]).permute(0, 3, 1, 2).float() / 255.0).numpy())
    path = os.path.join(S.GOLDEN_DIR, f"ref_outputs_seed{SEED}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})
    print("logits:\n", out["logits"])
    print("feature abs-mean", np.abs(out["features"]).mean(), "max", np.abs(out["features"]).max())
    print("frame_scores[0:8]", out["frame_scores"][:8])


if __name__ == "__main__":
    main()
