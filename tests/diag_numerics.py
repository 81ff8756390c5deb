"""GPU diagnostic: error of the CUDA path vs the fp32 oracle over many synthetic videos (not a test)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from deepfake_video_detection_b200 import FrameScorer, make_offsets
from deepfake_video_detection_b200.synthetic import load_checkpoint, synth_crops
from oracle import effnet_b0_oracle as O
import bf16_emulation as E

torch.set_num_threads(16)
sd = load_checkpoint(0)
lens = [4] * 24 + [1] * 8 + [8] * 4
crops, offs = synth_crops(int(sys.argv[1]) if len(sys.argv) > 1 else 99, len(lens), lens)
with torch.no_grad():
    f32 = O.trunk_features(sd, O.prep_u8_hwc(crops))
ref_logits = torch.stack([O.attention_pool_head(sd, f32[offs[v]:offs[v+1]][None])[0][0] for v in range(len(lens))])
out = {}
for prec in ("fp16", "bf16"):
    sc = FrameScorer(sd, prec, "cuda")
    logits, scores, feat = sc.score(torch.from_numpy(crops).cuda(), make_offsets(lens, "cuda"), return_features=True)
    feat, logits = feat.cpu(), logits.cpu()
    per_frame = ((feat - f32).norm(dim=1) / f32.norm(dim=1))
    dl = (logits - ref_logits).abs().max(dim=1)[0]
    flips = int(((logits[:, 1] >= logits[:, 0]) != (ref_logits[:, 1] >= ref_logits[:, 0])).sum())
    out[prec] = dict(feat_rel=float((feat - f32).norm() / f32.norm()), per_frame_max=float(per_frame.max()), per_frame_med=float(per_frame.median()),
                     dlogit_max=float(dl.max()), dlogit_med=float(dl.median()), dlogit_by_T={T: float(dl[[i for i, l in enumerate(lens) if l == T]].max()) for T in (1, 4, 8)},
                     flips=flips, logit_absmax=float(ref_logits.abs().max()))
with torch.no_grad():
    orig = E.bf
    E.bf = lambda t: t.to(torch.float16).float()
    emu = E.trunk_features_bf16(sd, O.prep_u8_hwc(crops[:16]))
    E.bf = orig
out["fp16_emulation_feat_rel_first16"] = float((emu - f32[:16]).norm() / f32[:16].norm())
print(json.dumps(out, indent=1))
