"""B200-native (sm_100a) implementation of the reference's EfficientNet-B0 frame-scoring hot path.

Public surface (mirrors the reference):
    PretrainedBackboneDetector, EnsembleDetector      src/pretrained_detector.py
    imagenet_normalize, decide                        app.py:1772-1780, 2090-2112
    LogicRNNLSTM, LogicCell, create_model             src/RNNModel.py (temporal head of BASELINE config 3)
    ViTFeatureExtractor, SimpleGCN, DeepfakeModel     src/models.py (ViT-B/16 frame encoder of config 5, ViT + GCN model)
    crop_resize, clamp_boxes                          app.py:1964-1978 (Pillow-exact crop + resize of face boxes)
    FrameScorer, PackedWeights, make_offsets          high-throughput engine over the C ABI (include/dfd_b200.h)
"""
from .decision import decide, imagenet_normalize
from .engine import DEFAULT_PRECISION, FrameScorer, GraphedScorer, PackedWeights, make_offsets
from .pretrained_detector import EnsembleDetector, PretrainedBackboneDetector
from .rnn_model import LogicCell, LogicRNNLSTM, create_model
from .sharding import gather_video_logits, score_videos_sharded, shard_bounds
from .vit_model import DeepfakeModel, SimpleGCN, ViTFeatureExtractor
from .crop_resize import clamp_boxes, crop_resize

__all__ = ["PretrainedBackboneDetector", "EnsembleDetector", "imagenet_normalize", "decide", "FrameScorer",
           "PackedWeights", "GraphedScorer", "make_offsets", "DEFAULT_PRECISION", "shard_bounds", "gather_video_logits",
           "score_videos_sharded", "LogicCell", "LogicRNNLSTM", "create_model", "ViTFeatureExtractor", "SimpleGCN", "DeepfakeModel",
           "crop_resize", "clamp_boxes"]
