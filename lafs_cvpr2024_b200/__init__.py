"""lafs_cvpr2024_b200 -- B200-native (sm_100a) implementation of the LAFS per-step hot path.

Drop-in objects (same names / signatures / state-dict keys as the reference):
    DINOLoss                               lafs_train.py:626-679
    DINOHead                               vision_transformer.py:265-301 (fused_loss=True: last_layer GEMM fused with DINOLoss)
    ema_update_                            lafs_train.py:610-613 (inline loop in the reference)
    extract_patches_pytorch_gridsample     face_pre_pro/ViT_face.py:1615-1656
    landmark_post                          face_pre_pro/ViT_face.py:1347-1378
    StudentUpdate                          utils.py:132-149 + torch.optim.AdamW + lafs_train.py:610-613 in one call
All compute goes through liblafs_b200.so (C ABI in include/lafs_b200.h); there is no CPU path.
"""
from . import _lib  # noqa: F401
from .dino_loss import DINOLoss  # noqa: F401
from .dino_head import DeferredLogits, DINOHead, dino_head_backward, dino_head_forward  # noqa: F401
from .ema import EmaPlan, ema_update_  # noqa: F401
from .margin_head import ArcFace, CosFace, label_to_shard, shard_bounds  # noqa: F401
from .optim import StudentUpdate, regularized_mask  # noqa: F401
from .patches import (PatchEmbedWeights, embed_backward_weight, extract_patches_pytorch_gridsample,  # noqa: F401
                      extract_tokens, gather_embed, gather_embed_train, landmark_post, new_token_buffer)

from .vit_face import ViT_face_landmark_patch8, face_landmark_4simmin_glo_loc  # noqa: F401,E402

__all__ = ["ViT_face_landmark_patch8", "face_landmark_4simmin_glo_loc", "ArcFace", "CosFace", "label_to_shard", "shard_bounds", "DINOLoss", "DINOHead", "DeferredLogits",
           "dino_head_forward", "dino_head_backward", "EmaPlan", "ema_update_", "extract_patches_pytorch_gridsample", "extract_tokens",
           "landmark_post", "gather_embed", "gather_embed_train", "PatchEmbedWeights", "StudentUpdate", "regularized_mask",
           "embed_backward_weight", "new_token_buffer"]
