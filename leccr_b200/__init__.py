"""leccr_b200 -- B200-native dense cross-modal similarity stage of LECCR.

Drop-in replacements for the reference's hot path:
  AllGather / allgather, get_contrastive_loss      models/xvlm.py:50-70, 260-292
  evaluation_coarse (image / video double_sim)     image_Retrieval_caption.py:83-163,
                                                   video_Retrieval_caption_double_sim.py:94-190
  itm_eval                                         image_Retrieval_caption.py:261-317
plus the additive fused_eval (similarity + top-k + Recall without materialising N x M).

The arithmetic lives in libleccr_b200.so (hand-written sm_100a CUDA behind a C ABI, see
include/leccr_b200.h).  There is no CPU path: calls raise if the library or a B200 is missing.
"""
__version__ = "0.1.0"

from .allgather import AllGather, allgather  # noqa: E402,F401
from .caption_loss import caption_contrastive_loss, get_caption_contrastive_loss  # noqa: E402,F401
from .contrastive import caption_vision_loss, contrastive_loss, get_contrastive_loss  # noqa: E402,F401
from .dstl_loss import dstl_loss, dstl_loss_gathered  # noqa: E402,F401
from .features import get_features, get_features_video, normalize_rows  # noqa: E402,F401
from .gallery import GallerySearchPlan  # noqa: E402,F401
from .evaluation import (FeatureGallery, FusedEvalPlan, StreamedEvalPlan, double_sim_matrix, evaluation_coarse, evaluation_coarse_video, fused_eval,  # noqa: E402,F401
                         fused_eval_sharded, itm_eval, prepare_gt, score_matrix, topk_gallery_sharded)
