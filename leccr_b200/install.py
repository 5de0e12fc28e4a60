"""Rebind the reference's hot-path callables to the B200 implementations.

    import leccr_b200.install as inst
    inst.install()                      # patches models.xvlm / models.xvlm_video if importable
    inst.install_scripts(image_module)  # patches evaluation_coarse / itm_eval of an imported task script

The reference looks these names up at call time (module globals, class attribute, `self.allgather`), so
rebinding is all a maintainer has to do; nothing else in the reference changes.
"""
import importlib

from .allgather import AllGather as _AllGather  # (the package attribute `allgather` is the function, not the module)
from .allgather import allgather as _allgather
from . import caption_loss as _cl
from . import contrastive as _ct
from .dstl_loss import dstl_loss as _dstl_loss  # (same shadowing: the package attribute is the function)
from . import evaluation as _ev
from . import features as _ft


def _patch_xvlm(mod, base_name):
    mod.AllGather = _AllGather
    mod.allgather = _allgather
    base = getattr(mod, base_name, None)
    if base is not None:
        base.get_contrastive_loss = _ct.get_contrastive_loss
        # models/xvlm.py:241 / models/xvlm_video.py:260: same name, the video variant pools with a frame mask
        base.get_features = _ft.get_features_video if base_name == "XVLMBase_video" else _ft.get_features


def install(modules=("models.xvlm", "models.xvlm_video")):
    """Patch the model modules that are importable; returns the list of patched module names."""
    done = []
    for name, base in zip(modules, ("XVLMBase", "XVLMBase_video")):
        try:
            mod = importlib.import_module(name)
        except Exception:
            continue
        _patch_xvlm(mod, base)
        done.append(name)
    return done


def install_caption_loss(modules=("models.model_retrieval_caption", "models.video_model_retrieval_caption")):
    """Rebind RetrievalModel.get_caption_contrastive_loss (models/model_retrieval_caption.py:145,
    models/video_model_retrieval_caption.py:171) and RetrievalModel.dstl_loss (:94) in the model modules that
    are importable."""
    done = []
    for name in modules:
        try:
            mod = importlib.import_module(name)
        except Exception:
            continue
        cls = getattr(mod, "RetrievalModel", None)
        if cls is not None and hasattr(cls, "get_caption_contrastive_loss"):
            cls.get_caption_contrastive_loss = _cl.get_caption_contrastive_loss
            if hasattr(cls, "dstl_loss"):
                cls.dstl_loss = _dstl_loss  # models/model_retrieval_caption.py:94
            if hasattr(cls, "caption_vision_loss"):
                cls.caption_vision_loss = _ct.caption_vision_loss  # models/model_retrieval_caption.py:118
            done.append(name)
    return done


def install_scripts(image_module=None, video_module=None):
    """Patch the task scripts' evaluation entry points (resolved as module globals in main())."""
    if image_module is not None:
        def evaluation_coarse(model, data_loader, tokenizer, device, config):
            args = getattr(image_module, "args", None)
            return _ev.evaluation_coarse(model, data_loader, tokenizer, device, config,
                                         distributed=bool(getattr(args, "distributed", False)),
                                         clip_tokenizer=getattr(image_module, "clip_tokenizer", None))
        image_module.evaluation_coarse = evaluation_coarse
        image_module.itm_eval = _ev.itm_eval
    if video_module is not None:
        def evaluation_coarse_v(model, data_loader, tokenizer, device, config, alpha=0.9):
            args = getattr(video_module, "args", None)
            return _ev.evaluation_coarse_video(model, data_loader, tokenizer, device, config, alpha=alpha,
                                               distributed=bool(getattr(args, "distributed", False)),
                                               clip_tokenizer=getattr(video_module, "clip_tokenizer", None))
        video_module.evaluation_coarse = evaluation_coarse_v
        video_module.itm_eval = _ev.itm_eval
