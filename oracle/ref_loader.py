"""TEST INFRASTRUCTURE -- loads the reference's own hot-path functions from /root/reference/LECCR.

Only usable in the build container (the GPU box has no /root/reference).  It is how the oracle is
pinned: oracle/make_golden.py runs these functions on seeded synthetic embeddings and commits the
outputs under tests/golden/.  The reference does not import as shipped (missing timm, ruamel.yaml,
ftfy, two unshipped modules, transformers.AdamW removed), so empty stub modules are installed first;
none of them is on the arithmetic path.
"""
import importlib.machinery
import os
import sys
import types

REF_ROOT = "/root/reference/LECCR"


def available():
    return os.path.isdir(REF_ROOT)


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__path__ = []
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


_loaded = {}


def load():
    """Returns a dict with AllGather, get_contrastive_loss, image / video evaluation_coarse, norm_score, itm_eval."""
    if _loaded:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present; golden fixtures are generated in the build container")
    import torch
    import yaml

    ident = lambda *a, **k: (a[0] if a else None)
    _stub("timm")
    _stub("timm.models")
    _stub("timm.models.vision_transformer", _cfg=lambda **k: {}, PatchEmbed=object)
    _stub("timm.models.registry", register_model=lambda f: f)
    _stub("timm.models.layers", trunc_normal_=ident, DropPath=torch.nn.Identity, to_2tuple=lambda x: (x, x))
    ruamel = _stub("ruamel")
    ruamel.yaml = _stub("ruamel.yaml", load=yaml.load, dump=yaml.dump, Loader=yaml.Loader)
    _stub("ftfy", fix_text=lambda s: s)
    import transformers.optimization as topt

    if not hasattr(topt, "AdamW"):
        topt.AdamW = torch.optim.AdamW
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import importlib

    # dataset/__init__.py:22 imports an unshipped module: provide it just before the package is imported
    pm = types.ModuleType("dataset.pretrain_dataset_multilingual")
    pm.__spec__ = importlib.machinery.ModuleSpec("dataset.pretrain_dataset_multilingual", None)
    for n in ("ImageMultiTextDataset", "RegionMultiTextDataset", "ImageMonoTextDataset", "ParaTextDataset"):
        setattr(pm, n, object)
    sys.modules["dataset.pretrain_dataset_multilingual"] = pm
    bo = types.ModuleType("models.box_ops")
    bo.__spec__ = importlib.machinery.ModuleSpec("models.box_ops", None)
    bo.box_cxcywh_to_xyxy = ident
    bo.generalized_box_iou = ident
    sys.modules["models.box_ops"] = bo

    xvlm = importlib.import_module("models.xvlm")
    _loaded["AllGather"] = xvlm.AllGather
    _loaded["get_contrastive_loss"] = xvlm.XVLMBase.get_contrastive_loss
    try:
        xv = importlib.import_module("models.xvlm_video")
        import models

        models.XVLMBase_video = xv.XVLMBase_video
        _loaded["get_contrastive_loss_video"] = xv.XVLMBase_video.get_contrastive_loss
    except Exception as e:  # the video duplicate is byte-identical logic; not needed for pinning
        _loaded["video_import_error"] = repr(e)
    img = importlib.import_module("image_Retrieval_caption")
    _loaded["image_module"] = img
    _loaded["image_evaluation_coarse"] = img.evaluation_coarse
    _loaded["itm_eval"] = img.itm_eval
    try:
        vid = importlib.import_module("video_Retrieval_caption_double_sim")
        _loaded["video_module"] = vid
        _loaded["video_evaluation_coarse"] = vid.evaluation_coarse
        _loaded["norm_score"] = vid.norm_score
        _loaded["video_itm_eval"] = vid.itm_eval
    except Exception as e:
        _loaded["video_script_import_error"] = repr(e)
    return _loaded
