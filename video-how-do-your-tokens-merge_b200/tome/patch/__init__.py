"""``tome.patch.<model>(model)`` and ``tome.patch.duplicate_<model>(model, layer, quantity)`` for the four
video transformers -- the public names of the reference's tome/patch package, bound to the sm_100a path."""
from . import motionformer as _motionformer
from . import timesformer as _timesformer
from . import videomae as _videomae
from . import vivit as _vivit

_MODELS = {"videomae": _videomae, "timesformer": _timesformer, "motionformer": _motionformer, "vivit": _vivit}
__all__ = []
for _name, _module in _MODELS.items():
    globals()[_name] = _module.apply_patch
    globals()["duplicate_" + _name] = _module.apply_duplicate_patch
    __all__ += [_name, "duplicate_" + _name]
del _name, _module
