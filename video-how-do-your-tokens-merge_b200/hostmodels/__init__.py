"""Plain-PyTorch, eval-oriented restatements of the four video transformers the reference
patches (slowfast/models/{videomae_video_model_builder,timesformer,motionformer_*,vivit_*}.py).
They exist so the token-merging path has callers on a box without slowfast/timm/fvcore:
same module/parameter names as the reference models (state dicts interchange), attention
and MLP through torch library kernels.  They are NOT the product -- the product is the
``tome`` package patched into them."""
from .videomae import VideoMAE, videomae_vit_base_patch16_224  # noqa: F401
from .timesformer import TimeSformer  # noqa: F401
from .motionformer import Motionformer  # noqa: F401
from .vivit import ViViT, VivitConfig  # noqa: F401
