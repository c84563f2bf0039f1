from .convnext import *  # noqa: F401,F403
from .efficientnet import *  # noqa: F401,F403
from .resnet import *  # noqa: F401,F403
from .swin import *  # noqa: F401,F403
from .vit import *  # noqa: F401,F403
