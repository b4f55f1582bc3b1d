from .config import MIMO_LANGUAGE_MODULE_KEY, MimoModelConfig  # noqa: F401
from .model import MimoModel  # noqa: F401
from .submodules import AudioModalitySubmodules, ModalitySubmodules, VisionModalitySubmodules  # noqa: F401
