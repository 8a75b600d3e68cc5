"""B200-native drop-in for the render path of Loveof1ife7/mini-3d-gaussian-splatting.

The directory name follows the build contract and is not a valid Python identifier; import it
through the `gsplat_b200` shim at the repo root (``from gsplat_b200 import GaussianRenderer``) or
``importlib.import_module("mini-3d-gaussian-splatting_b200")``.
"""
from .renderer import GaussianRenderer, RenderSettings  # noqa: F401
from .scene import Camera, GaussianModel  # noqa: F401

__all__ = ["GaussianRenderer", "RenderSettings", "GaussianModel", "Camera"]
from . import losses, multiview, training  # noqa: F401,E402
from .training import (ConfigManager, DensityController, GaussianOptimizer, GaussianTrainer, LearningRateScheduler, TrainingConfig,  # noqa: F401,E402
                       train_step)
from .io_utils import CameraUtils, IOUtils  # noqa: F401,E402
from .losses import GaussianLoss, SSIMLoss  # noqa: F401,E402
from .math_utils import MathUtils  # noqa: F401,E402
