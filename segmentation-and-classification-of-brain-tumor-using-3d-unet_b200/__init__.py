"""B200-native hot path of the enhanced 3D U-Net: drop-in `nn.Module` / loss / metric API of the reference, executed by
hand-written sm_100a CUDA kernels behind a C ABI (libb3d.so).  See DESIGN.md / INTEGRATION.md."""
from .modules import UNet3D, DoubleConv3D, AttentionGate3D, BrainTumorClassifier  # noqa: F401
from .losses import (CombinedLoss3D, TverskyLoss3D, DeepSupervisionLoss3D, CombinedLoss, DiceLoss,  # noqa: F401
                     FocalLoss)
from .metrics import (calculate_dice_score, dice_score, confusion_matrix, segment, tumor_volumes, classify,  # noqa: F401
                      CLASS_NAMES)
from .graph import GraphedTrainStep, GraphedInference  # noqa: F401
from .optim import FusedAdamW, make_adamw, make_scheduler  # noqa: F401
from . import preprocess  # noqa: F401,E402
from .preprocess import preprocess_image, preprocess_segmentation, preprocess_case, draw_augmentation, apply_augmentations  # noqa: F401,E402
