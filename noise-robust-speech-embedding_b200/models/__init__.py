from .byol import BYOLSpeechModel, byol_loss  # noqa: F401
from .encoder import (WavLMEncoder, install_b200_frontend, install_sync_free_spec_augment,  # noqa: F401
                      wavlm_large_config)
from .featproj import B200FeatureProjection  # noqa: F401
from .frontend import B200FeatureEncoder  # noqa: F401
from .posconv import B200PositionalConvEmbedding  # noqa: F401
from .multi_layer_heads import PredictionHead, ProjectionHead  # noqa: F401
from .pool import AttentiveStatisticsPooling, Pooling  # noqa: F401
from .emotion import EmotionClassifier  # noqa: F401
