from .byol import (  # noqa: F401
    EarlyStopping,
    byol_step,
    check_audio_tensor,
    check_audio_tensors,
    evaluate_embedding_similarity,
    train_one_epoch,
    validate_model,
)
from .distributed import GradArena, init_distributed, wrap_data_parallel  # noqa: F401
from .optim import FusedAdamWEma  # noqa: F401
from .emotion import ccc_loss, compute_ccc, emotion_dim_step, train_one_epoch_dimensional  # noqa: F401
