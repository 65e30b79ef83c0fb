from .config_utils import get_config, load_config  # noqa: F401
