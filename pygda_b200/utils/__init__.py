from .utility import logger
from .mmd import MMD, get_MMD, draw_indices

__all__ = ["logger", "MMD", "get_MMD", "draw_indices"]
