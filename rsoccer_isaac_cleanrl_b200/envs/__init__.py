from .vss import VSS, load_cfg  # noqa: F401
from .wrappers import CMA, DMA, RecordEpisodeStatisticsTorch, SingleAgent, make_env, random_ou  # noqa: F401
