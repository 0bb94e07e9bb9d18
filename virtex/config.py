from virtex_b200.config import *  # noqa: F401,F403
from virtex_b200.config import Config, CfgNode  # noqa: F401
