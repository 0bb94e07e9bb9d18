from virtex_b200.optim import Lookahead  # noqa: F401
from . import lookahead, lr_scheduler  # noqa: F401
